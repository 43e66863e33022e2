// te_pool.cpp - env-slot pool for learner threads (host code only; built on the C ABI of te_api.cu).
//
// The reference's A3C steps FLAGS.threads envs from that many Python threads, each in its own rollout loop
// (algorithms/a3c.py:66-72, 110-137).  te_pool_step() is what such a thread calls for its slot of ONE batched handle:
// it queues the slot's action; whichever caller finds no launch in flight becomes the leader, takes every action
// queued so far and advances exactly those slots with one te_step_masked launch; callers that arrive meanwhile are
// served by the next launch.  No lock-step round: a slow learner only delays itself.  A would-be leader waits ONCE,
// for at most `linger_us`, when fewer slots are queued than the previous launch served (bigger batches when all
// learners are fast).  Everything here runs without the Python GIL (ctypes releases it around the call).
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/traffic_b200.h"

struct te_pool {
  te_handle *h;
  int E, I, obs_len, k_ticks, linger_us;
  std::mutex mu;
  std::condition_variable cv;
  bool launching = false;
  int last_batch = 1;
  std::vector<uint8_t> state;            // per slot: 0 idle, 1 queued, 2 in flight, 3 result ready, 4 failed
  std::vector<uint8_t> actions, mask;    // [E][I], [E]
  uint8_t *h_actions = nullptr;          // page-locked staging of one launch
  float *obs = nullptr, *reward = nullptr;
  uint8_t *done = nullptr;
  std::string error;
  uint64_t launches = 0, stepped = 0;
};

extern "C" int te_pool_create(te_handle *h, int32_t k_ticks, int32_t linger_us, te_pool **out) {
  if (!h || !out || k_ticks < 1) return -1;
  te_dims d;
  if (te_get_dims(h, &d)) return -1;
  te_pool *p = new te_pool();
  p->h = h; p->E = d.num_envs; p->I = d.intersections; p->obs_len = d.obs_actor; p->k_ticks = k_ticks;
  p->linger_us = linger_us < 0 ? 0 : linger_us;
  p->state.assign(p->E, 0); p->actions.assign((size_t)p->E * p->I, 0); p->mask.assign(p->E, 0);
  void *a = nullptr, *b = nullptr, *c = nullptr, *e = nullptr;
  if (te_host_alloc((uint64_t)p->E * p->I, &a) || te_host_alloc((uint64_t)p->E * p->obs_len * 4, &b) ||
      te_host_alloc((uint64_t)p->E * p->I * 4, &c) || te_host_alloc((uint64_t)p->E, &e)) { delete p; return -1; }
  p->h_actions = (uint8_t *)a; p->obs = (float *)b; p->reward = (float *)c; p->done = (uint8_t *)e;
  *out = p;
  return 0;
}

extern "C" int te_pool_destroy(te_pool *p) {
  if (!p) return 0;
  te_host_free(p->h_actions); te_host_free(p->obs); te_host_free(p->reward); te_host_free(p->done);
  delete p;
  return 0;
}

// Blocks until slot `slot` has been advanced by one actor step with `action` (uint8[I], non-zero = 1); copies the
// slot's float observation / reward / done flag out.  Returns 0, or < 0 with te_pool_last_error().
extern "C" int te_pool_step(te_pool *p, int32_t slot, const uint8_t *action, float *obs, float *reward, uint8_t *done) {
  if (!p || slot < 0 || slot >= p->E || !action || !obs || !reward || !done) return -1;
  std::unique_lock<std::mutex> lk(p->mu);
  if (p->state[slot] != 0) { p->error = "te_pool_step: slot " + std::to_string(slot) + " is already stepping"; return -1; }
  memcpy(&p->actions[(size_t)slot * p->I], action, p->I);
  p->state[slot] = 1;
  bool lingered = false;
  for (;;) {
    if (p->state[slot] >= 3) break;
    if (!p->launching && p->state[slot] == 1) {
      int queued = 0;
      for (int e = 0; e < p->E; e++) queued += p->state[e] == 1;
      if (!lingered && p->linger_us > 0 && queued < p->last_batch) {
        lingered = true;
        p->cv.wait_for(lk, std::chrono::microseconds(p->linger_us));
        continue;
      }
      // leader: one masked launch for everything queued
      int n = 0;
      for (int e = 0; e < p->E; e++) {
        p->mask[e] = p->state[e] == 1;
        if (p->mask[e]) { p->state[e] = 2; n++; memcpy(p->h_actions + (size_t)e * p->I, &p->actions[(size_t)e * p->I], p->I); }
      }
      p->launching = true;
      p->last_batch = n;
      std::vector<uint8_t> mask = p->mask;   // (the launch runs without the lock: others keep queueing)
      lk.unlock();
      const int rc = te_step_masked(p->h, p->h_actions, mask.data(), p->k_ticks, p->obs, p->reward, p->done, TE_HOST, nullptr);
      lk.lock();
      if (rc < 0) p->error = te_last_error();
      for (int e = 0; e < p->E; e++) if (p->state[e] == 2) p->state[e] = rc < 0 ? 4 : 3;
      p->launching = false;
      p->launches++; p->stepped += n;
      p->cv.notify_all();
      continue;
    }
    p->cv.wait(lk);
  }
  const bool ok = p->state[slot] == 3;
  if (ok) {
    memcpy(obs, p->obs + (size_t)slot * p->obs_len, (size_t)p->obs_len * 4);
    memcpy(reward, p->reward + (size_t)slot * p->I, (size_t)p->I * 4);
    *done = p->done[slot];
  }
  p->state[slot] = 0;
  return ok ? 0 : -1;
}

// TrafficEnv._reset of one slot (init_phase uint8[I]); waits until no launch is in flight.
extern "C" int te_pool_reset(te_pool *p, int32_t slot, const uint8_t *init_phase) {
  if (!p || slot < 0 || slot >= p->E || !init_phase) return -1;
  std::unique_lock<std::mutex> lk(p->mu);
  p->cv.wait(lk, [&] { return !p->launching; });
  std::vector<uint8_t> mask(p->E, 0), phases((size_t)p->E * p->I, 0);
  mask[slot] = 1;
  memcpy(&phases[(size_t)slot * p->I], init_phase, p->I);
  const int rc = te_reset(p->h, mask.data(), phases.data(), TE_HOST, nullptr);   // (under the lock: no launch can start)
  if (rc < 0) p->error = te_last_error();
  return rc;
}

// cars_on_roads of one slot (int32[R]); waits until no launch is in flight.
extern "C" int te_pool_cars(te_pool *p, int32_t slot, int32_t *out) {
  if (!p || slot < 0 || slot >= p->E || !out) return -1;
  std::unique_lock<std::mutex> lk(p->mu);
  p->cv.wait(lk, [&] { return !p->launching; });
  te_dims d;
  te_get_dims(p->h, &d);
  std::vector<int32_t> all((size_t)p->E * d.roads);
  const int rc = te_cars_on_roads(p->h, all.data(), TE_HOST, nullptr);
  if (rc < 0) { p->error = te_last_error(); return rc; }
  memcpy(out, &all[(size_t)slot * d.roads], (size_t)d.roads * 4);
  return 0;
}

extern "C" int te_pool_counters(te_pool *p, uint64_t *launches, uint64_t *stepped) {
  if (!p) return -1;
  std::lock_guard<std::mutex> lk(p->mu);
  if (launches) *launches = p->launches;
  if (stepped) *stepped = p->stepped;
  return 0;
}

extern "C" const char *te_pool_last_error(te_pool *p) { return p ? p->error.c_str() : "null pool"; }
