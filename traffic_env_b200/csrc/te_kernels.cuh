// te_kernels.cuh - sm_100a kernels of the traffic-env simulation step.
//
// One CTA advances one env instance by K physics ticks (one actor step of the reference's
// Repeater wrapper, traffic_test.py:37-56).  The env's per-road ring buffers
// (traffic_env.py:364-366: state[R,10,20], leading[R], lastcar[R]) are staged in shared memory
// as two float planes x[Rp][20], v[Rp][20] - every car has the same archetype
// (traffic_env.py:35-43), so x and v are the only dynamic fields - and flushed once.
//
// Per tick (order of traffic_env.py:224-248, see DESIGN.md "tick phases"):
//   phase A  (a warp owns groups of 8 roads = 160 ring slots = 5 lane-passes)
//     - entry arrivals -> add_car                              (traffic_env.py:274-283, 97-114)
//     - virtual leader x from the light state                  (update_lights, :81-94)
//     - slot 0 <- slot 19 mirror, then Jacobi IDM update of every live slot from its
//       predecessor slot (sim, :50-62; move_cars, :187-212), waiting/detected counts
//     - pops: leading advances past cars with x > length        (advance_finished_cars, :117-135)
//   barrier
//   phase C  (a thread owns a destination road)
//     - popped cars of the unique upstream road are inserted with add_car semantics;
//       whether the insert sees the destination's pre- or post-pop `leading` follows the
//       reference's road-index order (upstream < dest: pre-pop)
//     - tail x of the road for the next tick's lights
//   barrier
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "te_math.cuh"

namespace te {

constexpr int CAP = 20;            // CAPACITY, traffic_env.py:24
constexpr int RING = CAP - 1;      // live ring positions 1..19
constexpr int YELLOW_TICKS = 6;    // traffic_env.py:21
constexpr int GROUP_ROADS = 8;     // 8 roads * 20 slots = 160 = 5 * 32 lanes
constexpr int GROUP_PASSES = 5;
constexpr int MAX_K = 64;

enum : int { F_LEARN_SWITCH = 1, F_REMI = 2, F_AUTO_RESET = 4, F_VALIDATE = 8 };
enum : int { ARR_NONE = 0, ARR_INJECTED = 1, ARR_PHILOX = 2 };

struct EnvScalars {
  float steps;            // np.float32 tick counter of the current episode (traffic_env.py:260,247)
  uint32_t ph_draw;       // next Philox draw index
  uint32_t ph_skip;       // empty ticks left before the next car
  uint32_t reset_count;
  long long sched_cursor; // ticks consumed from the arrival process since creation (never reset)
  int ep_step;            // actor steps in the current episode
  int done;               // last te_step overflowed
  double ep_ret, ep_disc, ep_mult;
};

struct DeviceStats {
  unsigned long long ticks, actor_steps, vehicle_updates, overflows, cars_generated, episodes, seq_fallback_ticks;
  double return_sum, disc_return_sum;
};

struct StepParams {
  int V, r, R, Rp, I, n_entry, G;
  int num_envs;
  float length;
  double det_thr;         // (double)length - 10.0  (traffic_env.py:201: float32 - int64 types as float64)
  int flags, arrival_mode, K, raw, episode_len;
  float gamma;
  IdmConst idm;
  // state (HBM)
  float *x, *v;           // [E][Rp][20]; slot 0 of a row carries packed ring indices (see pack_meta)
  int *elapsed;           // [E][I]
  uint8_t *phase, *passed_dst;  // [E][I]
  EnvScalars *env;
  DeviceStats *stats;
  // topology (HBM, read-only)
  const short *nexts, *up;      // [Rp], -1 = none
  const signed char *entry_idx; // [Rp], index into the entry list or -1
  const short *entry_roads;     // [n_entry]
  // per-call I/O (device pointers)
  const uint8_t *actions;       // [E][I]
  float *obs_f;                 // [E][2r+I]   (fused actor step)
  int *obs_i;                   // [E][2r+2I]  (raw tick)
  float *reward;                // [E][I]
  uint8_t *done;                // [E]
  // arrivals
  const long long *sched_off;   // [E*(horizon+1)]
  const short *sched_roads;
  int horizon;
  const uint32_t *gap_cdf;
  int n_gap;
  uint32_t seed;
  long long env_id_base;
};

// HBM row header: x[road][0] bits = leading | lastcar << 8 | detected << 16, v[road][0] bits = waiting.
__host__ __device__ inline uint32_t pack_meta(int leading, int lastcar, int detected) {
  return (uint32_t)leading | ((uint32_t)lastcar << 8) | ((uint32_t)detected << 16);
}

struct SmemLayout {
  int xs, vs, leadx, tailx, meta, wait, pd, nexts, up, eidx, phase, act, pdst, elapsed, ovf, cnt, snap, tabs, misc, total;
};

__host__ __device__ inline int align_up(int a, int b) { return (a + b - 1) / b * b; }

__host__ __device__ inline SmemLayout make_layout(int Rp, int I, int K, int n_entry) {
  SmemLayout L;
  int o = 0;
  L.xs = o; o += Rp * CAP * 4;
  L.vs = o; o += Rp * CAP * 4;
  L.tabs = o; o += (int)sizeof(PowfTables);            // 512 B, 8-aligned
  L.leadx = o; o += Rp * 4;
  L.tailx = o; o += Rp * 4;
  L.meta = o; o += Rp * 4;
  L.wait = o; o += Rp * 4;
  L.pd = o; o += Rp * 4;                               // passed_acc | detected << 16
  L.elapsed = o; o += align_up(I, 4) * 4;
  L.ovf = o; o += align_up(I, 4) * 4;
  L.snap = o; o += (MAX_K + 1) * 8;                    // Philox (draw, skip) before each tick
  L.misc = o; o += 32;
  L.nexts = o; o += Rp * 2;
  L.up = o; o += Rp * 2;
  L.eidx = o; o += Rp;
  L.phase = o; o += align_up(I, 4);
  L.act = o; o += align_up(I, 4);
  L.pdst = o; o += align_up(I, 4);
  L.cnt = o; o += align_up(K * (n_entry > 0 ? n_entry : 1), 16);
  L.total = align_up(o, 16);
  return L;
}

__device__ __forceinline__ int ring_wrap(int a) { return a >= CAP ? 1 : a; }
__device__ __forceinline__ int ring_count(int ld, int lc) { return lc - ld + (ld > lc ? RING : 0); }

// add_car (traffic_env.py:97-114) for one identical-archetype car at (xin, vin); `chk` is the value
// of leading[road] the reference would see at that moment.  Returns false when the ring is full.
__device__ __forceinline__ bool ring_push(float *xr, float *vr, int chk, int &lc, float xin, float vin,
                                          const IdmConst &c) {
  const int pos = ring_wrap(lc + 1);
  float start = __int_as_float(0x7f800000);
  if (lc != chk) start = __fsub_rn(__fsub_rn(xr[lc], c.len), c.s0);
  if (pos == chk) return false;
  xr[pos] = (start < xin) ? start : xin;
  vr[pos] = vin;
  lc = pos;
  return true;
}

struct Smem {
  float *xs, *vs, *leadx, *tailx;
  uint32_t *meta;        // leading | lastcar << 8 | pre-pop leading << 16 | npop << 24
  int *wait, *pd, *elapsed, *ovf;
  uint32_t *snap;
  int *misc;             // [0] first overflowing tick, [1] tick needing ordered transfers, [2] vehicle updates, [3] overflows, [4] generated, [5] ordered-transfer ticks
  short *nexts, *up;
  signed char *eidx;
  uint8_t *phase, *act, *pdst, *cnt;
  PowfTables *tabs;
};

__device__ __forceinline__ Smem carve(unsigned char *base, const SmemLayout &L) {
  Smem s;
  s.xs = (float *)(base + L.xs); s.vs = (float *)(base + L.vs);
  s.leadx = (float *)(base + L.leadx); s.tailx = (float *)(base + L.tailx);
  s.meta = (uint32_t *)(base + L.meta); s.wait = (int *)(base + L.wait); s.pd = (int *)(base + L.pd);
  s.elapsed = (int *)(base + L.elapsed); s.ovf = (int *)(base + L.ovf); s.snap = (uint32_t *)(base + L.snap);
  s.misc = (int *)(base + L.misc); s.nexts = (short *)(base + L.nexts); s.up = (short *)(base + L.up);
  s.eidx = (signed char *)(base + L.eidx); s.phase = base + L.phase; s.act = base + L.act;
  s.pdst = base + L.pdst; s.cnt = base + L.cnt; s.tabs = (PowfTables *)(base + L.tabs);
  return s;
}

// Move the popped cars of road u to the tail of road d (advance_finished_cars -> add_car,
// traffic_env.py:126-132).  u < d: the reference inserts before d's own pops of this tick.
__device__ __forceinline__ void transfer(const StepParams &p, const Smem &s, int u, int d, int t) {
  const uint32_t mu = s.meta[u];
  const int np = mu >> 24;
  if (np == 0) return;
  const uint32_t md = s.meta[d];
  const int dld = md & 0xff, dlp = (md >> 16) & 0xff;
  int dlc = (md >> 8) & 0xff;
  const int chk = (u < d) ? dlp : dld;
  int slot = (mu >> 16) & 0xff;
  float *xd = s.xs + d * CAP, *vd = s.vs + d * CAP;
  for (int k = 0; k < np; k++) {
    slot = ring_wrap(slot + 1);
    const float xin = __fsub_rn(s.xs[u * CAP + slot], p.length);
    const float vin = s.vs[u * CAP + slot];
    if (!ring_push(xd, vd, chk, dlc, xin, vin, p.idm)) {
      if (d < p.r) atomicAdd(&s.ovf[d % p.V], 1);
      atomicAdd(&s.misc[3], 1);
      atomicMin(&s.misc[0], t);
    }
  }
  s.meta[d] = (md & 0xffff00ffu) | ((uint32_t)dlc << 8);
}

constexpr int MAX_THREADS = 512;

__global__ void __launch_bounds__(MAX_THREADS) te_step_kernel(const StepParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const SmemLayout L = make_layout(p.Rp, p.I, p.K, p.n_entry);
  const Smem s = carve(smem_raw, L);
  const int env = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const unsigned FULL = 0xffffffffu;
  const float INF = __int_as_float(0x7f800000);
  const bool learn_switch = (p.flags & F_LEARN_SWITCH) != 0;
  EnvScalars *es = p.env + env;

  // ------------------------------------------------------------ prologue: stage the env
  {
    const float4 *gx = reinterpret_cast<const float4 *>(p.x + (size_t)env * p.Rp * CAP);
    const float4 *gv = reinterpret_cast<const float4 *>(p.v + (size_t)env * p.Rp * CAP);
    float4 *sx = reinterpret_cast<float4 *>(s.xs), *sv = reinterpret_cast<float4 *>(s.vs);
    const int n4 = p.Rp * (CAP / 4);
    for (int i = tid; i < n4; i += blockDim.x) { sx[i] = gx[i]; sv[i] = gv[i]; }
  }
  for (int i = tid; i < p.Rp; i += blockDim.x) {
    s.nexts[i] = p.nexts[i]; s.up[i] = p.up[i]; s.eidx[i] = p.entry_idx[i];
  }
  if (tid < 64) reinterpret_cast<unsigned long long *>(s.tabs)[tid] =
      reinterpret_cast<const unsigned long long *>(&g_powf_tables)[tid];
  for (int i = tid; i < p.I; i += blockDim.x) {
    // phase / elapsed update of the first tick (traffic_env.py:225-232); later ticks of the same
    // actor step repeat the same action and are derived in closed form (phase_at / elapsed_at).
    int ph = p.phase[(size_t)env * p.I + i] != 0;
    const int act = p.actions[(size_t)env * p.I + i] != 0;
    int el = p.elapsed[(size_t)env * p.I + i];
    int change;
    if (learn_switch) { change = act; ph ^= act; } else { change = ph ^ act; ph = act; }
    el = (el + 1) * (change ? 0 : 1);
    s.phase[i] = (uint8_t)ph; s.act[i] = (uint8_t)act; s.elapsed[i] = el;
    s.pdst[i] = p.passed_dst[(size_t)env * p.I + i]; s.ovf[i] = 0;
  }
  if (tid == 0) { s.misc[0] = 0x7fffffff; s.misc[1] = -1; s.misc[2] = 0; s.misc[3] = 0; s.misc[4] = 0; s.misc[5] = 0; }
  if (warp == 0) {
    // arrivals of the K ticks as per-tick, per-entry-road counts (cars are identical, so the
    // order of arrivals within a tick only matters per road, where it is preserved)
    const int ncnt = p.K * p.n_entry;
    for (int i = lane; i < ncnt; i += 32) s.cnt[i] = 0;
    __syncwarp();
    if (p.arrival_mode == ARR_INJECTED) {
      const long long cur = es->sched_cursor;
      for (int t = lane; t < p.K; t += 32) {
        const long long tick = cur + t;
        if (tick < p.horizon) {
          const long long *off = p.sched_off + (size_t)env * (p.horizon + 1) + tick;
          for (long long k = off[0]; k < off[1]; k++) {
            const int idx = p.entry_idx[p.sched_roads[k]];
            if (idx >= 0 && s.cnt[t * p.n_entry + idx] < 255) s.cnt[t * p.n_entry + idx]++;
          }
        }
      }
    } else if (p.arrival_mode == ARR_PHILOX && lane == 0) {
      uint32_t draw = es->ph_draw, skip = es->ph_skip;
      const uint32_t k0 = p.seed, k1 = (uint32_t)(p.env_id_base + env);
      for (int t = 0; t < p.K; t++) {
        s.snap[2 * t] = draw; s.snap[2 * t + 1] = skip;
        for (;;) {
          if (skip > 0) { skip--; break; }
          uint32_t o[4];
          philox4x32_10(draw, 0, 0, 0, k0, k1, o);
          draw++;
          const int idx = (int)__umulhi(o[1], (uint32_t)p.n_entry);
          if (s.cnt[t * p.n_entry + idx] < 255) s.cnt[t * p.n_entry + idx]++;
          skip = gap_from_u32(p.gap_cdf, p.n_gap, o[0]);
        }
      }
      s.snap[2 * p.K] = draw; s.snap[2 * p.K + 1] = skip;
    }
  }
  __syncthreads();
  for (int road = tid; road < p.Rp; road += blockDim.x) {
    const uint32_t w0 = __float_as_uint(s.xs[road * CAP]);
    const int ld = w0 & 0xff, lc = (w0 >> 8) & 0xff, det = (w0 >> 16) & 0xff;
    s.meta[road] = (uint32_t)ld | ((uint32_t)lc << 8) | ((uint32_t)ld << 16);
    s.wait[road] = __float_as_int(s.vs[road * CAP]);
    s.pd[road] = det << 16;
    s.leadx[road] = s.xs[road * CAP + ld];
    s.tailx[road] = (lc != ld) ? s.xs[road * CAP + lc] : INF;
  }
  __syncthreads();

  const IdmConst c = p.idm;
  int veh_local = 0, gen_local = 0;
  int t = 0;
  for (; t < p.K; t++) {
    // ---------------------------------------------------------------- phase A
    for (int g = warp; g < p.G; g += nwarps) {
      const int my_road = g * GROUP_ROADS + lane;  // per-road bookkeeping lane (lanes 0..7)
      const bool road_lane = lane < GROUP_ROADS && my_road < p.R;
      if (road_lane) {
        const uint32_t m = s.meta[my_road];
        const int ld = m & 0xff;
        int lc = (m >> 8) & 0xff;
        float *xr = s.xs + my_road * CAP, *vr = s.vs + my_road * CAP;
        const int ei = s.eidx[my_road];
        if (ei >= 0) {
          const int na = s.cnt[t * p.n_entry + ei];
          for (int k = 0; k < na; k++) {
            gen_local++;
            if (!ring_push(xr, vr, ld, lc, c.x_new, c.v_new, c)) {
              if (my_road < p.r) atomicAdd(&s.ovf[my_road % p.V], 1);
              atomicAdd(&s.misc[3], 1);
              atomicMin(&s.misc[0], t);
            }
          }
          s.meta[my_road] = (m & 0xffff00ffu) | ((uint32_t)lc << 8);
        }
        if (my_road < p.r) {  // update_lights, traffic_env.py:81-94
          const int dst = my_road % p.V;
          const bool ls_act = learn_switch && s.act[dst];
          const int ph_t = s.phase[dst] ^ (ls_act ? (t & 1) : 0);
          const int el_t = ls_act ? 0 : s.elapsed[dst] + t;
          const int road_phase = (my_road / p.V) < 2;
          float lx;
          if (road_phase == ph_t || el_t < YELLOW_TICKS) lx = p.length;
          else { const int nr = s.nexts[my_road]; lx = nr >= 0 ? __fadd_rn(s.tailx[nr], p.length) : INF; }
          s.leadx[my_road] = lx;
        }
        xr[0] = xr[CAP - 1]; vr[0] = vr[CAP - 1];  // mirror, traffic_env.py:203
      }
      __syncwarp();
      uint32_t accw = 0, accd = 0, accp = 0;
#pragma unroll 1
      for (int pass = GROUP_PASSES - 1; pass >= 0; --pass) {
        const int f = pass * 32 + lane;
        const int rl = f / CAP, slot = f - rl * CAP;
        const int road = g * GROUP_ROADS + rl;
        const uint32_t m = s.meta[road];
        const int ld = m & 0xff, lc = (m >> 8) & 0xff;
        const bool live = slot >= 1 && (ld < lc ? (slot > ld && slot <= lc) : (ld > lc && (slot > ld || slot <= lc)));
        float xn = 0.f, vn = 0.f;
        bool pw = false, pdet = false, pp = false;
        const int o = road * CAP + slot;
        if (live) {
          float x = s.xs[o], v = s.vs[o];
          const bool first = (slot == 1) ? (ld == CAP - 1) : (slot - 1 == ld);
          const float xl = first ? s.leadx[road] : s.xs[o - 1];
          const float vl = first ? 0.f : s.vs[o - 1];
          const float ll = first ? 0.f : c.len;
          idm_update(c, s.tabs, xl, vl, ll, x, v);
          xn = x; vn = v;
          // wrapped ring, low segment: the reference tests x, not v (traffic_env.py:210)
          const bool lowseg = ld > lc && slot <= lc;
          pw = (double)(lowseg ? xn : vn) < 0.2;
          pdet = (double)xn > p.det_thr;
          pp = xn > p.length;
        }
        const uint32_t bw = __ballot_sync(FULL, pw), bd = __ballot_sync(FULL, pdet), bp = __ballot_sync(FULL, pp);
        __syncwarp();  // every lane has read its leader before any lane overwrites a slot (Jacobi update)
        if (live) { s.xs[o] = xn; s.vs[o] = vn; }
        if (lane < GROUP_ROADS) {
          const int base = CAP * lane - 32 * pass;  // bit position of my road's slot 0 in this pass
          if (base < 32 && base > -CAP) {
            const uint32_t mask = (1u << CAP) - 1;
            accw |= (base >= 0 ? bw >> base : bw << -base) & mask;
            accd |= (base >= 0 ? bd >> base : bd << -base) & mask;
            accp |= (base >= 0 ? bp >> base : bp << -base) & mask;
          }
        }
      }
      __syncwarp();
      if (road_lane) {
        const uint32_t m = s.meta[my_road];
        const int ld = m & 0xff, lc = (m >> 8) & 0xff;
        const int n = ring_count(ld, lc);
        int newld = ld, npop = 0;
        if (n > 0) {
          veh_local += n;
          if (my_road < p.r) {
            s.wait[my_road] += __popc(accw);
            s.pd[my_road] = (s.pd[my_road] & 0xffff) | (__popc(accd) << 16);
          }
          const uint32_t q = (accp >> 1) & 0x7ffffu;   // bit i <-> slot i + 1
          const int sh = ring_wrap(ld + 1) - 1;
          const uint32_t rot = ((q >> sh) | (q << (RING - sh))) & 0x7ffffu;
          npop = __ffs(~rot) - 1;                       // leading run of cars past the end of the road
          if (npop > 0) {
            newld = (ld - 1 + npop) % RING + 1;
            const int nr = s.nexts[my_road];
            if (nr >= 0) {
              s.pd[my_road] += npop;                    // passed, traffic_env.py:127
              s.pdst[my_road % p.V] = 1;                // passed_dst, :128
              // Two or more pops while the upstream road has a higher index: its insert (which in the
              // reference runs after these pops were consumed) could reuse the slots the popped cars
              // still occupy.  Run this tick's transfers in strict road order instead.
              if (npop >= 2 && s.up[my_road] > my_road) s.misc[1] = t;
            }
          }
        }
        s.meta[my_road] = (uint32_t)newld | ((uint32_t)lc << 8) | ((uint32_t)ld << 16) | ((uint32_t)npop << 24);
      }
    }
    __syncthreads();
    // ---------------------------------------------------------------- phase C
    if (s.misc[1] == t) {
      if (tid == 0) {
        for (int e = 0; e < p.R; e++) { const int d = s.nexts[e]; if (d >= 0) transfer(p, s, e, d, t); }
        s.misc[5] += 1;
      }
      __syncthreads();
      for (int d = tid; d < p.R; d += blockDim.x) {
        const uint32_t md = s.meta[d];
        const int dld = md & 0xff, dlc = (md >> 8) & 0xff;
        s.tailx[d] = (dlc != dld) ? s.xs[d * CAP + dlc] : INF;
      }
    } else {
      for (int d = tid; d < p.R; d += blockDim.x) {
        const int u = s.up[d];
        if (u >= 0) transfer(p, s, u, d, t);
        const uint32_t md = s.meta[d];
        const int dld = md & 0xff, dlc = (md >> 8) & 0xff;
        s.tailx[d] = (dlc != dld) ? s.xs[d * CAP + dlc] : INF;
      }
    }
    __syncthreads();
    if (s.misc[0] <= t) { t++; break; }  // Repeater: `if done: break` (traffic_test.py:55)
  }
  const int ticks_run = t;  // >= 1
  const int last = ticks_run - 1;

  // ------------------------------------------------------------ epilogue
  // vehicle-update / generated-car counters
  for (int o = 16; o > 0; o >>= 1) {
    veh_local += __shfl_xor_sync(FULL, veh_local, o);
    gen_local += __shfl_xor_sync(FULL, gen_local, o);
  }
  if (lane == 0) { atomicAdd(&s.misc[2], veh_local); atomicAdd(&s.misc[4], gen_local); }

  const int obs_f_len = 2 * p.r + p.I, obs_i_len = 2 * p.r + 2 * p.I;
  // final light state after `ticks_run` ticks
  for (int i = tid; i < p.I; i += blockDim.x) {
    const bool ls_act = learn_switch && s.act[i];
    const int ph_f = s.phase[i] ^ (ls_act ? (last & 1) : 0);
    const int el_f = ls_act ? 0 : s.elapsed[i] + last;
    p.phase[(size_t)env * p.I + i] = (uint8_t)ph_f;
    p.elapsed[(size_t)env * p.I + i] = el_f;
    float rew = (float)(-10 * s.ovf[i]);  // OVERFLOW_PENALTY summed over the ticks: small integers, exact, +0 when none
    if (!p.raw && (p.flags & F_REMI)) {
      // remi, traffic_env.py:64-78, over the 4 approaches of intersection i in road order
      rew = 0.f;
      const bool pd = s.pdst[i] != 0;
      for (int dd = 0; dd < 4; dd++) {
        const int e = dd * p.V + i;
        const bool green = ((dd < 2) ? 1 : 0) != ph_f;
        const bool w = s.wait[e] > 0;
        if (w && !green && !pd) rew = __fsub_rn(rew, 0.5f);
        else if (pd && green && !w) rew = __fadd_rn(rew, 0.5f);
      }
    }
    p.reward[(size_t)env * p.I + i] = rew;
    if (p.raw) {
      p.obs_i[(size_t)env * obs_i_len + 2 * p.r + i] = ph_f;
      p.obs_i[(size_t)env * obs_i_len + 2 * p.r + p.I + i] = el_f;
    } else {
      // Repeater: obs[-I:] / 100 * (2 * phase - 1): int32 / int -> float64, cast to float32 on store
      p.obs_f[(size_t)env * obs_f_len + 2 * p.r + i] =
          __double2float_rn(__dmul_rn(__ddiv_rn((double)el_f, 100.0), (double)(2 * ph_f - 1)));
    }
    s.ovf[i] = __float_as_int(rew);  // reuse: reward for the return statistic below
  }
  __syncthreads();  // remi reads wait/pdst of all approaches before they are cleared
  const bool clear_remi = !p.raw && (p.flags & F_REMI);
  for (int i = tid; i < p.I; i += blockDim.x)
    p.passed_dst[(size_t)env * p.I + i] = clear_remi ? 0 : s.pdst[i];
  for (int e = tid; e < p.r; e += blockDim.x) {
    const int pd = s.pd[e];
    if (p.raw) {
      p.obs_i[(size_t)env * obs_i_len + e] = pd & 0xffff;
      p.obs_i[(size_t)env * obs_i_len + p.r + e] = pd >> 16;
    } else {
      p.obs_f[(size_t)env * obs_f_len + e] = (float)(pd & 0xffff);
      p.obs_f[(size_t)env * obs_f_len + p.r + e] = (float)(pd >> 16);
    }
  }
  // pack ring indices back into the row headers, restore the virtual leader's x, flush
  for (int road = tid; road < p.Rp; road += blockDim.x) {
    const uint32_t m = s.meta[road];
    const int ld = m & 0xff, lc = (m >> 8) & 0xff;
    s.xs[road * CAP + ld] = s.leadx[road];
    s.xs[road * CAP] = __uint_as_float(pack_meta(ld, lc, (s.pd[road] >> 16) & 0xff));
    s.vs[road * CAP] = __int_as_float((clear_remi || road >= p.r) ? 0 : s.wait[road]);
  }
  __syncthreads();
  {
    float4 *gx = reinterpret_cast<float4 *>(p.x + (size_t)env * p.Rp * CAP);
    float4 *gv = reinterpret_cast<float4 *>(p.v + (size_t)env * p.Rp * CAP);
    const float4 *sx = reinterpret_cast<const float4 *>(s.xs), *sv = reinterpret_cast<const float4 *>(s.vs);
    const int n4 = p.Rp * (CAP / 4);
    for (int i = tid; i < n4; i += blockDim.x) { gx[i] = sx[i]; gv[i] = sv[i]; }
  }
  if (tid == 0) {
    const bool overflowed = s.misc[0] != 0x7fffffff;
    p.done[env] = overflowed ? 1 : 0;
    es->steps = es->steps + (float)ticks_run;
    es->sched_cursor += ticks_run;
    if (p.arrival_mode == ARR_PHILOX) { es->ph_draw = s.snap[2 * ticks_run]; es->ph_skip = s.snap[2 * ticks_run + 1]; }
    if (!p.raw) {
      double mean = 0.0;
      for (int i = 0; i < p.I; i++) mean += (double)__int_as_float(s.ovf[i]);
      mean /= (double)p.I;
      es->ep_ret += mean; es->ep_disc += es->ep_mult * mean; es->ep_mult *= (double)p.gamma;
      es->ep_step += 1;
      es->done = overflowed ? 1 : 0;
      atomicAdd(&p.stats->actor_steps, 1ull);
    }
    atomicAdd(&p.stats->ticks, (unsigned long long)ticks_run);
    atomicAdd(&p.stats->vehicle_updates, (unsigned long long)s.misc[2]);
    if (s.misc[3]) atomicAdd(&p.stats->overflows, (unsigned long long)s.misc[3]);
    // cars generated in ticks that were not run (break on overflow) are not counted
    atomicAdd(&p.stats->cars_generated, (unsigned long long)s.misc[4]);
    if (s.misc[5]) atomicAdd(&p.stats->seq_fallback_ticks, (unsigned long long)s.misc[5]);
  }
}

// TrafficEnv._reset (traffic_env.py:259-272) on the HBM state.  mask == nullptr: every env;
// use_done != 0: envs whose last actor step ended the episode (TE_AUTO_RESET).
__global__ void te_reset_kernel(StepParams p, const uint8_t *mask, const uint8_t *init_phase, int use_done) {
  const int env = blockIdx.x;
  EnvScalars *es = p.env + env;
  bool doit = mask ? mask[env] != 0 : true;
  if (use_done) doit = es->done || (p.episode_len > 0 && es->ep_step >= p.episode_len);
  if (!doit) return;
  const int tid = threadIdx.x;
  float *x = p.x + (size_t)env * p.Rp * CAP, *v = p.v + (size_t)env * p.Rp * CAP;
  for (int road = tid; road < p.Rp; road += blockDim.x) {
    const int det = (__float_as_uint(x[road * CAP]) >> 16) & 0xff;  // `detected` survives a reset
    x[road * CAP] = __uint_as_float(pack_meta(1, 1, det));
    v[road * CAP] = __int_as_float(0);
    x[road * CAP + 1] = __int_as_float(0x7f800000);
    v[road * CAP + 1] = 0.f;
  }
  for (int i = tid; i < p.I; i += blockDim.x) {
    int ph;
    if (init_phase) ph = init_phase[(size_t)env * p.I + i] != 0;
    else {
      uint32_t o[4];
      philox4x32_10(es->reset_count, (uint32_t)(i >> 7), 1u, 0x5e5e7u, p.seed, (uint32_t)(p.env_id_base + env), o);
      ph = (o[(i >> 5) & 3] >> (i & 31)) & 1;
    }
    p.phase[(size_t)env * p.I + i] = (uint8_t)ph;
    p.elapsed[(size_t)env * p.I + i] = 0;
    p.passed_dst[(size_t)env * p.I + i] = 0;
  }
  __syncthreads();
  if (tid == 0) {
    if (es->ep_step > 0) {
      atomicAdd(&p.stats->episodes, 1ull);
      atomicAdd(&p.stats->return_sum, es->ep_ret);
      atomicAdd(&p.stats->disc_return_sum, es->ep_disc);
    }
    es->steps = 0.f; es->ep_step = 0; es->done = 0;
    es->ep_ret = 0.0; es->ep_disc = 0.0; es->ep_mult = 1.0;
    es->reset_count += 1;
  }
}

// remi_reward as a stand-alone call (traffic_env.py:384-387) for the single-tick API.
__global__ void te_remi_kernel(StepParams p, float *reward) {
  const int env = blockIdx.x;
  float *v = p.v + (size_t)env * p.Rp * CAP;
  for (int i = threadIdx.x; i < p.I; i += blockDim.x) {
    const int ph = p.phase[(size_t)env * p.I + i];
    const bool pd = p.passed_dst[(size_t)env * p.I + i] != 0;
    float rew = 0.f;
    for (int dd = 0; dd < 4; dd++) {
      const int e = dd * p.V + i;
      const bool green = ((dd < 2) ? 1 : 0) != ph;
      const bool w = __float_as_int(v[e * CAP]) > 0;
      if (w && !green && !pd) rew = __fsub_rn(rew, 0.5f);
      else if (pd && green && !w) rew = __fadd_rn(rew, 0.5f);
    }
    reward[(size_t)env * p.I + i] = rew;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < p.I; i += blockDim.x) p.passed_dst[(size_t)env * p.I + i] = 0;
  for (int e = threadIdx.x; e < p.r; e += blockDim.x) v[e * CAP] = __int_as_float(0);
}

// cars_on_roads (traffic_env.py:214-218): out[E][R]
__global__ void te_cars_kernel(StepParams p, int *out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)p.num_envs * p.R) return;
  const int env = (int)(i / p.R), road = (int)(i % p.R);
  const uint32_t w0 = __float_as_uint(p.x[((size_t)env * p.Rp + road) * CAP]);
  out[i] = ring_count(w0 & 0xff, (w0 >> 8) & 0xff);
}

// greedy controller (algorithms/greedy.py:14-16): cars_on_roads()[row, col, :] . [1,1,-1,-1] < 0
__global__ void te_greedy_kernel(StepParams p, uint8_t *actions) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)p.num_envs * p.I) return;
  const int env = (int)(i / p.I), it = (int)(i % p.I);
  int cnt[4];
  for (int dd = 0; dd < 4; dd++) {
    const uint32_t w0 = __float_as_uint(p.x[((size_t)env * p.Rp + dd * p.V + it) * CAP]);
    cnt[dd] = ring_count(w0 & 0xff, (w0 >> 8) & 0xff);
  }
  actions[i] = (cnt[0] + cnt[1] - cnt[2] - cnt[3]) < 0 ? 1 : 0;
}

// ---- test hooks
__global__ void te_test_powf_kernel(const float *x, float y, float *out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = powf_glibc(x[i], y, &g_powf_tables);
}
__global__ void te_test_idm_kernel(IdmConst c, const float *xl, const float *vl, const float *ll, const float *x,
                                   const float *v, float *xo, float *vo, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float xx = x[i], vv = v[i];
    idm_update(c, &g_powf_tables, xl[i], vl[i], ll[i], xx, vv);
    xo[i] = xx; vo[i] = vv;
  }
}
__global__ void te_test_philox_kernel(const uint32_t *ctr, const uint32_t *key, uint32_t *out) {
  uint32_t o[4];
  philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], o);
  for (int i = 0; i < 4; i++) out[i] = o[i];
}

}  // namespace te
