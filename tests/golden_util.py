"""Helpers shared by the golden-vector generator and the parity tests."""
import hashlib

import numpy as np

CAP = 20


def tick_digest(leading, lastcar, obs, waiting, passed_dst, rewards, done, live_x, live_v):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(leading, dtype=np.int32).tobytes())
    h.update(np.ascontiguousarray(lastcar, dtype=np.int32).tobytes())
    h.update(np.ascontiguousarray(obs, dtype=np.int32).tobytes())
    h.update(np.ascontiguousarray(waiting, dtype=np.int32).tobytes())
    h.update(np.ascontiguousarray(np.asarray(passed_dst).astype(bool), dtype=np.uint8).tobytes())
    h.update(np.ascontiguousarray(rewards, dtype=np.float32).tobytes())
    h.update(np.uint8(bool(done)).tobytes())
    h.update(np.ascontiguousarray(live_x, dtype=np.float32).tobytes())
    h.update(np.ascontiguousarray(live_v, dtype=np.float32).tobytes())
    return np.frombuffer(h.digest()[:8], dtype=np.uint64)[0]


def pack_schedule(sched):
    """list (per tick) of road lists -> CSR (offsets int32[T+1], roads int16[])."""
    off = np.zeros(len(sched) + 1, dtype=np.int32)
    for t, s in enumerate(sched):
        off[t + 1] = off[t] + len(s)
    roads = np.asarray([rd for s in sched for rd in s], dtype=np.int16)
    return off, roads


def unpack_schedule(off, roads):
    return [roads[off[t]:off[t + 1]].astype(np.int32) for t in range(len(off) - 1)]


def live_walk(leading, lastcar, x, v):
    """Ring-order walk over (leading, lastcar] for arrays x, v of shape [R, 20]."""
    xs, vs = [], []
    for e in range(len(leading)):
        s = int(leading[e])
        while s != int(lastcar[e]):
            s = 1 if s + 1 >= CAP else s + 1
            xs.append(x[e, s])
            vs.append(v[e, s])
    return np.asarray(xs, dtype=np.float32), np.asarray(vs, dtype=np.float32)
