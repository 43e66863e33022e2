"""TEST INFRASTRUCTURE ONLY - loads the UNMODIFIED reference simulator
(/root/reference, read-only, only present in the build container) so that the
C restatement in oracle/traffic_oracle.c can be pinned against it and golden
vectors can be generated (oracle/gen_golden.py).

Nothing in the product path (traffic_env_b200/) imports this module.  Nothing in
the ``-m gpu`` tests, ``smoke()`` or ``bench.py`` imports it either: the
reference cannot travel to the GPU box.

What is needed to import the reference unchanged:
  * an old-API ``gym`` (tests/support/gym_compat),
  * ``np.bool8`` (removed in numpy 2; used at traffic_env.py:380),
  * ``alg_flags`` imported before stepping (registers FLAGS.mode, read at
    traffic_env.py:240),
  * a writable NUMBA_CACHE_DIR (the jitted functions use cache=True and the
    reference tree is read-only),
  * inert ``tensorflow`` / ``matplotlib.pyplot`` modules for the baseline
    agents (algorithms/fixed.py:1, util.py:5).
"""
import importlib
import os
import sys
import types

REFERENCE_DIR = os.environ.get("TRAFFIC_ENV_REFERENCE", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
_SUPPORT = os.path.join(os.path.dirname(_HERE), "tests", "support")


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "gym_traffic", "envs", "traffic_env.py"))


_loaded = None


def load():
    """Import the reference and return a namespace with its modules."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_DIR)
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/traffic_env_numba_cache")
    os.environ.setdefault("NUMBA_NUM_THREADS", "1")
    import numpy as np
    if not hasattr(np, "bool8"):
        np.bool8 = np.bool_
    gym_dir = os.path.join(_SUPPORT, "gym_compat")
    for p in (REFERENCE_DIR, gym_dir):
        if p not in sys.path:
            sys.path.insert(0, p)
    # The product's drop-in gym_traffic must not shadow the reference here.
    for name in list(sys.modules):
        if name == "gym_traffic" or name.startswith("gym_traffic."):
            mod = sys.modules[name]
            f = getattr(mod, "__file__", "") or ""
            if not f.startswith(REFERENCE_DIR):
                del sys.modules[name]
    for stub in ("tensorflow", "matplotlib", "matplotlib.pyplot"):
        if stub not in sys.modules:
            try:
                importlib.import_module(stub)
            except Exception:
                sys.modules[stub] = types.ModuleType(stub)
    if "matplotlib" in sys.modules and "matplotlib.pyplot" in sys.modules:
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    import gym  # noqa: F401  (the compat one unless a real old gym exists)
    import args
    import alg_flags  # noqa: F401  registers FLAGS.mode et al.
    import gym_traffic
    from gym_traffic.envs import traffic_env, roadgraph
    from gym_traffic.spaces import gspace
    assert traffic_env.__file__.startswith(REFERENCE_DIR), traffic_env.__file__
    ns = types.SimpleNamespace(gym=gym, args=args, FLAGS=args.FLAGS, alg_flags=alg_flags,
                               gym_traffic=gym_traffic, traffic_env=traffic_env,
                               roadgraph=roadgraph, gspace=gspace, np=np)
    _loaded = ns
    return ns


class ReplayArrivals(object):
    """Injects a recorded arrival schedule into the unmodified reference env.

    TrafficEnv.add_new_cars (traffic_env.py:274-283) only touches
    ``self.rand_car`` (a generator yielding car rows / None) and
    ``self.rand.choice(entrypoints)``; replacing those two attributes replays
    a schedule without editing the reference.
    ``schedule[t]`` is the ordered list of entry-road ids for tick t.
    """

    def __init__(self, ref, schedule):
        self.ref = ref
        self.schedule = schedule
        self._pending = []

    def install(self, env):
        env.rand_car = self._cars()
        env.rand = self

    def _cars(self):
        arch = self.ref.traffic_env.archetypes
        t = 0
        while True:
            roads = self.schedule[t] if t < len(self.schedule) else ()
            for rd in roads:
                self._pending.append(int(rd))
                yield arch[0]
            yield None
            t += 1

    def choice(self, entrypoints):
        return self._pending.pop(0)


def record_schedule(ref, m, n, ticks, seed, local_cars_per_sec=None, entry="all"):
    """Run the reference's own generators (traffic_env.py:160-176, 274-283) with
    RandomState(seed) and record tick -> ordered entry-road list, without
    stepping the physics."""
    np = ref.np
    F = ref.FLAGS
    if local_cars_per_sec is not None:
        F.local_cars_per_sec = local_cars_per_sec
    F.entry = entry
    g = ref.roadgraph.GridRoad(m, n, 250)
    spec = 0 if entry == "all" else 0b1110
    g.generate_entrypoints(spec)
    F.cars_per_sec = F.local_cars_per_sec * g.m * ref.traffic_env.inv_popcount(spec)
    rand = np.random.RandomState(seed)
    gen = ref.traffic_env.poisson(rand) if F.poisson else ref.traffic_env.regular(rand)
    sched = []
    for _ in range(ticks):
        roads = []
        car = next(gen)
        while car is not None:
            roads.append(int(rand.choice(g.entrypoints)))
            car = next(gen)
        sched.append(roads)
    return sched


def make_env(ref, m=3, n=3, length=250, seed=0):
    """gym.make('traffic-v0') + set_graph/seed_generator/reset_entrypoints as in
    traffic_test.py:78-83."""
    env = ref.gym.make("traffic-v0")
    env.set_graph(ref.roadgraph.GridRoad(m, n, length))
    env.seed_generator(seed)
    env.reset_entrypoints()
    return env


def live_state(env):
    """Ring-order walk of the live slots of every road (SURVEY.md 8a quirks:
    only slots in (leading, lastcar] are meaningful)."""
    np = __import__("numpy")
    xs, vs = [], []
    for e in range(env.graph.roads):
        s = int(env.leading[e])
        while s != int(env.lastcar[e]):
            s = 1 if s + 1 >= 20 else s + 1
            xs.append(env.state[e, 0, s])
            vs.append(env.state[e, 1, s])
    return np.asarray(xs, dtype=np.float32), np.asarray(vs, dtype=np.float32)
