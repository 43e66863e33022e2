"""Developer tool: build kernel variants (extra -D flags) side by side and measure them in ONE gpurun call.

  python tools/ab_variants.py build  name1:-DFOO=1,-DBAR name2: ...     # here: variants/libtraffic_b200_<name>.so
  python tools/ab_variants.py run [--workloads a,b] [--steps N] name1 name2 ...   # on the GPU box

`run` executes the smoke check (bit-exact vs the oracle) and bench.py --no-e2e --no-cpu-baseline --no-secondary for every
variant and workload, several alternating repetitions, and prints one table.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VDIR = os.path.join(ROOT, "variants")
sys.path.insert(0, ROOT)


def so_of(name):
    return os.path.join(VDIR, "libtraffic_b200_%s.so" % name.split("@")[0])


def env_of(name):
    """`build@VAR=val@VAR2=val2`: the library build plus run-time study knobs (environment variables)."""
    env = dict(os.environ, TRAFFIC_B200_SO=so_of(name))
    for kv in name.split("@")[1:]:
        k, _, v = kv.partition("=")
        env[k] = v
    return env


def main():
    if sys.argv[1] == "build":
        from traffic_env_b200 import build as b
        os.makedirs(VDIR, exist_ok=True)
        for spec in sys.argv[2:]:
            name, _, defs = spec.partition(":")
            defines = [d[2:] if d.startswith("-D") else d for d in defs.split(",") if d]
            print("building", name, defines, flush=True)
            b.build(out=so_of(name), defines=defines)
        return
    args = sys.argv[2:]
    workloads, steps, reps, names, extra = ["grid10x10_L500_greedy", "grid3x3_L250_greedy"], 21, 2, [], []
    while args:
        a = args.pop(0)
        if a == "--workloads":
            workloads = args.pop(0).split(",")
        elif a == "--steps":
            steps = int(args.pop(0))
        elif a == "--reps":
            reps = int(args.pop(0))
        elif a.startswith("--bench-args"):
            extra = (a.split("=", 1)[1] if "=" in a else args.pop(0)).split()
        else:
            names.append(a)
    res = {}
    for n in names:
        env = env_of(n)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "__graft_entry__.py"), "--smoke-only"], env=env,
                           capture_output=True, text=True)
        res[(n, "smoke")] = "ok" if r.returncode == 0 and "smoke ok" in r.stdout else "FAIL " + (r.stdout + r.stderr)[-300:]
        print(n, "smoke:", res[(n, "smoke")], flush=True)
    for rep in range(reps):
        for wl in workloads:
            for n in names:
                env = env_of(n)
                r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", wl, "--steps", str(steps),
                                    "--warmup", "6", "--no-e2e", "--no-cpu-baseline", "--no-secondary"] + extra, env=env,
                                   capture_output=True, text=True)
                try:
                    d = json.loads(r.stdout.strip().splitlines()[-1])
                    val = (d["value"], d["roofline"]["kernel_ms"], d["clocks"]["sm_mhz"])
                except Exception:
                    val = ("ERR", (r.stdout + r.stderr)[-400:], None)
                res.setdefault((n, wl), []).append(val)
                print(rep, wl, n, val, flush=True)
    print("\n%-28s" % "variant" + "".join("%26s" % w[:24] for w in workloads))
    for n in names:
        row = "%-28s" % n
        for wl in workloads:
            vals = [v[0] for v in res.get((n, wl), []) if v[0] != "ERR"]
            row += "%26s" % ("%.4g (%s)" % (max(vals), ",".join("%.3g" % v for v in vals)) if vals else "ERR")
        print(row + "   smoke " + res[(n, "smoke")][:8])


if __name__ == "__main__":
    main()
