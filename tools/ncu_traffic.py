"""Developer tool: record the measured DRAM traffic of the step kernel for bench.py's `roofline.traffic`.

  python tools/ncu_traffic.py <workload> <capture.ncu-rep> [note]

Reads dram__bytes_read.sum + dram__bytes_write.sum of the (single) profiled te_step_kernel launch from an
`ncu --set full` capture and writes profiles/traffic_bytes.json[workload] together with the sha of the kernel sources
the capture was taken from; bench.py refuses the number when the sources have changed since (measured_traffic)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_sha  # noqa: E402


def main():
    wl, rep = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else os.path.basename(rep)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}

    def get(name):
        v, u = float(vals[col[name]].replace(",", "")), units[col[name]]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    total = get("dram__bytes_read.sum") + get("dram__bytes_write.sum")
    path = os.path.join(ROOT, "profiles", "traffic_bytes.json")
    try:
        d = json.load(open(path))
    except Exception:
        d = {}
    d[wl] = {"dram_bytes_per_launch": int(total), "kernel_source_sha": kernel_source_sha(), "capture": note,
             "kernel": vals[col["Kernel Name"]] if "Kernel Name" in col else "te_step_kernel",
             "duration_ms_under_ncu": float(vals[col["gpu__time_duration.sum"]].replace(",", "")) *
             {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "msecond": 1.0, "ms": 1.0, "nsecond": 1e-6, "second": 1e3}.get(units[col["gpu__time_duration.sum"]], 1.0)}
    json.dump(d, open(path, "w"), indent=1)
    print(wl, d[wl])


if __name__ == "__main__":
    main()
