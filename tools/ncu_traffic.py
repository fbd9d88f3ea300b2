"""profiles/r02_traffic.json from ncu reports: DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the
kernels bench.py reports a roofline for, tied to the build they were captured from (bench.source_id()).

    python tools/ncu_traffic.py "<command the reports were captured with>" name=report.ncu-rep[:launch_index] ...

e.g.  python tools/ncu_traffic.py "ncu --set full ... bench.py --steps 1 --warmup 1 --no-extra --no-cpu-baseline" \
          ballistic@config3=gpurun_out/r02_ballistic.ncu-rep contact@config3=gpurun_out/r02_contact.ncu-rep
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    hdr, units = r[0], r[1]
    return [dict(zip(hdr, x)) for x in r[2:]], dict(zip(hdr, units))


def to_bytes(v, unit):
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v) * f[unit]


def main():
    command = sys.argv[1]
    kernels, detail = {}, {}
    for arg in sys.argv[2:]:
        name, rep = arg.split("=")
        idx = 0
        if ":" in rep:
            rep, idx = rep.rsplit(":", 1)
            idx = int(idx)
        rs, units = rows(rep)
        d = rs[idx]
        rd = to_bytes(d["dram__bytes_read.sum"], units["dram__bytes_read.sum"])
        wr = to_bytes(d["dram__bytes_write.sum"], units["dram__bytes_write.sum"])
        kernels[name] = rd + wr
        detail[name] = {"kernel": d["Kernel Name"], "dram_read_bytes": rd, "dram_write_bytes": wr,
                        "duration_us_under_ncu": float(d["gpu__time_duration.sum"]) * {"us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}[units["gpu__time_duration.sum"]],
                        "report": os.path.basename(rep), "launch_index": idx}
    out = {"source_id": bench.source_id(), "command": command, "kernels": kernels, "detail": detail}
    with open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
