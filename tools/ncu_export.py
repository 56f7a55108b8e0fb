#!/usr/bin/env python3
"""Turn an .ncu-rep (scratch, gpurun_out/) into tracked evidence under profiles/:
   * profiles/<name>.raw.csv      the selected raw metrics of every captured launch (tools/ncu_raw.py format)
   * profiles/<name>.stalls.txt   opcode mix and stall-reason shares from the source page (tools/ncu_opmix.py)
   * profiles/ncu_traffic.json    key -> {dram_bytes_per_launch, units, src_sha, capture}: what bench.py quotes as
                                  roofline.traffic, valid only while the kernel sources hash to src_sha
usage: ncu_export.py report.ncu-rep name [traffic_key units [units_for_opmix]]"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench_large import source_sha  # noqa: E402


def main():
    rep, name = sys.argv[1], sys.argv[2]
    key = sys.argv[3] if len(sys.argv) > 3 else None
    units = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    raw = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_raw.py"), rep], capture_output=True, text=True).stdout
    open(os.path.join(ROOT, "profiles", name + ".raw.csv"), "w").write(raw)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    tmp = "/tmp/_ncu_src.csv"
    open(tmp, "w").write(src)
    op_units = sys.argv[5] if len(sys.argv) > 5 else (str(units) if units else "1")
    mix = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_opmix.py"), tmp, op_units], capture_output=True, text=True).stdout
    open(os.path.join(ROOT, "profiles", name + ".stalls.txt"), "w").write(
        "# %s: opcode mix per unit (units = %s) and stall-reason shares, from `ncu --page source`\n%s" % (os.path.basename(rep), op_units, mix))
    if key:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, unit_row, first = rows[0], rows[1], rows[2]
        d, u = dict(zip(hdr, first)), dict(zip(hdr, unit_row))
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = sum(float(d[k]) * scale[u[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        try:
            allrec = json.load(open(path))
        except Exception:  # noqa: BLE001
            allrec = {}
        allrec[key] = {"dram_bytes_per_launch": tot, "units": units, "src_sha": source_sha(key),
                       "capture": "profiles/%s.raw.csv (%s, launch %s: %s)" % (name, os.path.basename(rep), d["ID"], d["Kernel Name"][:60]),
                       "time_ms": float(d["gpu__time_duration.sum"]) * {"ms": 1.0, "us": 1e-3, "s": 1e3}.get(u["gpu__time_duration.sum"], 1.0)}
        json.dump(allrec, open(path, "w"), indent=1, sort_keys=True)
        print("traffic[%s] = %.4g bytes per launch" % (key, tot))
    print("wrote profiles/%s.raw.csv, .stalls.txt" % name)


if __name__ == "__main__":
    main()
