#!/usr/bin/env python3
"""Selected raw metrics of every kernel in an .ncu-rep (ncu -i X --page raw --csv), one kernel per block.
usage: ncu_raw.py X.ncu-rep [regex-on-metric-names]   (prints a tidy `metric,value` CSV per launch)"""
import csv
import io
import re
import subprocess
import sys

DEFAULT = (r"gpu__time_duration\.sum|launch__registers_per_thread|launch__occupancy|launch__grid_size|launch__block_size|"
           r"sm__warps_active\.avg\.pct|sm__pipe_fp64_cycles_active|sm__inst_executed_pipe_fp64|smsp__inst_executed\.sum$|"
           r"smsp__issue_active\.avg\.pct|smsp__thread_inst_executed_per_inst_executed|dram__bytes_(read|write)\.sum$|"
           r"dram__throughput\.avg\.pct|local_(load|store)|smsp__average_warps?_issue_stalled.*_per_issue_active|"
           r"sm__throughput\.avg\.pct|sm__cycles_elapsed\.max|lts__t_sector_hit_rate|smsp__inst_executed_op_.*fp64|"
           r"sm__sass_thread_inst_executed_op_d(fma|add|mul)_pred_on\.sum$|smsp__cycles_active\.avg$")


def main():
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else DEFAULT)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("# launch %s  %s  grid %s block %s" % (d["ID"], d["Kernel Name"], d["Grid Size"], d["Block Size"]))
        print("metric,unit,value")
        for k, u in zip(hdr, units):
            name = k.split(".", 2)[-1] if k.count(".") >= 2 and k.split(".")[1].startswith("Triage") else k
            if pat.search(k):
                print("%s,%s,%s" % (k, u, d[k]))
        print()


if __name__ == "__main__":
    main()
