#!/usr/bin/env python
"""Condense one kernel of an .ncu-rep into the small `metric,value,unit` CSV kept under profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r01_ncu_x.csv [row]

Keeps launch geometry, duration, DRAM traffic, pipe utilisation, issue/occupancy figures and every warp-stall ratio;
`ncu -i <rep> --page raw --csv` is the source (run here, no GPU needed).
"""
import csv
import io
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "launch__", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput",
        "sm__inst_executed_pipe_", "sm__pipe_", "smsp__inst_executed.sum", "smsp__issue_active", "smsp__warps_active",
        "smsp__warps_eligible", "sm__warps_active", "smsp__average_warps_issue_stalled", "sm__cycles_elapsed.avg",
        "smsp__thread_inst_executed_per_inst_executed", "l1tex__data_bank_conflicts", "lts__t_sector_hit_rate")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    row = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2 + row]
    d = dict(zip(hdr, zip(vals, units)))
    with open(out, "w") as f:
        f.write(f"Kernel Name,{d['Kernel Name'][0]},\nBlock Size,{d['Block Size'][0]},\nGrid Size,{d['Grid Size'][0]},\n")
        for k in sorted(d):
            if any(x in k for x in (".max.", ".min.", ".sum.pct_of_peak", ".sum.per_cycle", "_elapsed")) and not k.startswith("dram__"):
                continue          # one representative per counter: .avg (or the plain .sum for totals)
            if k.startswith(KEEP) and d[k][0] not in ("", "n/a"):
                f.write(f"{k},{d[k][0].replace(',', '')},{d[k][1]}\n")
    print(f"wrote {out}")


if __name__ == "__main__":
    main()
