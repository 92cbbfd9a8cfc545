#!/usr/bin/env python
"""Fold an `ncu --set full` capture into profiles/ncu_current.json, the file bench.py reads its `roofline.traffic` /
`roofline.ncu` blocks from (no counter in bench.py is a literal).

usage: python profiles/ncu_to_json.py <rep.ncu-rep> <key> <n_rows> [<kernel-name-substring>] [<note>]
   e.g. python profiles/ncu_to_json.py gpurun_out/r2_mc.ncu-rep "mlp_tc_kernel<MC>" 1000000 mlp_tc_kernel

dram_bytes = dram__bytes_read.sum + dram__bytes_write.sum of ONE launch; bench.py scales it by n / n_rows.
"""
import csv
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "ncu_current.json")
KEEP = {"gpu__time_duration.sum": "duration", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pct",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pct",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
        "smsp__inst_executed.sum": "warp_instructions", "launch__registers_per_thread": "registers",
        "launch__grid_size": "grid", "launch__block_size": "block"}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    rep, key, n_rows = sys.argv[1], sys.argv[2], int(sys.argv[3])
    match = sys.argv[4] if len(sys.argv) > 4 else ""
    note = sys.argv[5] if len(sys.argv) > 5 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    pick = [r for r in rows[2:] if match in r[hdr.index("Kernel Name")]]
    if not pick:
        raise SystemExit(f"no kernel matching {match!r} in {rep}")
    row = pick[-1]
    val = lambda name: float(row[hdr.index(name)].replace(",", ""))
    unit = lambda name: units[hdr.index(name)]
    rec = {"source": f"profiles/{os.path.basename(rep).replace(chr(46) + chr(110) + chr(99) + chr(117) + chr(45) + chr(114) + chr(101) + chr(112), chr(46) + chr(115) + chr(117) + chr(109) + chr(109) + chr(97) + chr(114) + chr(121) + chr(46) + chr(116) + chr(120) + chr(116))} (ncu --set full --clock-control none; the .ncu-rep itself is not committed)", "kernel_name": row[hdr.index('Kernel Name')][:120],
           "n": n_rows, "dram_bytes": val("dram__bytes_read.sum") * SCALE[unit("dram__bytes_read.sum")]
           + val("dram__bytes_write.sum") * SCALE[unit("dram__bytes_write.sum")]}
    for name, short in KEEP.items():
        if name in hdr:
            rec[short] = val(name)
            if short == "duration":
                rec["duration_unit"] = unit(name)
    if note:
        rec["note"] = note
    cur = {}
    if os.path.exists(OUT):
        with open(OUT) as f:
            cur = json.load(f)
    cur[key] = rec
    with open(OUT, "w") as f:
        json.dump(cur, f, indent=1, sort_keys=True)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()
