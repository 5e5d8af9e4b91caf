#!/usr/bin/env python
"""ncu raw page (`ncu -i X.ncu-rep --page raw --csv`) -> a small JSON summary for profiles/ (one object per profiled
launch: the counters the roofline discussion in DESIGN.md uses), optionally an entry of profiles/r2_gram_traffic.json
(DRAM bytes per launch tied to the hash of the kernel source the capture was taken from, read by bench.py).

usage: tools/ncu_summary.py raw.csv out.json [--traffic i8|e2m1|bf16 --source "how it was captured"]"""
import argparse
import csv
import hashlib
import json
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
KEEP = [
    "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second",
    "smsp__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__grid_size", "launch__block_size",
    "launch__cluster_size", "launch__cluster_max_active", "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_membar_per_warp_active.pct",
]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw")
    ap.add_argument("out")
    ap.add_argument("--traffic", default="")
    ap.add_argument("--source", default="")
    args = ap.parse_args()
    rows = list(csv.reader(open(args.raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        if len(vals) != len(hdr):
            continue
        rec = {}
        for h, u, v in zip(hdr, units, vals):
            if h in KEEP:
                rec[h] = (v + (" " + u if u else "")).strip()
        out.append(rec)
    Path(args.out).write_text(json.dumps(out, indent=1) + "\n")
    print(json.dumps(out[0], indent=1)[:1200] if out else "no rows")
    if args.traffic and out:
        def to_bytes(s):
            m = re.match(r"([0-9.eE+-]+)\s*(\w+)", s)
            return float(m.group(1)) * SCALE[m.group(2)]
        tot = to_bytes(out[0]["dram__bytes_read.sum"]) + to_bytes(out[0]["dram__bytes_write.sum"])
        tp = ROOT / "profiles" / "r2_gram_traffic.json"
        tj = json.loads(tp.read_text()) if tp.exists() else {}
        import sys
        sys.path.insert(0, str(ROOT))
        from spark_examples_b200 import native
        tj[args.traffic] = {"dram_bytes_per_launch": int(tot), "kernel_code_sha256_16": native.gramSourceFingerprint(),
                            "source": args.source or f"{args.out}: dram__bytes_read.sum + dram__bytes_write.sum of one warmed-up launch"}
        tp.write_text(json.dumps(tj, indent=1) + "\n")


main()
