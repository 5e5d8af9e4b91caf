#!/usr/bin/env python
"""Time the Gram kernel alone over (cta_group, kb_window, V) settings on one GPU; prints one JSON line per setting."""
import json, os, sys, itertools
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from spark_examples_b200 import native

def main():
    n = int(os.environ.get("SWEEP_N", "2504"))
    vs = [int(v) for v in os.environ.get("SWEEP_V", "1000000").split(",")]
    cgs = [int(v) for v in os.environ.get("SWEEP_CG", "2,1").split(",")]
    kbws = [int(v) for v in os.environ.get("SWEEP_KBW", "0,18,37,74,148,296").split(",")]
    dt = os.environ.get("SWEEP_DTYPE", "i8")
    eb = 1 if dt == "i8" else 2
    reps = int(os.environ.get("SWEEP_REPS", "10"))
    vmax = max(vs)
    ld = ((vmax + 127) // 128) * 128
    X = torch.empty((n, ld), dtype=torch.int8 if eb == 1 else torch.bfloat16, device="cuda")
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    stream = ts.cuda_stream
    with native.NativePca(n, dtype=native.DTYPE_I8 if eb == 1 else native.DTYPE_BF16, stream=stream, max_multiplicity=1) as g:
        g.synthDenseDevice(20240901, 0, vmax, 0, X.data_ptr(), ld)
    torch.cuda.synchronize()
    ref = None
    for v, cg, kbw in itertools.product(vs, cgs, kbws):
        os.environ["VPCA_CTA_GROUP"] = str(cg)
        if kbw > 0: os.environ["VPCA_KB_WINDOW"] = str(kbw)
        else: os.environ.pop("VPCA_KB_WINDOW", None)
        with native.NativePca(n, dtype=native.DTYPE_I8 if eb == 1 else native.DTYPE_BF16, stream=stream, max_multiplicity=1) as nat:
            ts = []
            for r in range(reps + 2):
                nat.reset()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); nat.accumulateDenseDevice(X.data_ptr(), v, ld); b.record(); b.synchronize()
                if r >= 2: ts.append(a.elapsed_time(b))
            nat.finalizeGram()
            S = torch.from_numpy(nat.getGram())
            chk = int(S.to(torch.int64).sum().item())
            ts.sort()
            ms = ts[len(ts) // 2]
            ops = n * (n + 1) * v
            print(json.dumps({"n": n, "v": v, "cg": cg, "kbw": kbw, "ms_med": round(ms, 4), "ms_min": round(ts[0], 4),
                              "tops_syrk": round(ops / ms / 1e9, 1), "cells_per_s": round(n * v / ms * 1e3 / 1e9, 2),
                              "checksum": chk}), flush=True)
main()
