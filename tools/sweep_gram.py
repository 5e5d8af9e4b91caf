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
    kbws = [int(v) for v in os.environ.get("SWEEP_KBW", "0,37,74,148").split(",")]
    leads = [int(v) for v in os.environ.get("SWEEP_LEAD", "0").split(",")]
    panels = [int(v) for v in os.environ.get("SWEEP_PANEL", "0").split(",")]
    os.environ["VPCA_GRAM_PROF"] = "1"
    dt = os.environ.get("SWEEP_DTYPE", "i8")
    eb = {"i8": 1, "bf16": 2, "e2m1": 0.5}[dt]
    ndt = {"i8": native.DTYPE_I8, "bf16": native.DTYPE_BF16, "e2m1": native.DTYPE_E2M1}[dt]
    reps = int(os.environ.get("SWEEP_REPS", "10"))
    vmax = max(vs)
    ld = ((vmax + 127) // 128) * 128
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    stream = ts.cuda_stream
    X = None
    cur_panel = None
    for panel, v, cg, kbw, lead in itertools.product(panels, vs, cgs, kbws, leads):
        if panel != cur_panel:
            del X
            with native.NativePca(n, dtype=ndt, stream=stream, max_multiplicity=1) as g:
                if panel == 0:
                    X = torch.empty((n, ld // 2 if dt == "e2m1" else ld), dtype=torch.bfloat16 if dt == "bf16" else torch.uint8, device="cuda")
                    g.synthDenseDevice(20240901, 0, vmax, 0, X.data_ptr(), ld)
                else:
                    X = torch.empty(g.panelBytes(vmax, panel) + 64, dtype=torch.uint8, device="cuda")
                    g.synthPanelsDevice(20240901, 0, vmax, 0, X.data_ptr(), panel)
            torch.cuda.synchronize()
            cur_panel = panel
        os.environ["VPCA_CTA_GROUP"] = str(cg)
        os.environ["VPCA_SYNC_LEAD"] = str(lead)
        if kbw > 0: os.environ["VPCA_KB_WINDOW"] = str(kbw)
        else: os.environ.pop("VPCA_KB_WINDOW", None)
        with native.NativePca(n, dtype=ndt, stream=stream, max_multiplicity=1) as nat:
            ts = []; seq = []
            for r in range(reps + 2):
                nat.reset()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                if panel == 0: nat.accumulateDenseDevice(X.data_ptr(), v, ld)
                else: nat.accumulatePanels(X.data_ptr(), v, panel)
                b.record(); b.synchronize()
                seq.append(round(a.elapsed_time(b), 3))
                if r >= 2: ts.append(a.elapsed_time(b))
            pr = nat.gramProfile()
            t0 = pr[:, 0].min()
            starts, mma, ends = pr[:, 0] - t0, pr[:, 2][pr[:, 2] > 0] - t0, pr[:, 3] - t0
            prof = {"start_spread_us": round(float(starts.max()) / 1e3, 1),
                    "mma_done_us_min_med_max": [round(float(x) / 1e3, 1) for x in (mma.min(), sorted(mma)[len(mma) // 2], mma.max())],
                    "end_us_min_max": [round(float(ends.min()) / 1e3, 1), round(float(ends.max()) / 1e3, 1)]}
            nat.finalizeGram()
            S = torch.from_numpy(nat.getGram())
            chk = int(S.to(torch.int64).sum().item())
            ts.sort()
            ms = ts[len(ts) // 2]
            ops = n * (n + 1) * v
            print(json.dumps({"n": n, "v": v, "panel": panel, "cg": cg, "kbw": kbw, "lead": lead, "prof": prof, "ms_seq": seq, "ms_med": round(ms, 4), "ms_min": round(ts[0], 4),
                              "tops_syrk": round(ops / ms / 1e9, 1), "cells_per_s": round(n * v / ms * 1e3 / 1e9, 2),
                              "checksum": chk}), flush=True)
main()
