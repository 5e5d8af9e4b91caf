#!/usr/bin/env python
"""Time centering + eigensolve (vpca_compute_pca) alone on structured Grams of several sizes."""
import json, os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from spark_examples_b200 import native

def main():
    sizes = [int(x) for x in os.environ.get("EIG_N", "1092,2504,4096").split(",")]
    reps = int(os.environ.get("EIG_REPS", "5"))
    ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
    for n in sizes:
        nv = 200_000
        X = torch.empty((n, nv), dtype=torch.int8, device="cuda")
        with native.NativePca(n, stream=ts.cuda_stream, max_multiplicity=1) as nat:
            nat.synthDenseDevice(20240901, 0, nv, 0, X.data_ptr(), nv)
            nat.accumulateDenseDevice(X.data_ptr(), nv, nv)
            nat.finalizeGram()
            for mode in os.environ.get("EIG_MODES", "direct,auto").split(","):
                os.environ["VPCA_EIG"] = mode
                k = int(os.environ.get("EIG_K", "2"))
                times = []
                for r in range(reps + 1):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); vecs, evals, nz = nat.computePca(k); b.record(); b.synchronize()
                    if r: times.append(a.elapsed_time(b))
                st = nat.stats()
                vecs = vecs[:, :2]
                S = torch.from_numpy(nat.getGram()).cuda().double()
                rs = S.sum(1); C = S - (rs / n)[:, None] - (rs / n)[None, :] + rs.sum() / n / n
                w, V = torch.linalg.eigh(C)
                Vt = V[:, [-1, -2]].cpu().numpy()
                import numpy as np
                for c in range(2):
                    i = int(np.argmax(np.abs(Vt[:, c])));  Vt[:, c] *= (1 if Vt[i, c] > 0 else -1)
                err = float((np.abs(vecs - Vt).max(0) / np.abs(Vt).max(0)).max())
                times.sort()
                if os.environ.get("VPCA_LZ_PROF") == "1":
                    pr = nat.lanczosProfile()
                    pr = pr[(pr[:, 0] > 0)]
                    if len(pr) > 2:
                        step = np.diff(pr[:, 0])
                        med = lambda x: round(float(np.median(x)) / 1e3, 2)
                        print(json.dumps({"n": n, "lanczos_phase_us": {
                            "step_total": med(step[step > 0]), "stage_w_and_norms": med(pr[:, 1] - pr[:, 0]),
                            "matvec": med(pr[:, 2] - pr[:, 1]), "y_and_basis_column": med(pr[:, 3] - pr[:, 2]),
                            "shares_of_VTy_and_VTv": med(pr[:, 4] - pr[:, 3]), "barrier_1": med(pr[:, 5] - pr[:, 4]),
                            "fused_gram_schmidt": med(pr[:, 6] - pr[:, 5]), "write_w": med(pr[:, 7] - pr[:, 6]),
                            "barrier_2": round(med(step[step > 0]) - med(pr[:, 7] - pr[:, 0]), 2)}}), flush=True)
                print(json.dumps({"n": n, "mode": mode, "k": k, "method": st["eig_method"], "iters": st["eig_iterations"],
                                  "device_ms": round(st["last_eig_ms"], 3), "eig_ms_med": round(times[len(times)//2], 3), "eig_ms_min": round(times[0], 3),
                                  "us_per_step": round(times[len(times)//2] * 1e3 / n, 2), "max_rel_err": err}), flush=True)
        del X
main()
