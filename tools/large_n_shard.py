#!/usr/bin/env python
"""One GPU's share of a biobank-scale Gram (BASELINE configs[3]: 100 000 samples x 500 000 variants over 8 GPUs =
62 500 variants per GPU, the 40 GB int32 Gram resident in HBM).  Times the Gram launch alone with CUDA events and
checks it with size-independent properties plus exact sub-blocks (no N x N oracle exists at this size):
  symmetry, diag(S) = carrier counts, S.1 = X (X^T 1), and random 256 x 256 blocks against an exact fp32 matmul.
vpca_compute_pca is not called: like MLlib's RowMatrix it is limited to 65 535 samples (VariantsPca.scala:226)."""
import argparse, json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from spark_examples_b200 import native

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=100_000)
    ap.add_argument("--variants", type=int, default=62_500)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--panel", type=int, default=8192)
    args = ap.parse_args()
    n, nv, P = args.samples, args.variants, args.panel
    dev = torch.device("cuda:0")
    ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
    npan = (nv + P - 1) // P
    S = torch.zeros((n, n), dtype=torch.int32, device=dev)
    X = torch.zeros((npan, n, P), dtype=torch.int8, device=dev)    # cells past nv in the last panel stay zero
    with native.NativePca(n, stream=ts.cuda_stream, d_gram=S.data_ptr(), max_multiplicity=1) as nat:
        nat.synthPanelsDevice(20240901, 0, nv, 0, X.data_ptr(), P)
        torch.cuda.synchronize()
        kt = []
        for r in range(args.reps + 2):
            nat.reset()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); nat.accumulatePanels(X.data_ptr(), nv, P); b.record(); b.synchronize()
            if r >= 2: kt.append(a.elapsed_time(b))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); nat.finalizeGram(); b.record(); b.synchronize()
        fin_ms = a.elapsed_time(b)
        st = nat.stats()
    ms = sum(kt) / len(kt)
    checks = {}
    # the generator leaves cells past nv in the last panel zero, so whole panels can be used below
    RB = 2048
    sym = True
    for r0 in range(0, n, RB):
        r1 = min(n, r0 + RB)
        sym = sym and bool(torch.equal(S[r0:r1, :], S[:, r0:r1].t()))
    checks["gram_symmetric"] = sym
    carriers = torch.zeros(n, dtype=torch.int64, device=dev)
    colsums = []
    for p in range(npan):
        nz = (X[p] != 0)
        carriers += nz.sum(dim=1)
        colsums.append(nz.sum(dim=0).to(torch.float64))
    checks["diag_equals_carrier_counts"] = bool(torch.equal(torch.diagonal(S).to(torch.int64), carriers))
    s1 = torch.zeros(n, dtype=torch.float64, device=dev)
    for p in range(npan):
        for r0 in range(0, n, 16384):
            r1 = min(n, r0 + 16384)
            s1[r0:r1] += (X[p, r0:r1] != 0).to(torch.float64) @ colsums[p]
    rows = torch.cat([S[r0:min(n, r0 + RB)].sum(dim=1, dtype=torch.int64) for r0 in range(0, n, RB)])
    checks["S_times_ones_equals_X_Xt1"] = bool(torch.equal(rows.to(torch.float64), s1))
    g = torch.Generator().manual_seed(3)
    ok = True
    for _ in range(6):
        i0 = int(torch.randint(0, n - 256, (1,), generator=g)); j0 = int(torch.randint(0, n - 256, (1,), generator=g))
        acc = torch.zeros((256, 256), dtype=torch.float32, device=dev)
        for p in range(npan):
            acc += (X[p, i0:i0 + 256] != 0).float() @ (X[p, j0:j0 + 256] != 0).float().t()   # exact: sums < 2^24
        ok = ok and bool(torch.equal(S[i0:i0 + 256, j0:j0 + 256], acc.to(torch.int32)))
    checks["random_blocks_equal_exact_matmul"] = ok
    ops = float(n) * (n + 1) * nv
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    peak = 2.0 * float(peaks.get("bf16_tflops", 1590.0))
    print(json.dumps({"workload": f"{n} samples x {nv} variants int8 (one GPU's share), Gram {4 * n * n / 1e9:.1f} GB in HBM",
                      "gram_ms": ms, "gram_ms_all": kt, "finalize_ms": fin_ms, "cells_per_s": n * nv / (ms * 1e-3),
                      "syrk_min_tops": ops / (ms * 1e-3) / 1e12, "peak_tops_2x_measured_bf16": peak,
                      "frac": ops / (ms * 1e-3) / 1e12 / peak, "gram_launches_total": st["gram_launches"],
                      "resident": st["gram_resident"], "checks": checks}))

main()
