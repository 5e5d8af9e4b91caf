#!/usr/bin/env python
"""BASELINE configs[3] on all GPUs of one box (two forms, --mode): 100 000 samples x 500 000 variants, the 40 GB int32 Gram held ONCE across the
box as row bands (rank q allocates only the rows it owns: 5 GB at 8 GPUs), variants sharded over the GPUs
(`for (c1 <- callset; c2 <- callset) matrix(c1, c2) += 1`, VariantsPca.scala:186-188, one partition matrix per GPU, with the
sizing note of :176-177 answered by never materialising a second copy).  One process drives every GPU -- the process model of
the reference's `local[*]` driver JVM (VariantsPca.scala:38-50) -- through vpca_gram_set_peers_local: every Gram kernel's
epilogue adds its tile straight into the band of the rank that owns the row (red.relaxed.sys over NVLink), so there is no
reduce step and no gather; the bands ARE the result.  `--mode owner-computes` is the other form SURVEY 8e names: every GPU
holds ALL variants (50 GB of int8 genotypes) and its Gram kernel enumerates only the tiles of the rows it stores -- no peers,
no traffic between the GPUs at all.

Times the Gram launches with CUDA events per device (max over devices = the job), and checks without an N x N oracle:
  * diag(S) = carrier counts of the whole cohort, per band;
  * whole sampled rows of every band (its first, its last and random ones: the lower-triangle part, columns 0..row) and
    random 256 x 256 blocks against an exact fp32 matmul of the same shards (0/1 cells, counts < 2^24, TF32 off).
vpca_compute_pca is not called: like MLlib's RowMatrix it is limited to 65 535 samples (VariantsPca.scala:226)."""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
from spark_examples_b200 import native


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=100_000)
    ap.add_argument("--variants", type=int, default=500_000, help="whole cohort; split evenly over the GPUs")
    ap.add_argument("--gpus", type=int, default=0, help="0 = all visible")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--panel", type=int, default=8192)
    ap.add_argument("--check-rows", type=int, default=3, help="sampled whole rows per band checked against fp32 matmul")
    ap.add_argument("--mode", choices=["owner-flush", "owner-computes"], default="owner-flush",
                    help="owner-flush: every GPU holds a variant shard and its kernel adds each tile into the band of the row's "
                         "owner over NVLink; owner-computes: every GPU holds ALL variants and computes only its own band (no "
                         "traffic between GPUs at all; 50 GB of genotypes per GPU)")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    world = args.gpus or torch.cuda.device_count()
    n, P = args.samples, args.panel
    computes = args.mode == "owner-computes"
    per = args.variants if computes else (args.variants + world - 1) // world
    bands = native.ownerRowBands(n, world)
    ctxs, bufs, streams = [], [], []
    total_v = per if computes else per * world
    report = {"config": f"{n} samples x {total_v} variants over {world} GPU(s), one process, band-only Grams, {args.mode}",
              "n": n, "variants_per_gpu": per, "world": world, "panel": P,
              "band_rows": [b[1] for b in bands], "band_gb": [round(b[1] * n * 4 / 2 ** 30, 2) for b in bands]}
    try:
        for r in range(world):
            torch.cuda.set_device(r)
            s = torch.cuda.Stream(device=r)
            streams.append(s)
            ctxs.append(native.NativePca(n, device=r, stream=s.cuda_stream, max_multiplicity=1, gram_band=bands[r]))
        if not computes:
            native.setPeersLocal(ctxs, "owner_rows")
        for r, c in enumerate(ctxs):
            torch.cuda.set_device(r)
            buf = torch.zeros(c.panelBytes(per, P), dtype=torch.uint8, device=f"cuda:{r}")
            bufs.append(buf)
            c.synthPanelsDevice(20240901, 0 if computes else r * per, per, 0, buf.data_ptr(), P)
        for c in ctxs:
            c.synchronize()
        times = []
        for rep in range(args.reps + 1):
            for c in ctxs:
                c.reset()
            for c in ctxs:
                c.synchronize()       # every band is zero before any rank adds into it
            ev = []
            for r, c in enumerate(ctxs):
                torch.cuda.set_device(r)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(streams[r]):
                    a.record()
                    c.accumulatePanels(bufs[r].data_ptr(), per, P)
                    b.record()
                ev.append((a, b))
            if not computes:
                for c in ctxs:
                    c.gatherGram()    # closing all-rank barrier on every stream: all remote adds have landed
            for c in ctxs:
                c.synchronize()
            per_dev = [a.elapsed_time(b) for a, b in ev]
            if rep:
                times.append(per_dev)
        ms_dev = np.median(np.array(times), axis=0)
        ms = float(ms_dev.max())
        ops = float(n) * (n + 1) * total_v
        st = [c.stats() for c in ctxs]
        report.update({
            "gram_ms_per_device_median": [round(float(x), 2) for x in ms_dev],
            "gram_ms_job": round(ms, 2),
            "cells_per_s": n * total_v / (ms * 1e-3),
            "syrk_tops_per_gpu": round(ops / world / (ms * 1e-3) / 1e12, 1),
            "frac_of_nominal_int8_4500": round(ops / world / (ms * 1e-3) / 1e12 / 4500.0, 3),
            "resident_schedule": [s["gram_resident"] for s in st],
            # lower triangle of S, minus the part whose owner is the writer itself, crosses NVLink once per rank
            "remote_red_bytes_per_gpu_upper_bound": 0 if computes else int(n * (n + 1) // 2 * 4 * (world - 1) / world),
            "genotype_bytes_per_gpu": int(n) * int(per),
        })
        print(json.dumps(dict(report, checks="pending")), flush=True)     # the timings survive a failure in the checks below
        if args.out:
            Path(args.out).write_text(json.dumps(dict(report, checks="pending")) + "\n")
        # ---- checks: the rows of X on one device as fp32 (6.25 GB per 62 500-variant shard would be 25 GB as fp32, so
        #      the reference values are computed shard by shard from the int8 panels)
        npan = (per + P - 1) // P

        def shard_rows(r, rows):
            """int8 panels of shard r restricted to `rows` -> (len(rows), npan * P) fp32 on device r"""
            x = bufs[r].view(torch.int8)[: npan * n * P].view(npan, n, P)
            sel = x[:, rows, :]                                            # slice or index list
            return sel.permute(1, 0, 2).reshape(sel.shape[1], npan * P).to(torch.float32)

        shard_devs = [0] if computes else list(range(world))     # owner-computes: every device holds the whole cohort
        carriers = torch.zeros(n, dtype=torch.int64)
        for r in shard_devs:
            x = bufs[r].view(torch.int8)[: npan * n * P].view(npan, n, P)
            for p_ in range(npan):                                   # panel by panel: the int64 sum of a whole shard would not fit
                carriers += (x[p_] != 0).sum(dim=1).cpu()
        ok_diag, ok_rows, ok_blocks = True, True, True
        rng = np.random.default_rng(7)
        for q, c in enumerate(ctxs):
            row0, rows = bands[q]
            pick = sorted({row0, row0 + rows - 1} | set(int(v) for v in rng.integers(row0, row0 + rows, args.check_rows)))
            got = {row: c.gramBand(row, 1)[0] for row in pick}
            for row in pick:
                ok_diag = ok_diag and int(got[row][row]) == int(carriers[row])
            want = {row: torch.zeros(row + 1, dtype=torch.float64) for row in pick}
            for r in shard_devs:
                torch.cuda.set_device(r)
                xr = shard_rows(r, pick)                                   # (len(pick), K)
                for r0 in range(0, max(pick) + 1, 8192):             # columns beyond the last sampled row are never needed
                    r1 = min(n, r0 + 8192)
                    blk = (xr @ shard_rows(r, slice(r0, r1)).t()).to(torch.float64).cpu()   # exact: counts < 2^24
                    for i, row in enumerate(pick):
                        hi = min(r1, row + 1)
                        if hi > r0:
                            want[row][r0:hi] += blk[i, : hi - r0]
            for row in pick:
                ok_rows = ok_rows and bool(np.array_equal(got[row][: row + 1].astype(np.int64), want[row].numpy().astype(np.int64)))
            # one random 256 x 256 block strictly inside the band's lower-triangle part
            br = int(rng.integers(row0, max(row0 + 1, row0 + rows - 256)))
            bc = int(rng.integers(0, max(1, br - 256)))
            gb = c.gramBand(br, min(256, row0 + rows - br))[:, bc:bc + 256].astype(np.int64)
            wb = torch.zeros(gb.shape, dtype=torch.float64)
            for r in shard_devs:
                torch.cuda.set_device(r)
                wb += (shard_rows(r, slice(br, br + gb.shape[0])) @ shard_rows(r, slice(bc, bc + gb.shape[1])).t()).to(torch.float64).cpu()
            ok_blocks = ok_blocks and bool(np.array_equal(gb, wb.numpy().astype(np.int64)))
        report["checks"] = {"diag_equals_carrier_counts": ok_diag, "sampled_rows_exact_vs_fp32_matmul": ok_rows,
                            "random_256_blocks_exact": ok_blocks, "rows_checked_per_band": args.check_rows + 2}
    finally:
        for c in ctxs:
            try:
                c.synchronize()
            except Exception:
                pass
        for c in ctxs:
            c.close()
    line = json.dumps(report)
    print(line, flush=True)
    if args.out:
        Path(args.out).write_text(line + "\n")


main()
