#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the oracle (numpy restatement).  These are REGRESSION vectors (reference-produced
vectors live in reference_twin.json, see make_reference_twin_golden.py).  For the eigen step PARITY IS UNPINNED: the reference
has no fixtures of its own and cannot be run here (no JVM/Spark), so these vectors pin the ORACLE (and,
through the GPU tests, the CUDA path) against regressions; the hand-computed cases in tests/test_oracle.py
and the README.md:109-119 magnitude property are the independent anchors.

    python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as o   # noqa: E402

SEED = 20240901
OUT = Path(__file__).resolve().parent


def cohort(n, nv, name, k=2):
    X = o.np_synth_dense(SEED, n, 0, nv, 0)
    off, idx = o.dense_to_calls(X)
    S = o.np_similarity(n, [idx[off[v]:off[v + 1]] for v in range(len(off) - 1)])
    C, rs, nz = o.np_center(S)
    U, sv = o.mllib_principal_components(C, k)
    U = o.sign_normalise(U)
    w = np.linalg.eigvalsh(C)[::-1][:k]
    np.savez_compressed(OUT / f"{name}.npz", n=n, nv=nv, seed=SEED, offsets=off, idx=idx, S=S, row_sums=rs,
                        non_zero_rows=nz, C_checksum=np.array([C.sum(), np.abs(C).sum(), C[0, 0], C[n // 2, n // 3]]),
                        U=U, cov_singular_values=sv, eigenvalues=w,
                        X_packed=np.packbits(X.astype(np.uint8), axis=1))
    print(name, "S.sum", int(S.sum()), "top eig", w)


def generator_vectors():
    v = np.array([0, 1, 2, 12345, 999_999, 39_999_999, 2 ** 33 + 7], dtype=np.uint64)
    thr = o.np_variant_thresholds(SEED, v)
    tile = o.np_synth_dosage(SEED, 37, 1000, 53)
    np.savez_compressed(OUT / "generator.npz", seed=SEED, variants=v, thresholds=thr, dosage_n37_v1000_53=tile,
                        pop_bounds_2504=o.np_pop_bounds(2504), pop_bounds_1092=o.np_pop_bounds(1092))
    print("generator", thr[0])


if __name__ == "__main__":
    cohort(48, 200, "cohort_n48_v200")
    cohort(200, 1500, "cohort_n200_v1500")
    generator_vectors()
