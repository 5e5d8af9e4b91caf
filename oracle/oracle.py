"""CPU oracle for the VariantsPca hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module.  The product path (``spark_examples_b200``) never does and fails
loudly when its CUDA library is missing.

PINNING: the reference ships no tests/golden vectors for this path and its Scala driver cannot run here (no JVM, no
Spark).  Its Python twin (src/main/python/variants_pca.py) is Python-2 + py4j, but the three functions of it that restate
the encode, similarity and centering steps (:19-121) are pure Python over RDD operations: tests/golden/
make_reference_twin_golden.py EXECUTES them (mechanical 2-to-3 rewrites, an in-memory RDD stand-in) and
tests/test_reference_twin.py holds this module to their output bit for bit.  The eigen step (`perform_pca`, :123-152)
needs the JVM: for that step PARITY IS UNPINNED (upstream's RowMatrixSuite known-answer case is the only anchor).
This module holds

* ``np_*``  -- a pure-numpy restatement, written independently of the C one, and
* ``c_*``   -- ctypes bindings to ``oracle/vpca_oracle.c`` (the timed CPU baseline),

both following /root/reference/src/main/scala/com/google/cloud/genomics/spark/examples/
VariantsPca.scala (cited per function) and, for the eigen step, the published source of the
un-vendored dependency spark-mllib 1.6.1 (build.sbt:11,25).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libvpca_oracle.so"
_lib = None

NPOP = 5
POP_CUM_PCT = (26, 40, 60, 80, 100)
POP_FST = (0.15, 0.10, 0.07, 0.05, 0.03)
TAG_VARIANT = 0xA11E1E00
TAG_CELL = 0xC0FFEE00
IH4_SD = 147.80054127


def build(force: bool = False) -> Path:
    """Compile oracle/vpca_oracle.c with the committed Makefile."""
    src = _HERE / "vpca_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(str(_LIB_PATH))
        i32p = ctypes.POINTER(ctypes.c_int32)
        i64p = ctypes.POINTER(ctypes.c_int64)
        f64p = ctypes.POINTER(ctypes.c_double)
        i8p = ctypes.POINTER(ctypes.c_int8)
        u32p = ctypes.POINTER(ctypes.c_uint32)
        L.vo_encode_calls.restype = ctypes.c_int64
        L.vo_encode_calls.argtypes = [ctypes.c_int32, ctypes.c_int64, i64p, i32p, i64p, i32p, i64p, i32p]
        L.vo_similarity.restype = ctypes.c_int
        L.vo_similarity.argtypes = [ctypes.c_int32, ctypes.c_int64, i64p, i32p, ctypes.c_int32, i32p]
        L.vo_similarity_stream.restype = ctypes.c_int
        L.vo_similarity_stream.argtypes = [ctypes.c_int32, ctypes.c_int64, i64p, i32p, i32p]
        L.vo_center.restype = None
        L.vo_center.argtypes = [ctypes.c_int32, i32p, f64p, f64p, i32p]
        L.vo_mllib_covariance.restype = None
        L.vo_mllib_covariance.argtypes = [ctypes.c_int32, f64p, f64p]
        L.vo_philox4x32_10.restype = None
        L.vo_philox4x32_10.argtypes = [u32p, u32p, u32p]
        L.vo_pop_bounds.restype = None
        L.vo_pop_bounds.argtypes = [ctypes.c_int32, i32p]
        L.vo_variant_thresholds.restype = None
        L.vo_variant_thresholds.argtypes = [ctypes.c_uint64, ctypes.c_int64, u32p]
        L.vo_synth_dense.restype = None
        L.vo_synth_dense.argtypes = [ctypes.c_uint64, ctypes.c_int32, ctypes.c_int64, ctypes.c_int64,
                                     ctypes.c_int, ctypes.c_int64, i8p]
        L.vo_synth_calls.restype = ctypes.c_int64
        L.vo_synth_calls.argtypes = [ctypes.c_uint64, ctypes.c_int32, ctypes.c_int64, ctypes.c_int64,
                                     i64p, i32p, i64p]
        L.vo_num_threads.restype = ctypes.c_int
        L.vo_set_threads.restype = None
        L.vo_set_threads.argtypes = [ctypes.c_int]
        _lib = L
    return _lib


def _p(a: np.ndarray, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


# ----------------------------------------------------------------------------------------------
# C bindings
# ----------------------------------------------------------------------------------------------
def c_num_threads() -> int:
    return int(lib().vo_num_threads())


def c_set_threads(n: int) -> None:
    lib().vo_set_threads(int(n))


def c_encode_calls(n_samples, call_off, callset, gt_off, genotype):
    """VariantsPca.scala:56-60 + :153-168 on columnar Variant records -> CSR (offsets, idx)."""
    call_off = np.ascontiguousarray(call_off, np.int64)
    callset = np.ascontiguousarray(callset, np.int32)
    gt_off = np.ascontiguousarray(gt_off, np.int64)
    genotype = np.ascontiguousarray(genotype, np.int32)
    nv = len(call_off) - 1
    out_off = np.zeros(nv + 1, np.int64)
    out_idx = np.zeros(max(len(callset), 1), np.int32)
    nv_out = lib().vo_encode_calls(n_samples, nv, _p(call_off, ctypes.c_int64), _p(callset, ctypes.c_int32),
                                   _p(gt_off, ctypes.c_int64), _p(genotype, ctypes.c_int32),
                                   _p(out_off, ctypes.c_int64), _p(out_idx, ctypes.c_int32))
    if nv_out < 0:
        raise IndexError("callset index outside [0, n_samples)")
    out_off = out_off[: nv_out + 1].copy()
    return out_off, out_idx[: out_off[-1]].copy()


def c_similarity(n, off, idx, n_partitions=1):
    """VariantsPca.scala:182-191 (getSimilarityMatrix): int32 N x N, all entries present."""
    off = np.ascontiguousarray(off, np.int64)
    idx = np.ascontiguousarray(idx, np.int32)
    S = np.zeros((n, n), np.int32)
    rc = lib().vo_similarity(n, len(off) - 1, _p(off, ctypes.c_int64), _p(idx, ctypes.c_int32),
                             int(n_partitions), _p(S, ctypes.c_int32))
    if rc != 0:
        raise IndexError("sample index outside [0, n)")
    return S


def c_similarity_stream(n, off, idx):
    """VariantsPca.scala:262-279 (getSimilarityMatrixStream), dense result."""
    off = np.ascontiguousarray(off, np.int64)
    idx = np.ascontiguousarray(idx, np.int32)
    S = np.zeros((n, n), np.int32)
    rc = lib().vo_similarity_stream(n, len(off) - 1, _p(off, ctypes.c_int64), _p(idx, ctypes.c_int32),
                                    _p(S, ctypes.c_int32))
    if rc != 0:
        raise IndexError("sample index outside [0, n)")
    return S


def c_center(S):
    """VariantsPca.scala:199-223 -> (C float64 N x N, rowSums, nonZeroRows)."""
    S = np.ascontiguousarray(S, np.int32)
    n = S.shape[0]
    C = np.zeros((n, n), np.float64)
    rs = np.zeros(n, np.float64)
    nz = ctypes.c_int32(0)
    lib().vo_center(n, _p(S, ctypes.c_int32), _p(C, ctypes.c_double), _p(rs, ctypes.c_double), ctypes.byref(nz))
    return C, rs, int(nz.value)


def c_mllib_covariance(C):
    C = np.ascontiguousarray(C, np.float64)
    n = C.shape[0]
    Cov = np.zeros((n, n), np.float64)
    lib().vo_mllib_covariance(n, _p(C, ctypes.c_double), _p(Cov, ctypes.c_double))
    return Cov


def c_philox(ctr, key):
    c = np.asarray(ctr, np.uint32).copy()
    k = np.asarray(key, np.uint32).copy()
    o = np.zeros(4, np.uint32)
    lib().vo_philox4x32_10(_p(c, ctypes.c_uint32), _p(k, ctypes.c_uint32), _p(o, ctypes.c_uint32))
    return o


def c_variant_thresholds(seed, v):
    o = np.zeros(NPOP, np.uint32)
    lib().vo_variant_thresholds(ctypes.c_uint64(seed), ctypes.c_int64(v), _p(o, ctypes.c_uint32))
    return o


def c_synth_dense(seed, n, v0, nv, mode=0):
    """Dense sample-major int8 tile X[s, v - v0]; mode 0 = binary carrier, 1 = dosage."""
    X = np.zeros((n, nv), np.int8)
    lib().vo_synth_dense(ctypes.c_uint64(seed), n, v0, nv, mode, nv, _p(X, ctypes.c_int8))
    return X


def c_synth_calls(seed, n, v0, nv):
    """RDD[Seq[Int]] form (CSR offsets, idx) of the synthetic cohort, empty variants dropped."""
    off = np.zeros(nv + 1, np.int64)
    kept = ctypes.c_int64(0)
    nnz = lib().vo_synth_calls(ctypes.c_uint64(seed), n, v0, nv, _p(off, ctypes.c_int64), None, ctypes.byref(kept))
    idx = np.zeros(max(int(nnz), 1), np.int32)
    lib().vo_synth_calls(ctypes.c_uint64(seed), n, v0, nv, _p(off, ctypes.c_int64), _p(idx, ctypes.c_int32),
                         ctypes.byref(kept))
    return off[: kept.value + 1].copy(), idx[: int(nnz)]


# ----------------------------------------------------------------------------------------------
# numpy restatement (independent of the C code above)
# ----------------------------------------------------------------------------------------------
def np_extract_call_info(calls, mapping):
    """VariantsPca.scala:56-60.  calls: iterable of (callsetId:str, genotype:list[int]) or None.
    Returns [(hasVariation, callsetIndex)]; unknown id raises KeyError (NoSuchElementException)."""
    out = []
    for callset_id, genotype in (calls or []):
        has_variation = False
        for allele in genotype:                      # foldLeft(false)(_ || _ > 0)
            has_variation = has_variation or allele > 0
        out.append((has_variation, mapping[callset_id]))
    return out


def np_get_calls(variants, mapping):
    """VariantsPca.scala:153-168, single-dataset branch (:157, :163-167) -> list[list[int]]."""
    rows = []
    for calls in variants:
        info = np_extract_call_info(calls, mapping)
        kept = [idx for has_var, idx in info if has_var]          # :164
        if len(kept) > 0:                                         # :166
            rows.append(kept)                                     # :167
    return rows


# -- multi-dataset keying (VariantsPca.scala:62-78, :115-148).  The hash is Guava's `Hashing.murmur3_128()` --
#    com.google.guava, pulled in through google-genomics-utils and shaded at build.sbt:44, NOT under /root/reference: the
#    published MurmurHash3_x64_128 algorithm (Austin Appleby, public domain; seed 0) restated here with numpy uint64
#    arithmetic, pinned on Guava's known answers in tests/test_oracle.py.
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _u64(x):
    return np.uint64(int(x) & 0xFFFFFFFFFFFFFFFF)


def _rotl(x, r):
    x = int(x)
    return _u64((x << r) | (x >> (64 - r)))


def _mul(a, b):
    return _u64(int(a) * int(b))


def _fmix(k):
    k = int(k)
    k ^= k >> 33
    k = (k * 0xFF51AFD7ED558CCD) & 0xFFFFFFFFFFFFFFFF
    k ^= k >> 33
    k = (k * 0xC4CEB9FE1A85EC53) & 0xFFFFFFFFFFFFFFFF
    k ^= k >> 33
    return _u64(k)


def np_murmur3_128(data: bytes) -> bytes:
    """`Hashing.murmur3_128().hashBytes(data).asBytes()`: 16 bytes, h1 then h2, little-endian."""
    c1, c2 = _u64(0x87C37B91114253D5), _u64(0x4CF5AD432745937F)
    h1 = h2 = _u64(0)
    n = len(data)
    body = np.frombuffer(data[: n - n % 16], dtype="<u8").reshape(-1, 2)
    for k1, k2 in body:
        h1 = _u64(int(h1) ^ int(_mul(_rotl(_mul(k1, c1), 31), c2)))
        h1 = _u64(int(_rotl(h1, 27)) + int(h2))
        h1 = _u64(int(h1) * 5 + 0x52DCE729)
        h2 = _u64(int(h2) ^ int(_mul(_rotl(_mul(k2, c2), 33), c1)))
        h2 = _u64(int(_rotl(h2, 31)) + int(h1))
        h2 = _u64(int(h2) * 5 + 0x38495AB5)
    tail = data[n - n % 16:]
    if len(tail) > 8:
        k2 = _u64(int.from_bytes(tail[8:], "little"))
        h2 = _u64(int(h2) ^ int(_mul(_rotl(_mul(k2, c2), 33), c1)))
    if len(tail) > 0:
        k1 = _u64(int.from_bytes(tail[:8], "little"))
        h1 = _u64(int(h1) ^ int(_mul(_rotl(_mul(k1, c1), 31), c2)))
    h1 = _u64(int(h1) ^ n)
    h2 = _u64(int(h2) ^ n)
    h1 = _u64(int(h1) + int(h2))
    h2 = _u64(int(h2) + int(h1))
    h1, h2 = _fmix(h1), _fmix(h2)
    h1 = _u64(int(h1) + int(h2))
    h2 = _u64(int(h2) + int(h1))
    return int(h1).to_bytes(8, "little") + int(h2).to_bytes(8, "little")


def np_variant_key(contig, start, end, reference_bases, alternate_bases):
    """VariantsPca.scala:62-78 getVariantKey: putString(contig) putLong(start) putLong(end) putString(referenceBases)
    putString(alternateBases.mkString("")) -> HashCode.toString (hex of asBytes).  Guava writes longs little-endian
    and, for `putString(s, UTF_8)`, the UTF-8 bytes."""
    alt = "".join(alternate_bases or [])
    payload = (contig.encode("utf-8") + int(start).to_bytes(8, "little", signed=True) +
               int(end).to_bytes(8, "little", signed=True) + (reference_bases or "").encode("utf-8") + alt.encode("utf-8"))
    return np_murmur3_128(payload).hex()


def np_join_datasets(left, right):
    """VariantsPca.scala:115-128 joinDatasets on two datasets of (key, calls) records: `keyBy` + `join` (inner, one output
    per pair of records with equal keys) and `related._1 ++ related._2`.  Spark leaves the order of the joined RDD
    unspecified; this restatement lists pairs by left record, then right record."""
    by_key = {}
    for key, calls in right:
        by_key.setdefault(key, []).append(calls)
    return [list(lc) + list(rc) for key, lc in left for rc in by_key.get(key, ())]


def np_merge_datasets(datasets, variant_set_count):
    """VariantsPca.scala:136-148 mergeDatasets: union of all datasets, `groupByKey`, keep the keys whose group has exactly
    `variantSetCount` records (:144), flatten the group's calls (:145).  Groups are listed by first appearance, members
    in union order."""
    groups = {}
    for ds in datasets:
        for key, calls in ds:
            groups.setdefault(key, []).append(calls)
    return [[c for calls in g for c in calls] for g in groups.values() if len(g) == variant_set_count]


def np_similarity(n, rows):
    """VariantsPca.scala:182-191 with one partition; rows: list of index lists."""
    S = np.zeros((n, n), np.int64)
    for callset in rows:
        cs = np.asarray(callset, np.int64)
        if len(cs) and (cs.min() < 0 or cs.max() >= n):
            raise IndexError("sample index outside [0, n)")
        np.add.at(S, (cs[:, None], cs[None, :]), 1)               # for (c1 <- cs; c2 <- cs) M(c1,c2) += 1
    assert S.max(initial=0) < 2 ** 31
    return S.astype(np.int32)


def np_similarity_dense(X):
    """Same S from the dense multiplicity matrix X (samples x variants): S = X X^T, exact in int64."""
    X64 = np.asarray(X, np.int64)
    S = X64 @ X64.T
    assert S.max(initial=0) < 2 ** 31
    return S.astype(np.int32)


def np_center(S):
    """VariantsPca.scala:199-223; returns (C, rowSums, nonZeroRows)."""
    n = S.shape[0]
    Sd = S.astype(np.float64)
    row_sums = np.zeros(n)
    for j in range(n):                                            # foldLeft(0D)(_ + _._2), column order
        row_sums = row_sums + Sd[:, j]
    non_zero = int((row_sums > 0).sum())                          # :207
    matrix_sum = 0.0
    for i in range(n):                                            # reduce(_ + _)
        matrix_sum = matrix_sum + row_sums[i] if i else row_sums[0]
    row_count = float(n)
    matrix_mean = matrix_sum / row_count / row_count              # :211
    row_mean = row_sums / row_count                               # :216 / :220
    C = ((Sd - row_mean[:, None]) - row_mean[None, :]) + matrix_mean   # :221, left to right
    return C, row_sums, non_zero


def mllib_principal_components(C, k):
    """spark-mllib 1.6.1 RowMatrix.computePrincipalComponents(k) as called at VariantsPca.scala:225-226.

    computeCovariance: mean = column means; G = C^T C; Cov = G/(m-1) - m/(m-1) mean mean^T.
    Then brzSvd(Cov) (LAPACK dgesdd, which numpy.linalg.svd also calls) and the first k columns
    of U, returned column-major like ``pca.toArray`` (:227).  Requires n <= 65535 as upstream does.
    """
    n = C.shape[0]
    if n > 65535:
        raise ValueError("RowMatrix.computeCovariance: Argument with more than 65535 cols")
    m = float(n)
    mean = C.sum(axis=0) / m
    G = C.T @ C
    Cov = G / (m - 1.0) - (m / (m - 1.0)) * np.outer(mean, mean)
    U, s, _ = np.linalg.svd(Cov, full_matrices=True)
    return U[:, :k].copy(), s[:k].copy()


def compute_pca(S, k=2):
    """VariantsPca.scala:198-231 end to end on a dense S -> (N x k eigvec matrix, singular values of Cov)."""
    C, _, _ = np_center(S)
    return mllib_principal_components(C, k)


def sign_normalise(U):
    """flip each column so its largest-|.| entry (lowest index on ties) is positive."""
    U = np.array(U, np.float64, copy=True)
    if U.ndim == 1:
        U = U[:, None]
    for c in range(U.shape[1]):
        i = int(np.argmax(np.abs(U[:, c])))
        if U[i, c] < 0:
            U[:, c] = -U[:, c]
    return U


def eigvec_rel_err(U_test, U_ref):
    """max_i |u^_i - u_i| / max_i |u_i| per component after sign normalisation."""
    A, B = sign_normalise(U_test), sign_normalise(U_ref)
    return np.max(np.abs(A - B), axis=0) / np.max(np.abs(B), axis=0)


def emit_result_lines(result, names):
    """VariantsPca.scala:233-240: 'name\\tdataset\\tpc1\\tpc2', sorted by name."""
    rows = []
    for callset_id, pc1, pc2 in result:
        dataset = callset_id.split("-")[0]
        rows.append((names[callset_id], pc1, pc2, dataset))
    rows.sort(key=lambda t: t[0])
    return [f"{t[0]}\t{t[3]}\t{_jdouble(t[1])}\t{_jdouble(t[2])}" for t in rows]


def _jdouble(x: float) -> str:
    """Java Double.toString layout: plain decimal for 1e-3 <= |x| < 1e7, otherwise d.dddE[-]n; the digits
    are the shortest that round-trip (JDK >= 19 behaviour; older JDKs may print one more digit)."""
    x = float(x)
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Infinity" if x > 0 else "-Infinity"
    if x == 0:
        return "-0.0" if np.signbit(x) else "0.0"
    sign = "-" if x < 0 else ""
    digits, exp = f"{abs(x):.17e}".split("e")          # placeholder to get the exponent
    exp = int(exp)
    r = repr(abs(x))
    if "e" in r:
        mant, e2 = r.split("e")
        exp = int(e2)
    else:
        mant = r
    d = mant.replace(".", "").lstrip("0").rstrip("0") or "0"     # significant digits
    if 1e-3 <= abs(x) < 1e7:
        if exp >= 0:
            ip, fp = d[: exp + 1].ljust(exp + 1, "0"), d[exp + 1:]
        else:
            ip, fp = "0", "0" * (-exp - 1) + d
        return f"{sign}{ip}.{fp or '0'}"
    return f"{sign}{d[0]}.{d[1:] or '0'}E{exp}"


# ----------------------------------------------------------------------------------------------
# numpy mirror of the synthetic generator (see vpca_oracle.c for the specification)
# ----------------------------------------------------------------------------------------------
def np_philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(x, np.uint64) for x in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint64(k0)
    k1 = np.uint64(k1)
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    mask = np.uint64(0xFFFFFFFF)
    s32 = np.uint64(32)
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        n0 = (p1 >> s32) ^ c1 ^ k0
        n1 = p1 & mask
        n2 = (p0 >> s32) ^ c3 ^ k1
        n3 = p0 & mask
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return c0, c1, c2, c3


def np_pop_bounds(n):
    return np.array([(c * n) // 100 for c in POP_CUM_PCT], np.int64)


def np_pop_of_sample(n):
    b = np_pop_bounds(n)
    return np.searchsorted(b, np.arange(n), side="right").astype(np.int64)


def np_variant_thresholds(seed, v):
    v = np.asarray(v, np.uint64)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    vlo, vhi = v & np.uint64(0xFFFFFFFF), v >> np.uint64(32)
    r0, r1, r2, r3 = np_philox4x32_10(vlo, vhi, np.uint64(0), np.uint64(TAG_VARIANT), k0, k1)
    r4, r5, _, _ = np_philox4x32_10(vlo, vhi, np.uint64(1), np.uint64(TAG_VARIANT), k0, k1)
    u = (r0.astype(np.float64) + 0.5) * (1.0 / 4294967296.0)
    p = 0.02 + 0.48 * u
    pq = p * (1.0 - p)
    thr = np.zeros(v.shape + (NPOP,), np.uint32)
    for k, w in enumerate((r1, r2, r3, r4, r5)):
        w = w.astype(np.int64)
        sum4 = (w & 255) + ((w >> 8) & 255) + ((w >> 16) & 255) + (w >> 24)
        z = (sum4 - 510).astype(np.float64) / IH4_SD
        pk = p + np.sqrt(POP_FST[k] * pq) * z
        pk = np.where(pk < 0.001, 0.001, pk)
        pk = np.where(pk > 0.999, 0.999, pk)
        thr[..., k] = (pk * 4294967296.0).astype(np.uint64).astype(np.uint32)
    return thr


def np_synth_dosage(seed, n, v0, nv):
    """Dosage matrix G[s, j] in {0,1,2} for variants v0 .. v0+nv."""
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    v = np.arange(v0, v0 + nv, dtype=np.uint64)
    thr = np_variant_thresholds(seed, v)                        # (nv, 5)
    pop = np_pop_of_sample(n)                                   # (n,)
    npair = (n + 1) // 2
    pair = np.arange(npair, dtype=np.uint64)
    w = np_philox4x32_10((v & np.uint64(0xFFFFFFFF))[None, :], (v >> np.uint64(32))[None, :],
                         pair[:, None], np.uint64(TAG_CELL), k0, k1)      # 4 x (npair, nv)
    G = np.zeros((2 * npair, nv), np.int8)
    T = thr[:, pop].T.astype(np.uint64) if n else np.zeros((0, nv), np.uint64)   # (n, nv)
    Tpad = np.zeros((2 * npair, nv), np.uint64)
    Tpad[:n] = T
    G[0::2] = (w[0] < Tpad[0::2]).astype(np.int8) + (w[1] < Tpad[0::2]).astype(np.int8)
    G[1::2] = (w[2] < Tpad[1::2]).astype(np.int8) + (w[3] < Tpad[1::2]).astype(np.int8)
    return G[:n]


def np_synth_dense(seed, n, v0, nv, mode=0):
    G = np_synth_dosage(seed, n, v0, nv)
    return G if mode else (G > 0).astype(np.int8)


def dense_to_calls(X):
    """Dense multiplicity matrix (samples x variants) -> RDD[Seq[Int]] rows (ascending sample order,
    multiplicity m listed m times, empty variants dropped) as CSR (offsets int64, idx int32)."""
    X = np.asarray(X)
    rows = []
    for j in range(X.shape[1]):
        col = X[:, j]
        s = np.repeat(np.arange(X.shape[0]), col.astype(np.int64))
        if len(s):
            rows.append(s.astype(np.int32))
    off = np.zeros(len(rows) + 1, np.int64)
    if rows:
        off[1:] = np.cumsum([len(r) for r in rows])
        idx = np.concatenate(rows)
    else:
        idx = np.zeros(0, np.int32)
    return off, idx
