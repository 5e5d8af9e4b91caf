"""CPU tests of the oracle itself (no GPU): hand-computed cases, the two independent restatements against each
other, the committed golden vectors, and the reference's own known-answer properties (README.md:109-119)."""
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).resolve().parent / "golden"
SEED = 20240901


def test_philox_known_answers(oracle):
    """Random123 known-answer vectors for Philox4x32-10."""
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
         [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ]
    for ctr, key, want in kat:
        assert [int(x) for x in oracle.c_philox(ctr, key)] == want
        got = oracle.np_philox4x32_10(*ctr, *key)
        assert [int(x) for x in got] == want


def test_extract_call_info_rule(oracle):
    """VariantsPca.scala:58: variation = any allele index > 0; -1 (no-call) and 0 (reference) are not."""
    mapping = {"a": 0, "b": 1, "c": 2, "d": 3}
    calls = [("a", [0, 0]), ("b", [0, 1]), ("c", [-1, -1]), ("d", [2]), ("a", [])]
    info = oracle.np_extract_call_info(calls, mapping)
    assert info == [(False, 0), (True, 1), (False, 2), (True, 3), (False, 0)]
    assert oracle.np_extract_call_info(None, mapping) == []
    with pytest.raises(KeyError):
        oracle.np_extract_call_info([("zz", [1])], mapping)


def test_get_calls_drops_empty_variants_and_keeps_duplicates(oracle):
    mapping = {"a": 0, "b": 1, "c": 2}
    variants = [
        [("a", [0, 1]), ("b", [0, 0]), ("c", [1, 1])],     # -> [0, 2]
        [("a", [0, 0]), ("b", [-1, 0])],                    # no variation -> dropped (:166)
        None,                                               # calls = None -> dropped
        [("b", [1]), ("b", [2])],                           # same callset twice -> [1, 1]
    ]
    assert oracle.np_get_calls(variants, mapping) == [[0, 2], [1, 1]]
    # columnar C restatement of the same thing
    call_off = [0, 3, 5, 5, 7]
    callset = [0, 1, 2, 0, 1, 1, 1]
    gt = [[0, 1], [0, 0], [1, 1], [0, 0], [-1, 0], [1], [2]]
    gt_off = np.concatenate([[0], np.cumsum([len(g) for g in gt])])
    off, idx = oracle.c_encode_calls(3, call_off, callset, gt_off, np.concatenate(gt))
    assert off.tolist() == [0, 2, 4] and idx.tolist() == [0, 2, 1, 1]
    with pytest.raises(IndexError):
        oracle.c_encode_calls(2, [0, 1], [2], [0, 1], [1])


def test_similarity_hand_computed(oracle):
    """Rows {0,2}, {1,1} (sample 1 listed twice), {0,1,2}: full cartesian square per row (:186-188)."""
    rows = [[0, 2], [1, 1], [0, 1, 2]]
    want = np.array([[2, 1, 2],
                     [1, 5, 1],
                     [2, 1, 2]], np.int32)
    off = np.array([0, 2, 4, 7], np.int64)
    idx = np.array([0, 2, 1, 1, 0, 1, 2], np.int32)
    assert np.array_equal(oracle.np_similarity(3, rows), want)
    for parts in (1, 2, 3):
        assert np.array_equal(oracle.c_similarity(3, off, idx, parts), want)
    X = np.array([[1, 0, 1], [0, 2, 1], [1, 0, 1]])
    assert np.array_equal(oracle.np_similarity_dense(X), want)
    with pytest.raises(IndexError):
        oracle.c_similarity(3, np.array([0, 1], np.int64), np.array([3], np.int32))


def test_similarity_stream_same_matrix(oracle):
    off, idx = oracle.c_synth_calls(SEED, 60, 0, 200)
    assert np.array_equal(oracle.c_similarity_stream(60, off, idx), oracle.c_similarity(60, off, idx, 3))


def test_centering_hand_computed(oracle):
    S = np.array([[2, 1, 2], [1, 5, 1], [2, 1, 2]], np.int32)
    C, rs, nz = oracle.np_center(S)
    assert rs.tolist() == [5.0, 7.0, 5.0] and nz == 3
    mm = 17.0 / 3 / 3
    want = np.array([[S[i, j] - rs[i] / 3 - rs[j] / 3 + mm for j in range(3)] for i in range(3)])
    assert np.array_equal(C, want)
    Cc, rsc, nzc = oracle.c_center(S)
    assert np.array_equal(Cc, C) and np.array_equal(rsc, rs) and nzc == nz
    assert np.allclose(C.sum(axis=0), 0, atol=1e-12) and np.allclose(C.sum(axis=1), 0, atol=1e-12)


def test_c_and_numpy_restatements_agree(oracle):
    n, nv = 97, 400
    Xc = oracle.c_synth_dense(SEED, n, 5, nv)
    Xn = oracle.np_synth_dense(SEED, n, 5, nv)
    assert np.array_equal(Xc, Xn)
    assert np.array_equal(oracle.c_synth_dense(SEED, n, 5, nv, 1), oracle.np_synth_dense(SEED, n, 5, nv, 1))
    off, idx = oracle.c_synth_calls(SEED, n, 5, nv)
    off2, idx2 = oracle.dense_to_calls(Xn)
    assert np.array_equal(off, off2) and np.array_equal(idx, idx2)
    S1 = oracle.c_similarity(n, off, idx, 4)
    S2 = oracle.np_similarity(n, [idx[off[v]:off[v + 1]] for v in range(len(off) - 1)])
    assert np.array_equal(S1, S2) and np.array_equal(S1, oracle.np_similarity_dense(Xn))
    C1, _, _ = oracle.c_center(S1)
    C2, _, _ = oracle.np_center(S1)
    assert np.array_equal(C1, C2)
    assert np.allclose(oracle.c_mllib_covariance(C1), (C1.T @ C1) / (n - 1) - n / (n - 1) * np.outer(C1.mean(0), C1.mean(0)),
                       rtol=1e-10, atol=1e-6)


def test_mllib_recipe_equals_eigh_of_centered(oracle):
    """C = J S J is PSD, so svd(Cov)[:, :k] are the top-k eigenvectors of C."""
    n, nv = 150, 1200
    S = oracle.np_similarity_dense(oracle.np_synth_dense(SEED, n, 0, nv))
    C, _, _ = oracle.np_center(S)
    U, sv = oracle.mllib_principal_components(C, 3)
    w, V = np.linalg.eigh(C)
    assert w[0] > -1e-8 * w[-1]
    assert np.all(oracle.eigvec_rel_err(U, V[:, ::-1][:, :3]) < 1e-9)
    assert np.allclose(sv, w[::-1][:3] ** 2 / (n - 1), rtol=1e-9)
    with pytest.raises(ValueError):                       # upstream RowMatrix refuses more than 65535 columns
        oracle.mllib_principal_components(np.empty((65536, 0)), 1)


def test_mllib_recipe_known_answer_of_upstream_pca_test(oracle):
    """The 4 x 3 matrix and the expected components of the "pca" test in upstream Spark's RowMatrixSuite
    (mllib/src/test/scala/org/apache/spark/mllib/linalg/distributed/RowMatrixSuite.scala -- spark-mllib is the
    un-vendored dependency behind VariantsPca.scala:225-226; the suite is not fetchable here, so the numbers are restated
    and re-derived: Cov = [[15, 0, 0], [0, 10, 10], [0, 10, 10]] exactly, eigenvalues 20, 15, 0).  Upstream compares
    columns up to sign (`assertColumnEqualUpToSign`)."""
    rows = np.array([[0.0, 1.0, 2.0], [3.0, 4.0, 5.0], [6.0, 7.0, 8.0], [9.0, 0.0, 1.0]])
    want = np.array([[0.0, 1.0, 0.0],
                     [np.sqrt(2.0) / 2.0, 0.0, np.sqrt(2.0) / 2.0],
                     [np.sqrt(2.0) / 2.0, 0.0, -np.sqrt(2.0) / 2.0]])
    m, mu = 4.0, rows.mean(axis=0)
    cov = rows.T @ rows / (m - 1.0) - m / (m - 1.0) * np.outer(mu, mu)       # computeCovariance, as oracle.py restates it
    assert np.allclose(cov, [[15, 0, 0], [0, 10, 10], [0, 10, 10]], atol=1e-12)
    U, sv = oracle.mllib_principal_components(rows, 3)
    assert np.allclose(sv, [20.0, 15.0, 0.0], atol=1e-12)
    assert np.allclose(sv[:2] / sv.sum(), [4.0 / 7.0, 3.0 / 7.0])          # explainedVariance of the later suite versions
    for c in range(3):
        assert np.allclose(U[:, c], want[:, c], atol=1e-12) or np.allclose(U[:, c], -want[:, c], atol=1e-12)


@pytest.mark.parametrize("name", ["cohort_n48_v200", "cohort_n200_v1500"])
def test_golden_cohorts(oracle, name):
    g = np.load(GOLD / f"{name}.npz")
    n, nv = int(g["n"]), int(g["nv"])
    X = np.unpackbits(g["X_packed"], axis=1)[:, :nv].astype(np.int8)
    assert np.array_equal(oracle.c_synth_dense(int(g["seed"]), n, 0, nv), X)
    off, idx = oracle.c_synth_calls(int(g["seed"]), n, 0, nv)
    assert np.array_equal(off, g["offsets"]) and np.array_equal(idx, g["idx"])
    S = oracle.c_similarity(n, off, idx, 2)
    assert np.array_equal(S, g["S"])
    C, rs, nz = oracle.c_center(S)
    assert np.array_equal(rs, g["row_sums"]) and nz == int(g["non_zero_rows"])
    assert np.allclose([C.sum(), np.abs(C).sum(), C[0, 0], C[n // 2, n // 3]], g["C_checksum"], rtol=1e-12, atol=1e-9)
    U, sv = oracle.mllib_principal_components(C, 2)
    assert np.all(oracle.eigvec_rel_err(U, g["U"]) < 1e-9)
    assert np.allclose(sv, g["cov_singular_values"], rtol=1e-9)


def test_golden_generator(oracle):
    g = np.load(GOLD / "generator.npz")
    for v, thr in zip(g["variants"], g["thresholds"]):
        assert np.array_equal(oracle.c_variant_thresholds(int(g["seed"]), int(v)), thr)
    assert np.array_equal(oracle.c_synth_dense(int(g["seed"]), 37, 1000, 53, 1), g["dosage_n37_v1000_53"])
    assert oracle.np_pop_bounds(2504).tolist() == g["pop_bounds_2504"].tolist() == [651, 1001, 1502, 2003, 2504]


def test_readme_known_answer_properties(oracle):
    """README.md:109-119: outputs are unit-norm eigenvector entries (|v| ~ 1/sqrt(1092) = 0.0303), printed as
    name<TAB>dataset<TAB>pc1<TAB>pc2 sorted by name (VariantsPca.scala:238-239)."""
    readme = [("NA20811", 0.0286308791579312, -0.008456233951873527), ("NA20818", -0.033609576645005836, -0.026655905606186293)]
    for _, pc1, _ in readme:
        assert 0.5 / np.sqrt(1092) < abs(pc1) < 1.5 / np.sqrt(1092)
    n, nv = 1092, 3000
    S = oracle.np_similarity_dense(oracle.c_synth_dense(SEED, n, 0, nv))
    U, _ = oracle.compute_pca(S, 2)
    assert np.allclose(np.linalg.norm(U, axis=0), 1.0, atol=1e-12)
    assert 0.5 / np.sqrt(n) < np.abs(U[:, 0]).mean() < 1.5 / np.sqrt(n)
    names = {f"synth-{i}": nm for i, nm in enumerate(["NA20818", "NA20811"])}
    lines = oracle.emit_result_lines([("synth-0", readme[1][1], readme[1][2]), ("synth-1", readme[0][1], readme[0][2])], names)
    assert lines == ["NA20811\tsynth\t0.0286308791579312\t-0.008456233951873527",
                     "NA20818\tsynth\t-0.033609576645005836\t-0.026655905606186293"]


def test_synthetic_cohort_has_separated_structure(oracle):
    """unequal drift plants well separated top eigenvalues (otherwise 1e-6 eigenvector parity is moot)."""
    n, nv = 400, 4000
    X = oracle.c_synth_dense(SEED, n, 0, nv)
    assert 0.35 < X.mean() < 0.5
    C, _, _ = oracle.np_center(oracle.np_similarity_dense(X))
    w = np.linalg.eigvalsh(C)[::-1]
    assert (w[0] - w[1]) / w[0] > 0.05 and (w[1] - w[2]) / w[1] > 0.03 and w[3] / w[4] > 2


def test_murmur3_128_known_answers_and_variant_key(oracle):
    """Guava `Hashing.murmur3_128()` (un-vendored dependency, build.sbt:44) known answers, printed like HashCode.toString;
    the oracle's restatement and the product's host mirror agree on random byte strings of every tail length."""
    from spark_examples_b200.variants_pca import murmur3_128 as product_murmur, getVariantKey
    import spark_examples_b200 as pkg
    assert oracle.np_murmur3_128(b"").hex() == "0" * 32
    assert oracle.np_murmur3_128(b"hello").hex() == "029bbd41b3a7d8cb191dae486a901e5b"
    assert oracle.np_murmur3_128(b"The quick brown fox jumps over the lazy dog").hex() == "6c1b07bc7bbc4be347939ac4a93c437a"
    rng = np.random.default_rng(8)
    for l in list(range(0, 50)) + [63, 64, 65, 200]:
        data = bytes(rng.integers(0, 256, l, dtype=np.uint8))
        assert oracle.np_murmur3_128(data).hex() == product_murmur(data)
    v = pkg.Variant("17", start=41196311, end=41196312, referenceBases="A", alternateBases=["C", "T"])
    assert oracle.np_variant_key("17", 41196311, 41196312, "A", ["C", "T"]) == getVariantKey(v)
    assert oracle.np_variant_key("17", 41196311, 41196312, "A", ["CT"]) == getVariantKey(v)        # mkString("") (:63-64)
    assert oracle.np_variant_key("17", 41196311, 41196313, "A", ["CT"]) != getVariantKey(v)


def test_join_and_merge_restatements_on_hand_cases(oracle):
    """:115-128 inner join with duplicates on both sides (cartesian product per key); :136-148 merge keeps groups of
    exactly variantSetCount records, whatever dataset they come from."""
    left = [("k1", [0]), ("k2", [1, 2]), ("k1", [3])]
    right = [("k1", [7]), ("k3", [8]), ("k1", [])]
    assert oracle.np_join_datasets(left, right) == [[0, 7], [0], [3, 7], [3]]
    ds = [[("a", [0]), ("b", [1])], [("a", [2]), ("c", [3])], [("a", [4]), ("b", [5]), ("b", [6])]]
    assert oracle.np_merge_datasets(ds, 3) == [[0, 2, 4], [1, 5, 6]]       # "b": three records although only two datasets have it
    assert oracle.np_merge_datasets(ds, 1) == [[3]]
