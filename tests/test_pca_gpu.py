"""GPU parity tests of centering + eigensolve (VariantsPca.scala:198-231) against the oracle, through the C ABI.

Tolerance (BASELINE.json north_star): top-k eigenvectors within 1e-6 relative, sign-normalised:
max_i |u^_i - u_i| / max_i |u_i| <= 1e-6 per component.  Centering is compared bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED = 20240901
TOL = 1e-6


def _native(n, **kw):
    from spark_examples_b200 import native
    return native.NativePca(n, **kw)


def _structured_gram(oracle, n, nv, seed=SEED):
    X = oracle.c_synth_dense(seed, n, 0, nv, 0)
    return oracle.np_similarity_dense(X)


@pytest.mark.parametrize("n,nv", [(64, 500), (257, 2000), (1092, 8000)])
def test_centering_bit_exact(oracle, n, nv):
    S = _structured_gram(oracle, n, nv)
    want, row_sums, nz = oracle.c_center(S)
    with _native(n) as nat:
        nat.setGram(S)
        C = nat.getCentered()
    assert np.array_equal(C, want)


@pytest.mark.parametrize("n,nv", [(2, 40), (3, 60), (50, 700), (257, 3000), (1092, 8000)])
def test_top2_eigenvectors_match_mllib_recipe(oracle, n, nv):
    S = _structured_gram(oracle, n, nv)
    want, sv = oracle.compute_pca(S, 2 if n > 2 else 1)
    k = want.shape[1]
    with _native(n) as nat:
        nat.setGram(S)
        vecs, evals, nz = nat.computePca(k)
        d, e = nat.getTridiagonal()
    C, rs, nz_want = oracle.np_center(S)
    assert nz == nz_want
    w = np.linalg.eigvalsh(C)[::-1]
    # eigenvalues of C (the reference's singular values of Cov are w^2/(n-1))
    assert np.allclose(evals, w[:k], rtol=1e-10, atol=1e-8 * abs(w[0]))
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    assert np.allclose(np.linalg.eigvalsh(T)[::-1][: min(n, 8)], w[: min(n, 8)], rtol=1e-9, atol=1e-8 * abs(w[0]))
    # unit norm + sign rule
    assert np.allclose(np.linalg.norm(vecs, axis=0), 1.0, atol=1e-12)
    for c in range(k):
        assert vecs[np.argmax(np.abs(vecs[:, c])), c] > 0
    err = oracle.eigvec_rel_err(vecs, want)
    assert np.all(err <= TOL), err


def test_top5_eigenvectors_random_psd(oracle):
    rng = np.random.default_rng(7)
    n = 300
    # integer Gram of a random low-rank-plus-noise design: well separated top-5
    B = rng.integers(0, 2, size=(n, 4000)).astype(np.int64)
    B[:60] |= rng.integers(0, 2, size=(1, 4000))
    B[60:150] |= rng.integers(0, 2, size=(1, 4000))
    B[150:210, ::3] = 1
    S = (B @ B.T).astype(np.int32)
    C, _, _ = oracle.np_center(S)
    w, V = np.linalg.eigh(C)
    with _native(n, num_pc=5) as nat:
        nat.setGram(S)
        vecs, evals, _ = nat.computePca(5)
    assert np.allclose(evals, w[::-1][:5], rtol=1e-10)
    gaps = np.abs(np.diff(w[::-1][:6]))
    assert gaps.min() > 1e-6 * w[-1]
    err = oracle.eigvec_rel_err(vecs, V[:, ::-1][:, :5])
    assert np.all(err <= TOL), err
    assert np.allclose(vecs.T @ vecs, np.eye(5), atol=1e-10)


def test_zero_gram_does_not_crash(oracle):
    n = 32
    with _native(n) as nat:
        nat.setGram(np.zeros((n, n), np.int32))
        vecs, evals, nz = nat.computePca(2)
    assert nz == 0
    assert np.all(np.isfinite(vecs)) and np.allclose(evals, 0.0)


def test_end_to_end_calls_to_pcs(oracle):
    """RDD[Seq[Int]] rows -> encode -> Gram -> PCA, all on device, vs the full oracle chain."""
    n, nv = 1092, 8000
    off, idx = oracle.c_synth_calls(SEED, n, 0, nv)
    S_want = oracle.c_similarity(n, off, idx, 4)
    want, _ = oracle.compute_pca(S_want, 2)
    with _native(n) as nat:
        nat.accumulateCalls(0, off, idx)
        nat.commit(0)
        nat.finalizeGram()
        vecs, evals, nz = nat.computePca(2)
    assert nz == n
    err = oracle.eigvec_rel_err(vecs, want)
    assert np.all(err <= TOL), err
    # README.md:109-119 property: unit-norm eigenvector entries of magnitude ~ 1/sqrt(N)
    assert 0.2 / np.sqrt(n) < np.abs(vecs[:, 0]).mean() < 2.0 / np.sqrt(n)
