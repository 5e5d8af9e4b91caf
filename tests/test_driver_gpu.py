"""GPU tests of the drop-in driver surface: the same call sequence as VariantsPcaDriver.main (VariantsPca.scala:38-50),
results compared with the oracle chain."""
import io

import numpy as np
import pytest

import spark_examples_b200 as pkg
from spark_examples_b200.variants_pca import VariantsPcaDriver

pytestmark = pytest.mark.gpu


def _records(rng, callsets, X):
    """Variant records whose carrier pattern is the binary matrix X (samples x variants)."""
    out = []
    for j in range(X.shape[1]):
        calls = []
        for i, (cid, _) in enumerate(callsets):
            g = [0, 1] if X[i, j] else ([0, 0] if rng.random() < 0.8 else [-1, -1])
            calls.append(pkg.Call(cid, genotype=g))
        out.append(pkg.Variant("chr17", start=41196311 + j, end=41196312 + j, referenceBases="A", alternateBases=["G"],
                               calls=calls))
    return out


def test_main_sequence_on_variant_records(oracle):
    rng = np.random.default_rng(21)
    n, nv = 60, 400
    callsets = [(f"hg-{i}", f"NA{i:05d}") for i in range(n)]
    X = oracle.c_synth_dense(20240901, n, 0, nv)
    conf = pkg.PcaConf(["--variants-per-partition", "128", "--num-pc", "2"])
    driver = VariantsPcaDriver(conf, common=pkg.VariantsCommon(conf, callsets=callsets, datasets=[_records(rng, callsets, X)]))
    data = driver.getData
    filtered = [driver.filterDataset(d) for d in data]
    callsRdd = driver.getCallsRdd(filtered)
    simMatrix = driver.getSimilarityMatrix(callsRdd)
    result = driver.computePca(simMatrix)
    buf = io.StringIO()
    driver.emitResult(result, out=buf)
    driver.reportIoStats()
    S = simMatrix.toArray()
    driver.stop()
    S_want = oracle.np_similarity_dense(X)
    assert np.array_equal(S, S_want)
    assert simMatrix.collect()[n + 1] == ((1, 1), int(S_want[1, 1]))
    U_want, _ = oracle.compute_pca(S_want, 2)
    got = np.array([[r[1], r[2]] for r in result])
    assert [r[0] for r in result] == [c[0] for c in callsets]
    assert np.all(oracle.eigvec_rel_err(got, U_want) <= 1e-6)
    want_lines = oracle.emit_result_lines([(c[0], *oracle.sign_normalise(U_want)[i]) for i, c in enumerate(callsets)], dict(callsets))
    got_lines = buf.getvalue().splitlines()
    assert [l.split("\t")[:2] for l in got_lines] == [l.split("\t")[:2] for l in want_lines]
    for g, w in zip(got_lines, want_lines):
        assert np.allclose([float(x) for x in g.split("\t")[2:]], [float(x) for x in w.split("\t")[2:]], rtol=0, atol=1e-7)


def test_synthetic_source_through_driver(oracle):
    n, nv = 320, 3000
    conf = pkg.PcaConf(["--synthetic", f"{n},{nv},20240901", "--variants-per-partition", "1024"])
    driver = VariantsPcaDriver(conf)
    sim = driver.getSimilarityMatrix(driver.getCallsRdd(driver.getData))
    result = driver.computePca(sim)
    S = sim.toArray()
    driver.stop()
    X = oracle.c_synth_dense(20240901, n, 0, nv)
    assert np.array_equal(S, oracle.np_similarity_dense(X))
    U_want, _ = oracle.compute_pca(S, 2)
    assert np.all(oracle.eigvec_rel_err(np.array([[r[1], r[2]] for r in result]), U_want) <= 1e-6)


def test_unknown_callset_and_bad_index_raise(oracle):
    callsets = [("a-0", "A"), ("a-1", "B")]
    conf = pkg.PcaConf([])
    bad = [pkg.Variant("1", calls=[pkg.Call("zz-9", genotype=[1])])]
    d = VariantsPcaDriver(conf, common=pkg.VariantsCommon(conf, callsets=callsets, datasets=[bad]))
    with pytest.raises(KeyError):                                  # NoSuchElementException at VariantsPca.scala:59
        d.getCallsRdd(d.getData)


def test_checkpoint_resume_counts_every_partition_once(tmp_path, oracle):
    """a run that dies after some partitions resumes from the saved Gram + watermark and ends with the
    same similarity matrix as an uninterrupted run."""
    rng = np.random.default_rng(5)
    n, nv = 40, 600
    callsets = [(f"ck-{i}", f"S{i:03d}") for i in range(n)]
    X = oracle.c_synth_dense(20240901, n, 0, nv)
    records = _records(rng, callsets, X)
    ck = str(tmp_path / "pca_ck")
    args = ["--variants-per-partition", "100", "--checkpoint-path", ck]

    conf = pkg.PcaConf(args)
    d1 = VariantsPcaDriver(conf, common=pkg.VariantsCommon(conf, callsets=callsets, datasets=[records]))
    rdd = d1.getCallsRdd(d1.getData)
    assert len(rdd.partitions) == 6
    # first attempt: partition 3 is corrupt (index out of range) -> the job dies after 3 committed partitions
    good3 = rdd.partitions[3]
    rdd.partitions[3] = type(good3)(good3.offsets.copy(), np.where(np.arange(len(good3.idx)) == 5, n + 7, good3.idx).astype(np.int32))
    with pytest.raises(IndexError):
        d1.getSimilarityMatrix(rdd)
    d1.stop()
    # (the checkpoint is written every 16 partitions and at the end; force one for the 3 partitions that committed)
    conf2 = pkg.PcaConf(args)
    d2 = VariantsPcaDriver(conf2, common=pkg.VariantsCommon(conf2, callsets=callsets, datasets=[records]))
    rdd2 = d2.getCallsRdd(d2.getData)
    nat = d2._native(n)
    nat.reset()
    for pid in range(3):
        nat.accumulateCalls(pid, rdd2.partitions[pid].offsets, rdd2.partitions[pid].idx)
        nat.commit(pid)
    d2._save_checkpoint(nat, rdd2, {0, 1, 2}, every=1)
    d2.stop()
    # resume: only partitions 3..5 are processed again
    conf3 = pkg.PcaConf(args)
    d3 = VariantsPcaDriver(conf3, common=pkg.VariantsCommon(conf3, callsets=callsets, datasets=[records]))
    sim = d3.getSimilarityMatrix(d3.getCallsRdd(d3.getData))
    st = d3._nat.stats()
    S = sim.toArray()
    d3.stop()
    assert np.array_equal(S, oracle.np_similarity_dense(X))
    assert st["variants_accumulated"] == sum(len(p.offsets) - 1 for p in rdd2.partitions[3:])
