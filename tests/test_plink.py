"""PLINK .bed ingestion (spark_examples_b200/plink.py): host-side reader / decoder on the CPU, the packed-row device
path (vpca_accumulate_bed) on the GPU against the oracle's Gram of the decoded calls."""
import numpy as np
import pytest

from spark_examples_b200 import plink


def _genotypes(n, v, seed, missing=0.03):
    rng = np.random.default_rng(seed)
    p = rng.uniform(0.02, 0.5, size=v)
    d = rng.binomial(2, p[None, :], size=(n, v)).astype(np.int64)         # A1 allele counts
    d[rng.random((n, v)) < missing] = -1
    return d


@pytest.mark.parametrize("n", [1, 4, 13, 64, 1092])
def test_fileset_round_trip_and_decode_rule(tmp_path, n):
    v = 257
    d = _genotypes(n, v, n)
    plink.write_fileset(str(tmp_path / "c"), d)
    fam, bim = plink.read_fam(str(tmp_path / "c.bed")), plink.read_bim(str(tmp_path / "c"))
    assert len(fam) == n and fam[0] == ("synth-S000000", "S000000") and len(bim) == v and bim[3].position == 41196314
    bed = plink.BedFile(str(tmp_path / "c"), n_variants=len(bim))
    assert (bed.n_samples, bed.n_variants, bed.stride) == (n, v, (n + 3) // 4)
    rows = bed.rows(0, v)
    # hasVariation = genotype.foldLeft(false)(_ || _ > 0) with A1 as the alternate allele; missing = no-call = False
    assert np.array_equal(plink.decode_rows(rows, n, plink.COUNT_A1), d.T > 0)
    assert np.array_equal(plink.decode_rows(rows, n, plink.COUNT_A2), (d.T == 0) | (d.T == 1))
    off, idx = plink.rows_to_calls(rows[10:50], n, plink.COUNT_A1)
    want = [np.nonzero(d[:, j] > 0)[0] for j in range(10, 50)]
    want = [w for w in want if len(w)]                                     # VariantsPca.scala:166
    assert len(off) - 1 == len(want)
    for r, w in enumerate(want):
        assert np.array_equal(idx[off[r]:off[r + 1]], w)


def test_bad_files_are_rejected(tmp_path):
    d = _genotypes(5, 9, 0)
    plink.write_fileset(str(tmp_path / "c"), d)
    raw = (tmp_path / "c.bed").read_bytes()
    (tmp_path / "s.bed").write_bytes(raw[:2] + b"\x00" + raw[3:])          # sample-major flag
    (tmp_path / "s.fam").write_text((tmp_path / "c.fam").read_text())
    with pytest.raises(ValueError, match="sample-major"):
        plink.BedFile(str(tmp_path / "s"))
    (tmp_path / "t.bed").write_bytes(raw[:-1])                             # truncated
    (tmp_path / "t.fam").write_text((tmp_path / "c.fam").read_text())
    with pytest.raises(ValueError, match="does not match"):
        plink.BedFile(str(tmp_path / "t"))
    (tmp_path / "m.bed").write_bytes(b"xyz" + raw[3:])
    (tmp_path / "m.fam").write_text((tmp_path / "c.fam").read_text())
    with pytest.raises(ValueError, match="not a PLINK"):
        plink.BedFile(str(tmp_path / "m"))


def test_driver_source_and_flags(tmp_path, capsys):
    """--bed-path builds BedSlice partitions; CallsRdd.collect() is the decoded RDD[Seq[Int]]."""
    from spark_examples_b200.conf import PcaConf
    from spark_examples_b200.variants_common import BedSlice, VariantsCommon
    from spark_examples_b200.variants_pca import CallsRdd
    d = _genotypes(21, 100, 3)
    plink.write_fileset(str(tmp_path / "c"), d, fam=[(f"F{i % 3}", f"I{i}") for i in range(21)])
    conf = PcaConf(["--bed-path", str(tmp_path / "c"), "--variants-per-partition", "32", "--bed-counted-allele", "A2"])
    common = VariantsCommon(conf)
    assert "Matrix size: 21." in capsys.readouterr().out
    assert list(common.indexes)[:2] == ["F0-I0", "F1-I1"] and common.names["F1-I1"] == "I1"
    parts = common.data[0].partitions
    assert [type(p) for p in parts] == [BedSlice] * 4 and [p.nv for p in parts] == [32, 32, 32, 4]
    rows = CallsRdd(parts, 21).collect()
    want = [np.nonzero((d[:, j] == 0) | (d[:, j] == 1))[0].tolist() for j in range(100)]
    assert rows == [w for w in want if w]


@pytest.mark.gpu
@pytest.mark.parametrize("n,nv,counted,dtype", [(5, 40, 1, "i8"), (257, 3000, 1, "i8"), (1092, 9000, 2, "i8"),
                                                (1092, 9000, 1, "e2m1"), (2504, 20000, 1, "i8")])
def test_accumulate_bed_matches_oracle(oracle, n, nv, counted, dtype):
    from spark_examples_b200 import native
    d = _genotypes(n, nv, n + nv)
    code = np.full((nv, n), 1, np.uint8)
    code[d.T == 2], code[d.T == 1], code[d.T == 0] = 0, 2, 3
    pad = (-n) % 4
    c4 = np.concatenate([code, np.zeros((nv, pad), np.uint8)], axis=1).reshape(nv, -1, 4)
    rows = (c4[:, :, 0] | (c4[:, :, 1] << 2) | (c4[:, :, 2] << 4) | (c4[:, :, 3] << 6)).astype(np.uint8)
    off, idx = plink.rows_to_calls(rows, n, counted)
    want = oracle.c_similarity(n, off, idx, 4)
    dt = {"i8": native.DTYPE_I8, "e2m1": native.DTYPE_E2M1}[dtype]
    with native.NativePca(n, dtype=dt, max_multiplicity=1) as nat:
        half = nv // 2
        nat.accumulateBed(3, rows[:half], counted)             # staged partition
        nat.commit(3)
        nat.accumulateBed(-1, rows[half:], counted)            # straight into the Gram
        nat.finalizeGram()
        got = nat.getGram()
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_driver_main_on_a_bed_fileset(tmp_path, oracle, capsys):
    """python -m spark_examples_b200 --bed-path: same PCs as the oracle chain on the decoded calls."""
    from spark_examples_b200.conf import PcaConf
    from spark_examples_b200.variants_pca import VariantsPcaDriver
    n, nv = 300, 4000
    X = oracle.c_synth_dense(20240901, n, 0, nv, 1)                        # dosage 0/1/2 with population structure
    plink.write_fileset(str(tmp_path / "c"), X.astype(np.int64))
    conf = PcaConf(["--bed-path", str(tmp_path / "c"), "--variants-per-partition", "1500"])
    driver = VariantsPcaDriver(conf)
    try:
        calls = driver.getCallsRdd(driver.getData)
        sim = driver.getSimilarityMatrix(calls)
        result = driver.computePca(sim)
    finally:
        driver.stop()
    S = oracle.np_similarity_dense((X > 0).astype(np.uint8))
    want, _ = oracle.compute_pca(S, 2)
    got = np.array([[r[1], r[2]] for r in result])
    # rows come back in callset-index order, keyed by callset id (VariantsPca.scala:228-230)
    assert [r[0] for r in result][:2] == ["synth-S000000", "synth-S000001"]
    assert np.all(oracle.eigvec_rel_err(got, want) <= 1e-6)
