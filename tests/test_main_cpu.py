"""`VariantsPcaDriver.main` (VariantsPca.scala:38-50) from every offline source, on the CPU: the driver's plumbing --
which rows reach `accumulateCalls` / `accumulateBed`, exactly-once commits, callset order, output format -- with the
native library replaced by a TEST DOUBLE that computes through the oracle.  The double lives here only; the product has
no CPU path (tests/test_abi.py), and the same flows run against libvpca.so in the `-m gpu` tests."""
import io

import numpy as np
import pytest

import spark_examples_b200 as pkg
from spark_examples_b200 import parquet_calls, plink, vcf
from spark_examples_b200 import variants_pca
from spark_examples_b200.variants_pca import VariantsPcaDriver


class OracleBackedNative:
    """Same method surface as native.NativePca for the calls the driver makes."""

    def __init__(self, oracle, n):
        self.o, self.n = oracle, n
        self.S = np.zeros((n, n), np.int64)
        self.staged = {}
        self.log = []

    def reset(self):
        self.S[:] = 0
        self.staged.clear()

    def accumulateCalls(self, pid, off, idx):
        assert pid not in self.staged, "one accumulate per partition in this driver"
        self.log.append(("calls", pid, len(off) - 1))
        self.staged[pid] = self.o.c_similarity(self.n, np.asarray(off, np.int64), np.asarray(idx, np.int32), 1)

    def accumulateBed(self, pid, rows, counted):
        off, idx = plink.rows_to_calls(rows, self.n, counted)
        self.log.append(("bed", pid, rows.shape[0]))
        self.staged[pid] = self.o.c_similarity(self.n, off, idx, 1)

    def joinRows(self, mode, keys, offsets, idx, n_left=0, variant_set_count=2):
        from spark_examples_b200.variants_common import JoinedSlice
        self.joined = variants_pca.joined_rows_on_host(JoinedSlice(mode, keys, offsets, idx, n_left, variant_set_count))
        self.log.append(("join", mode, len(keys)))
        return len(self.joined.offsets) - 1, int(self.joined.offsets[-1])

    def accumulateJoined(self, pid):
        self.log.append(("joined", pid, len(self.joined.offsets) - 1))
        self.staged[pid] = self.o.c_similarity(self.n, self.joined.offsets, self.joined.idx, 1)

    def commit(self, pid):
        self.S += self.staged.pop(pid)

    def abort(self, pid):
        self.staged.pop(pid, None)

    def finalizeGram(self):
        assert not self.staged, "every partition must be committed or aborted before finalize"

    def getGram(self):
        return self.S.astype(np.int32)

    def computePca(self, k):
        U, sv = self.o.compute_pca(self.getGram(), k)
        C, rs, nz = self.o.np_center(self.getGram())
        return self.o.sign_normalise(U), sv, nz

    def stats(self):
        return {"variants_accumulated": 0, "gram_launches": 0, "kernel_launches": 0, "h2d_bytes": 0, "last_gram_ms": 0.0,
                "last_eig_ms": 0.0}

    def close(self):
        pass


@pytest.fixture
def fake_native(monkeypatch, oracle):
    made = []

    def _native(self, n):
        if self._nat is None:
            self._nat = OracleBackedNative(oracle, n)
            made.append(self._nat)
        return self._nat
    monkeypatch.setattr(VariantsPcaDriver, "_native", _native)
    return made


def _cohort(oracle, n=60, nv=400):
    return oracle.c_synth_dense(20240901, n, 0, nv, 1).astype(np.int64)          # dosage 0/1/2, population structure


def _expected_lines(oracle, X_has, names, datasets):
    S = oracle.np_similarity_dense(X_has.astype(np.uint8))
    U, _ = oracle.compute_pca(S, 2)
    U = oracle.sign_normalise(U)
    rows = sorted(zip(names, datasets, U[:, 0], U[:, 1]), key=lambda t: t[0])
    return [f"{nm}\t{ds}\t{pkg.jformat.jdouble(a)}\t{pkg.jformat.jdouble(b)}" for nm, ds, a, b in rows]


def _run_main(argv, capsys):
    variants_pca.main(argv)
    out = capsys.readouterr().out.splitlines()
    return out, [ln for ln in out if ln.count("\t") == 3]


def test_main_from_vcf(tmp_path, capsys, oracle, fake_native):
    d = _cohort(oracle)
    n, nv = d.shape
    samples = [f"NA{i:05d}" for i in range(n)]
    gt = {0: "0/0", 1: "0/1", 2: "1/1"}
    recs = [dict(chrom="chr17", pos=1000 + j, ref="A", alt=["C"], info={"AF": [0.3]}, gts=[gt[int(x)] for x in d[:, j]])
            for j in range(nv)]
    path = str(tmp_path / "trial.vcf.gz")
    vcf.write_vcf(path, samples, recs)
    out, lines = _run_main(["--vcf-path", path, "--variants-per-partition", "150", "--output-path", str(tmp_path / "o")],
                           capsys)
    assert "Matrix size: 60." in out and any(ln.startswith("Non zero rows in matrix: 60 / 60.") for ln in out)
    assert lines == _expected_lines(oracle, d > 0, samples, ["trial"] * n)
    assert [e[0] for e in fake_native[0].log] == ["calls"] * 3                   # 400 records in partitions of 150
    part = (tmp_path / "o-pca.tsv" / "part-00000").read_text().splitlines()      # saveAsTextFile layout (:241-245)
    assert len(part) == n and part[0].split("\t")[0] == "NA00000" and part[0].split("\t")[3] == "trial"


def test_main_from_two_vcf_files_joins_on_the_variant_key(tmp_path, capsys, oracle, fake_native):
    """Two variant sets (--vcf-path a,b): main takes the join branch of getCallsRdd (VariantsPca.scala:159); the rows of
    both files and their key bytes go to the native join in ONE call, the joined rows are accumulated as one partition."""
    d = _cohort(oracle, n=50, nv=300)
    na = 30
    gt = {0: "0/0", 1: "0/1", 2: "1/1"}
    keep_a = np.arange(300) % 3 != 0                    # each file misses a third of the sites: the join keeps the overlap
    keep_b = np.arange(300) % 3 != 1
    def recs(rows, keep):
        return [dict(chrom="chr2", pos=500 + j, ref="G", alt=["T"], gts=[gt[int(x)] for x in d[rows, j]]) for j in range(300) if keep[j]]
    pa, pb = str(tmp_path / "setA.vcf"), str(tmp_path / "setB.vcf")
    sa, sb = [f"A{i:03d}" for i in range(na)], [f"B{i:03d}" for i in range(50 - na)]
    vcf.write_vcf(pa, sa, recs(slice(0, na), keep_a))
    vcf.write_vcf(pb, sb, recs(slice(na, 50), keep_b))
    out, lines = _run_main(["--vcf-path", f"{pa},{pb}"], capsys)
    both = keep_a & keep_b
    assert lines == _expected_lines(oracle, (d > 0)[:, both], sa + sb, ["setA"] * na + ["setB"] * (50 - na))
    log = fake_native[0].log
    assert [e[0] for e in log] == ["join", "joined"] and log[0][2] == int(keep_a.sum() + keep_b.sum())


def test_main_from_bed(tmp_path, capsys, oracle, fake_native):
    d = _cohort(oracle, 50, 300)
    n, nv = d.shape
    plink.write_fileset(str(tmp_path / "c"), d, fam=[(f"fam{i % 2}", f"I{i:03d}") for i in range(n)])
    out, lines = _run_main(["--bed-path", str(tmp_path / "c"), "--variants-per-partition", "128"], capsys)
    assert lines == _expected_lines(oracle, d > 0, [f"I{i:03d}" for i in range(n)], [f"fam{i % 2}" for i in range(n)])
    assert fake_native[0].log == [("bed", 0, 128), ("bed", 1, 128), ("bed", 2, 44)]


def test_main_from_calls_parquet(tmp_path, capsys, oracle, fake_native):
    d = _cohort(oracle, 40, 500)
    n, nv = d.shape
    has = d.T > 0
    counts = has.sum(axis=1)
    off = np.zeros(nv + 1, np.int64)
    np.cumsum(counts, out=off[1:])
    idx = np.nonzero(has)[1].astype(np.int32)
    callsets = [(f"pq-{i}", f"P{i:02d}") for i in range(n)]
    path = str(tmp_path / "calls.parquet")
    parquet_calls.write_calls(path, callsets, off, idx, row_group_variants=200)
    out, lines = _run_main(["--calls-parquet-path", path], capsys)
    assert lines == _expected_lines(oracle, d > 0, [c[1] for c in callsets], ["pq"] * n)
    assert [e[:2] for e in fake_native[0].log] == [("calls", 0), ("calls", 1), ("calls", 2)]


def test_failed_partition_is_aborted_and_nothing_is_half_applied(tmp_path, oracle, fake_native):
    d = _cohort(oracle, 30, 100)
    plink.write_fileset(str(tmp_path / "c"), d)
    driver = VariantsPcaDriver(pkg.PcaConf(["--bed-path", str(tmp_path / "c"), "--variants-per-partition", "40"]))
    calls = driver.getCallsRdd(driver.getData)
    nat = driver._native(30)
    real = nat.accumulateBed

    def flaky(pid, rows, counted):
        if pid == 1:
            raise RuntimeError("task failed")
        return real(pid, rows, counted)
    nat.accumulateBed = flaky
    with pytest.raises(RuntimeError, match="task failed"):
        driver.getSimilarityMatrix(calls)
    assert not nat.staged                                                        # partition 1 left nothing behind
    off, idx = plink.rows_to_calls(calls.partitions[0].rows(), 30, 1)
    assert np.array_equal(nat.getGram(), oracle.c_similarity(30, off, idx, 1))    # partition 0 applied exactly once
