"""Parquet calls files (spark_examples_b200/parquet_calls.py): `RDD[Seq[Int]]` at rest, one row group per partition."""
import numpy as np
import pytest

import spark_examples_b200 as pkg
from spark_examples_b200 import parquet_calls as pc
from spark_examples_b200.variants_pca import VariantsPcaDriver


def _rows(seed, n, nv):
    rng = np.random.default_rng(seed)
    rows = [np.nonzero(rng.random(n) < 0.2)[0].astype(np.int32) for _ in range(nv)]
    rows[3] = np.zeros(0, np.int32)                                  # an empty variant (dropped at :166)
    rows[5] = np.array([2, 2, 7], np.int32)                          # a sample listed twice (multiplicity 2)
    off = np.zeros(nv + 1, np.int64)
    np.cumsum([len(r) for r in rows], out=off[1:])
    return rows, off, np.concatenate(rows)


def test_round_trip_row_groups_are_partitions(tmp_path):
    n, nv = 40, 100
    rows, off, idx = _rows(1, n, nv)
    callsets = [(f"ds-{i}", f"S{i}") for i in range(n)]
    path = str(tmp_path / "calls.parquet")
    pc.write_calls(path, callsets, off, idx, row_group_variants=32, start=np.arange(nv) + 1000)
    f = pc.CallsParquet(path)
    assert f.callsets == callsets and [s.nv for s in f.slices] == [32, 32, 32, 4]
    got = []
    for s in f.slices:
        b = s.load()
        assert b.offsets.dtype == np.int64 and b.idx.dtype == np.int32 and b.offsets[0] == 0
        got += [b.idx[b.offsets[v]:b.offsets[v + 1]].tolist() for v in range(len(b.offsets) - 1)]
    assert got == [r.tolist() for r in rows if len(r)]


def test_driver_reads_calls_parquet(tmp_path, capsys, oracle):
    n, nv = 25, 64
    rows, off, idx = _rows(2, n, nv)
    path = str(tmp_path / "calls.parquet")
    pc.write_calls(path, [(f"ds-{i}", f"S{i}") for i in range(n)], off, idx, row_group_variants=20)
    d = VariantsPcaDriver(pkg.PcaConf(["--calls-parquet-path", path]))
    assert "Matrix size: 25." in capsys.readouterr().out
    calls = d.getCallsRdd(d.getData)
    assert len(calls.partitions) == 4 and calls.collect() == [r.tolist() for r in rows if len(r)]
    # the same rows through the oracle's Gram: what getSimilarityMatrix computes on the GPU from these partitions
    b = [s.load() for s in calls.partitions]
    S = sum(oracle.c_similarity(n, x.offsets, x.idx, 1) for x in b)
    assert np.array_equal(S, oracle.c_similarity(n, off, idx, 1)) and S[2, 2] >= 4          # the doubled sample: 2 x 2
    with pytest.raises(ValueError, match="min-allele-frequency"):
        d2 = VariantsPcaDriver(pkg.PcaConf(["--calls-parquet-path", path, "--min-allele-frequency", "0.1"]))
        d2.filterDataset(d2.getData[0])


def test_rejects_foreign_files(tmp_path):
    import pyarrow as pa
    import pyarrow.parquet as pq
    p = str(tmp_path / "x.parquet")
    pq.write_table(pa.table({"a": [1, 2]}), p)
    with pytest.raises(ValueError, match="callsets"):
        pc.CallsParquet(p)
    with pytest.raises(ValueError, match="offsets"):
        pc.write_calls(p, [("a-0", "A")], np.array([0, 3, 2]), np.array([0, 0], np.int32))
