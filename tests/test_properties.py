"""Property tests (hypothesis) of the host-side formats: whatever the cohort shape, the packed forms decode to the same
`RDD[Seq[Int]]` rows that the reference's rule (VariantsPca.scala:56-60, 153-168) yields on the unpacked genotypes."""
import numpy as np
from hypothesis import given, settings, strategies as st

from spark_examples_b200 import parquet_calls, plink, vcf

_small = settings(max_examples=40, deadline=None)


@st.composite
def dosages(draw):
    n = draw(st.integers(1, 37))
    v = draw(st.integers(1, 29))
    flat = draw(st.lists(st.integers(-1, 2), min_size=n * v, max_size=n * v))
    return np.array(flat, np.int64).reshape(n, v)


@_small
@given(dosages())
def test_bed_rows_decode_to_the_reference_rule(tmp_path_factory, d):
    n, v = d.shape
    prefix = str(tmp_path_factory.mktemp("bed") / "c")
    plink.write_fileset(prefix, d)
    bed = plink.BedFile(prefix)
    rows = bed.rows(0, v)
    assert rows.shape == (v, (n + 3) // 4)
    assert np.array_equal(plink.decode_rows(rows, n, plink.COUNT_A1), d.T > 0)            # any A1 allele; missing = no-call
    assert np.array_equal(plink.decode_rows(rows, n, plink.COUNT_A2), (d.T == 0) | (d.T == 1))
    off, idx = plink.rows_to_calls(rows, n, plink.COUNT_A1)
    want = [np.nonzero(d[:, j] > 0)[0] for j in range(v)]
    want = [w for w in want if len(w)]
    assert len(off) - 1 == len(want) and all(np.array_equal(idx[off[r]:off[r + 1]], w) for r, w in enumerate(want))


@_small
@given(st.lists(st.lists(st.integers(0, 50), max_size=12), min_size=1, max_size=40), st.integers(1, 9))
def test_parquet_row_groups_preserve_rows(tmp_path_factory, rows, group):
    off = np.zeros(len(rows) + 1, np.int64)
    np.cumsum([len(r) for r in rows], out=off[1:])
    idx = np.array([x for r in rows for x in r], np.int32)
    path = str(tmp_path_factory.mktemp("pq") / "c.parquet")
    parquet_calls.write_calls(path, [(f"d-{i}", f"s{i}") for i in range(51)], off, idx, row_group_variants=group)
    got = []
    for s in parquet_calls.CallsParquet(path).slices:
        b = s.load()
        got += [b.idx[b.offsets[i]:b.offsets[i + 1]].tolist() for i in range(len(b.offsets) - 1)]
    assert got == [r for r in rows if r]                                                  # empty variants dropped (:166)


@_small
@given(st.lists(st.one_of(st.just("."), st.integers(0, 9).map(str)), min_size=1, max_size=4), st.booleans())
def test_gt_field_parses_to_allele_indices(alleles, phased):
    sep = "|" if phased else "/"
    genotype, phaseset = vcf._parse_gt(sep.join(alleles))
    assert genotype == tuple(-1 if a == "." else int(a) for a in alleles)
    assert phaseset == ("*" if phased and len(alleles) > 1 else "")
