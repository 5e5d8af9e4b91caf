"""CPU tests of the host-side mirror of the reference interface (flags, records, encode rule, keys, output)."""
import io

import numpy as np
import pytest

import spark_examples_b200 as pkg
from spark_examples_b200 import variants_common as vc
from spark_examples_b200.jformat import jdouble
from spark_examples_b200.variants_pca import (CallsRdd, VariantsPcaDriver, _rows_to_batch, extractCallInfo, getVariantKey,
                                              murmur3_128)


def test_flag_names_and_defaults_match_reference():
    """GenomicsConf.scala:35-56, :77-85: camelCase vals -> --kebab-case flags, same defaults."""
    c = pkg.PcaConf([])
    assert c.basesPerPartition() == 1000000
    assert c.numReducePartitions() == 10
    assert c.references() == ["chr17:41196311:41277499"]
    assert c.variantSetId() == [pkg.GoogleGenomicsPublicData.Platinum_Genomes] == ["3049512673186936334"]
    assert c.numPc() == 2
    assert not c.allReferences() and not c.debugDatasets()
    assert not c.minAlleleFrequency.isDefined and not c.outputPath.isDefined and not c.inputPath.isDefined
    with pytest.raises(KeyError):
        c.outputPath()
    c = pkg.PcaConf(["--bases-per-partition", "5000", "--num-reduce-partitions", "3", "--output-path", "/tmp/x",
                     "--references", "1:0:10,2:5:9", "17:1:2", "--variant-set-id", "a", "b", "--spark-master", "local[4]",
                     "--all-references", "--debug-datasets", "--min-allele-frequency", "0.05", "--num-pc", "4",
                     "--client-secrets", "s.json", "--input-path", "in"])
    assert (c.basesPerPartition(), c.numReducePartitions(), c.outputPath(), c.sparkMaster()) == (5000, 3, "/tmp/x", "local[4]")
    assert c.references() == ["1:0:10,2:5:9", "17:1:2"] and c.variantSetId() == ["a", "b"]
    assert c.allReferences() and c.debugDatasets() and c.numPc() == 4 and abs(c.minAlleleFrequency() - 0.05) < 1e-12
    with pytest.raises(SystemExit):
        pkg.PcaConf(["--no-such-flag"])


def test_extract_call_info_matches_reference_rule(oracle):
    mapping = {"a": 0, "b": 1, "c": 2}
    v = pkg.Variant("17", calls=[pkg.Call("a", genotype=[0, 0]), pkg.Call("b", genotype=[0, 2]), pkg.Call("c", genotype=[-1, -1])])
    got = extractCallInfo(v, mapping)
    assert got == [pkg.CallData(False, 0), pkg.CallData(True, 1), pkg.CallData(False, 2)]
    assert [(c.hasVariation, c.callsetId) for c in got] == oracle.np_extract_call_info(
        [(c.callsetId, c.genotype) for c in v.calls], mapping)
    assert extractCallInfo(pkg.Variant("17", calls=None), mapping) == []
    with pytest.raises(KeyError):
        extractCallInfo(pkg.Variant("17", calls=[pkg.Call("zz", genotype=[1])]), mapping)


def _random_variants(rng, callsets, nv):
    out = []
    for i in range(nv):
        calls = []
        for cid, _ in callsets:
            if rng.random() < 0.9:
                calls.append(pkg.Call(cid, genotype=rng.choice([-1, 0, 0, 0, 1, 2], size=2).tolist()))
        out.append(pkg.Variant("chr1", start=i, end=i + 1, referenceBases="A", alternateBases=["C"],
                               info={"AF": [str(round(rng.random(), 3))]}, calls=calls if rng.random() < 0.95 else None))
    return out


def _driver(variants, callsets, extra=()):
    conf = pkg.PcaConf(list(extra))
    common = pkg.VariantsCommon(conf, callsets=callsets, datasets=[variants])
    return VariantsPcaDriver(conf, common=common)


def test_get_calls_rdd_equals_oracle_encode(oracle):
    rng = np.random.default_rng(11)
    callsets = [(f"ds-{i}", f"N{i:03d}") for i in range(23)]
    variants = _random_variants(rng, callsets, 300)
    d = _driver(variants, callsets, ["--variants-per-partition", "64"])
    rdd = d.getCallsRdd(d.getData)
    assert isinstance(rdd, CallsRdd) and len(rdd.partitions) == 5
    want = oracle.np_get_calls([None if v.calls is None else [(c.callsetId, c.genotype) for c in v.calls] for v in variants],
                               d.common.indexes)
    assert rdd.collect() == want
    assert all(len(r) > 0 for r in want)


def test_filter_dataset_min_allele_frequency():
    rng = np.random.default_rng(12)
    callsets = [("ds-0", "A"), ("ds-1", "B")]
    variants = _random_variants(rng, callsets, 50) + [pkg.Variant("chr1", info={}, calls=[])]
    d = _driver(variants, callsets, ["--min-allele-frequency", "0.4"])
    kept = [v for p in d.filterDataset(d.getData[0]).partitions for v in p]
    want = [v for v in variants if "AF" in v.info and np.float32(v.info["AF"][0]) >= np.float32(0.4)]
    assert kept == want and 0 < len(kept) < len(variants)
    d2 = _driver(variants, callsets)
    assert d2.filterDataset(d2.getData[0]) is d2.getData[0]


def test_variant_key_is_guava_murmur3_128():
    # MurmurHash3_x64_128 known answers, printed the way Guava's HashCode.toString does (little-endian bytes)
    assert murmur3_128(b"") == "0" * 32
    assert murmur3_128(b"hello") == "029bbd41b3a7d8cb191dae486a901e5b"
    assert murmur3_128(b"The quick brown fox jumps over the lazy dog") == "6c1b07bc7bbc4be347939ac4a93c437a"
    a = pkg.Variant("17", start=41196311, end=41196312, referenceBases="A", alternateBases=["C", "T"])
    b = pkg.Variant("17", start=41196311, end=41196312, referenceBases="A", alternateBases=["CT"])
    c = pkg.Variant("17", start=41196311, end=41196313, referenceBases="A", alternateBases=["CT"])
    assert getVariantKey(a) == getVariantKey(b) != getVariantKey(c)          # mkString("") (:63-64)
    assert len(getVariantKey(a)) == 32


def test_join_and_merge_datasets(oracle):
    callsets = [("x-0", "A"), ("x-1", "B"), ("y-0", "C")]
    mk = lambda pos, calls: pkg.Variant("1", start=pos, end=pos + 1, referenceBases="A", alternateBases=["G"], calls=calls)
    ds1 = [mk(1, [pkg.Call("x-0", genotype=[0, 1]), pkg.Call("x-1", genotype=[0, 0])]), mk(2, [pkg.Call("x-1", genotype=[1, 1])])]
    ds2 = [mk(1, [pkg.Call("y-0", genotype=[1, 0])]), mk(3, [pkg.Call("y-0", genotype=[1, 1])])]
    conf = pkg.PcaConf([])
    common = pkg.VariantsCommon(conf, callsets=callsets, datasets=[ds1, ds2])
    d = VariantsPcaDriver(conf, common=common)
    assert d.getCallsRdd(d.getData).collect() == [[0, 2]]          # only position 1 is in both sets
    conf3 = pkg.PcaConf([])
    common3 = pkg.VariantsCommon(conf3, callsets=callsets, datasets=[ds1, ds2, ds1])
    d3 = VariantsPcaDriver(conf3, common=common3)
    assert d3.getCallsRdd(d3.getData).collect() == [[0, 2, 0]]     # union of the three sides, duplicates kept


def test_rows_to_batch_filters_like_reference():
    rows = [[pkg.CallData(True, 3), pkg.CallData(False, 1)], [pkg.CallData(False, 0)], [], [pkg.CallData(True, 2), pkg.CallData(True, 2)]]
    b = _rows_to_batch(rows)
    assert b.offsets.tolist() == [0, 1, 3] and b.idx.tolist() == [3, 2, 2]


def test_jdouble_matches_java_layout_and_oracle(oracle):
    cases = {0.0286308791579312: "0.0286308791579312", -0.008456233951873527: "-0.008456233951873527", 1.0: "1.0",
             1e7: "1.0E7", 9999999.0: "9999999.0", 1.5e-5: "1.5E-5", 0.001: "0.001", -2.5e10: "-2.5E10", 0.0: "0.0",
             123456.75: "123456.75", float("nan"): "NaN", float("inf"): "Infinity"}
    for x, want in cases.items():
        assert jdouble(x) == want
        assert oracle._jdouble(x) == want
    rng = np.random.default_rng(5)
    for x in rng.standard_normal(200) * 10.0 ** rng.integers(-9, 9, 200):
        s = jdouble(x)
        assert float(s.replace("E", "e")) == x and s == oracle._jdouble(x)


def test_emit_result_format_and_file(tmp_path, oracle):
    callsets = [("dsA-2", "NA20811"), ("dsB-1", "NA20818"), ("dsA-9", "NA00001")]
    conf = pkg.PcaConf(["--output-path", str(tmp_path / "run")])
    d = VariantsPcaDriver(conf, common=pkg.VariantsCommon(conf, callsets=callsets, datasets=[[]]))
    result = [("dsA-2", 0.0286308791579312, -0.008456233951873527), ("dsB-1", -0.033609576645005836, -0.026655905606186293),
              ("dsA-9", 1e-5, 2.0)]
    buf = io.StringIO()
    d.emitResult(result, out=buf)
    lines = buf.getvalue().splitlines()
    assert lines == ["NA00001\tdsA\t1.0E-5\t2.0", "NA20811\tdsA\t0.0286308791579312\t-0.008456233951873527",
                     "NA20818\tdsB\t-0.033609576645005836\t-0.026655905606186293"]           # :238-239
    assert lines == oracle.emit_result_lines(result, dict(callsets))
    part = (tmp_path / "run-pca.tsv" / "part-00000").read_text().splitlines()
    assert part[0] == "NA20811\t0.0286308791579312\t-0.008456233951873527\tdsA"                # :243 column order


def test_variants_file_round_trip(tmp_path):
    rng = np.random.default_rng(3)
    callsets = [(f"f-{i}", f"S{i}") for i in range(7)]
    variants = _random_variants(rng, callsets, 40)
    path = tmp_path / "v.jsonl"
    vc.write_variants_file(str(path), callsets, variants)
    conf = pkg.PcaConf(["--input-path", str(path), "--variants-per-partition", "16"])
    common = pkg.VariantsCommon(conf)
    assert common.indexes == {cid: i for i, (cid, _) in enumerate(callsets)} and common.names["f-3"] == "S3"
    back = [v for p in common.data[0].partitions for v in p]
    assert [(v.start, None if v.calls is None else [(c.callsetId, tuple(c.genotype)) for c in v.calls]) for v in back] == \
           [(v.start, None if v.calls is None else [(c.callsetId, tuple(c.genotype)) for c in v.calls]) for v in variants]
    assert len(common.data[0].partitions) == 3


def test_no_source_is_an_error_not_a_silent_default():
    with pytest.raises(RuntimeError, match="retired"):
        pkg.VariantsCommon(pkg.PcaConf([]))


def test_synthetic_source_shapes():
    conf = pkg.PcaConf(["--synthetic", "100,1000,7", "--variants-per-partition", "300"])
    common = pkg.VariantsCommon(conf)
    assert len(common.indexes) == 100 and common.names["synth-000042"] == "S000042"
    sl = common.data[0].partitions
    assert [(s.v0, s.nv, s.seed) for s in sl] == [(0, 300, 7), (300, 300, 7), (600, 300, 7), (900, 100, 7)]


def test_compute_pca_needs_two_components():
    conf = pkg.PcaConf(["--num-pc", "1"])
    d = VariantsPcaDriver(conf, common=pkg.VariantsCommon(conf, callsets=[("a-0", "A"), ("a-1", "B")], datasets=[[]]))
    with pytest.raises(IndexError):
        d.computePca(None)


def test_joined_slice_rows_equal_the_record_level_join_and_merge():
    """getCallsRdd hands 2+ datasets to the GPU as a JoinedSlice; its host rendering (collect) must equal what the
    record-level joinDatasets / mergeDatasets (the dict-based mirror of :115-148) produce after :164-167."""
    from spark_examples_b200.variants_pca import _rows_to_batch
    rng = np.random.default_rng(17)
    callsets = [(f"p-{i}", f"P{i}") for i in range(9)] + [(f"q-{i}", f"Q{i}") for i in range(7)]

    def ds(cs, positions):
        return [pkg.Variant("1", start=int(p), end=int(p) + 1, referenceBases="A", alternateBases=["G"],
                            calls=[pkg.Call(cid, genotype=[int(rng.random() < 0.4), 0]) for cid, _ in cs]) for p in positions]

    d1, d2 = ds(callsets[:9], rng.integers(0, 60, 80)), ds(callsets[9:], rng.integers(0, 60, 80))   # repeated positions
    for datasets in ([d1, d2], [d1, d2, d2[:30]]):
        conf = pkg.PcaConf([])
        d = VariantsPcaDriver(conf, common=pkg.VariantsCommon(conf, callsets=callsets, datasets=datasets))
        got = sorted(d.getCallsRdd(d.getData).collect())
        recs = d.joinDatasets(d.getData) if len(datasets) == 2 else d.mergeDatasets(d.getData, len(datasets))
        b = _rows_to_batch(recs)
        want = sorted(b.idx[b.offsets[i]:b.offsets[i + 1]].tolist() for i in range(len(b.offsets) - 1))
        assert got == want and len(got) > 5
