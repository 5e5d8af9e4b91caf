"""VCF ingestion (spark_examples_b200/vcf.py): records field for field as the reference's VariantsBuilder would build
them from the API (rdd/VariantsRDD.scala:103-160), then the reference's own host path on them: extractCallInfo,
--min-allele-frequency, the 2-way join and the N-way merge (VariantsPca.scala:56-168).  CPU only: from `RDD[Seq[Int]]`
on it is the accumulateCalls path the GPU tests cover."""
import numpy as np
import pytest

import spark_examples_b200 as pkg
from spark_examples_b200 import vcf
from spark_examples_b200.variants_pca import VariantsPcaDriver

SAMPLES = ["NA1", "NA2", "NA3", "NA4"]


def _records():
    return [
        dict(chrom="chr17", pos=41196312, id="rs1;rs1b", ref="G", alt=["A"], info={"AF": [0.25], "DB": None},
             gts=["0/1", "0|0", "1|1", "./."]),
        dict(chrom="17", pos=41196400, ref="AT", alt=["A", "ATT"], info={"AF": [0.01, 0.6]}, gts=["0/2", "0/0", "0/0", "1/0"]),
        dict(chrom="X", pos=100, ref="C", alt=["T"], info={"AF": [0.5]}, gts=["1/1", "1/1", "1/1", "1/1"]),
        dict(chrom="chrMT", pos=5, ref="C", alt=["T"], gts=["1", "0", "0", "0"]),
        dict(chrom="2", pos=7, ref="C", alt=["T"], info={"AF": [0.3]}, gts=["0/0", "0/0", "0/0", "0/0"]),
        dict(chrom="2", pos=9, ref="C", alt=[], gts=["0/0", "0/0", ".", "0/0"]),
    ]


@pytest.mark.parametrize("ext", [".vcf", ".vcf.gz"])
def test_records_field_for_field(tmp_path, ext):
    path = str(tmp_path / ("plat-genomes" + ext))
    vcf.write_vcf(path, SAMPLES, _records())
    callsets, meta = vcf.read_header(path)
    assert callsets == [("plat_genomes-0", "NA1"), ("plat_genomes-1", "NA2"), ("plat_genomes-2", "NA3"),
                        ("plat_genomes-3", "NA4")]
    assert meta[0] == "##fileformat=VCFv4.2"
    vs = list(vcf.read_variants(path))
    assert [v.contig for v in vs] == ["17", "17", "2", "2"]            # chr17 -> 17; X and chrMT dropped (:103-135)
    v0, v1, v2, v3 = vs
    assert (v0.start, v0.end, v0.referenceBases, v0.alternateBases, v0.names) == (41196311, 41196312, "G", ["A"],
                                                                                  ["rs1", "rs1b"])
    assert v0.info == {"AF": ["0.25"], "DB": []} and v0.variantSetId == "plat_genomes"
    assert [c.genotype for c in v0.calls] == [(0, 1), (0, 0), (1, 1), (-1, -1)]
    assert [c.phaseset for c in v0.calls] == ["", "*", "*", ""]
    assert [c.callsetId for c in v0.calls] == [c[0] for c in callsets] and v0.calls[2].callsetName == "NA3"
    assert (v1.start, v1.end, v1.alternateBases, v1.names) == (41196399, 41196401, ["A", "ATT"], None)
    assert v1.info["AF"] == ["0.01", "0.6"]
    assert v3.alternateBases is None and [c.genotype for c in v3.calls][2] == (-1,)


def test_region_filter_and_contig_rule(tmp_path):
    path = str(tmp_path / "a.vcf")
    vcf.write_vcf(path, SAMPLES, _records())
    regions = vcf.parse_regions(["chr17:41196311:41196399", "2:0:8"])
    assert regions == [("chr17", 41196311, 41196399), ("2", 0, 8)]
    assert [(v.contig, v.start) for v in vcf.read_variants(path, regions)] == [("17", 41196311), ("2", 6)]
    assert vcf.normalize_contig("chr17") == "17" and vcf.normalize_contig("17") == "17"
    assert vcf.normalize_contig("X") is None and vcf.normalize_contig("chrX") is None and vcf.normalize_contig("") == ""
    (tmp_path / "bad.vcf").write_text("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n1\t5\n")
    with pytest.raises(ValueError, match="8 tab-separated"):
        list(vcf.read_variants(str(tmp_path / "bad.vcf")))
    (tmp_path / "nohdr.vcf").write_text("1\t5\t.\tA\tC\t.\t.\t.\n")
    with pytest.raises(ValueError, match="#CHROM"):
        vcf.read_header(str(tmp_path / "nohdr.vcf"))


def test_single_file_through_the_reference_host_path(tmp_path, capsys):
    path = str(tmp_path / "cohort.vcf")
    vcf.write_vcf(path, SAMPLES, _records())
    conf = pkg.PcaConf(["--vcf-path", path])
    d = VariantsPcaDriver(conf)
    assert "Matrix size: 4." in capsys.readouterr().out
    rows = d.getCallsRdd(d.getData).collect()
    # hasVariation = any allele > 0 (:58); the no-call ./. has none; variants without carriers are dropped (:166)
    assert rows == [[0, 2], [0, 3]]
    # --min-allele-frequency keeps INFO AF[0] >= threshold (:96-111); records without AF are dropped
    d2 = VariantsPcaDriver(pkg.PcaConf(["--vcf-path", path, "--min-allele-frequency", "0.2"]))
    filtered = [d2.filterDataset(ds) for ds in d2.getData]                 # as main does (:41-42)
    assert d2.getCallsRdd(filtered).collect() == [[0, 2]]
    # explicit --references restricts the records like the API request ranges did
    d3 = VariantsPcaDriver(pkg.PcaConf(["--vcf-path", path, "--references", "17:41196390:41196500"]))
    assert d3.getCallsRdd(d3.getData).collect() == [[0, 3]]


def _two_files(tmp_path):
    a = [dict(chrom="1", pos=10, ref="A", alt=["C"], gts=["0/1", "0/0"]),
         dict(chrom="1", pos=20, ref="A", alt=["G"], gts=["1/1", "0/1"]),
         dict(chrom="1", pos=30, ref="T", alt=["C"], gts=["0/0", "0/1"])]
    b = [dict(chrom="chr1", pos=20, ref="A", alt=["G"], gts=["0/1", "0/0", "1/1"]),        # same key as a[1]
         dict(chrom="1", pos=30, ref="T", alt=["G"], gts=["0/1", "0/1", "0/1"]),            # different ALT: no match
         dict(chrom="1", pos=10, ref="A", alt=["C"], gts=["0/0", "0/0", "0/1"])]            # same key as a[0]
    pa, pb = str(tmp_path / "setA.vcf"), str(tmp_path / "setB.vcf")
    vcf.write_vcf(pa, ["a0", "a1"], a)
    vcf.write_vcf(pb, ["b0", "b1", "b2"], b)
    return pa, pb


def test_two_files_are_joined_on_the_variant_key(tmp_path, capsys):
    pa, pb = _two_files(tmp_path)
    d = VariantsPcaDriver(pkg.PcaConf(["--vcf-path", f"{pa},{pb}"]))
    assert "Matrix size: 5." in capsys.readouterr().out
    assert list(d.common.indexes) == ["setA-0", "setA-1", "setB-0", "setB-1", "setB-2"]
    rows = sorted(d.getCallsRdd(d.getData).collect())
    # inner join (:115-128): pos 10 -> a0 + b2; pos 20 -> a0, a1 + b0, b2; pos 30 differs in ALT -> no row
    assert rows == [[0, 1, 2, 4], [0, 4]]


def test_three_files_are_merged(tmp_path):
    pa, pb = _two_files(tmp_path)
    pc = str(tmp_path / "setC.vcf")
    vcf.write_vcf(pc, ["c0"], [dict(chrom="1", pos=20, ref="A", alt=["G"], gts=["0/1"]),
                               dict(chrom="1", pos=30, ref="T", alt=["C"], gts=["1/1"])])
    d = VariantsPcaDriver(pkg.PcaConf(["--vcf-path", f"{pa},{pb},{pc}"]))
    rows = d.getCallsRdd(d.getData).collect()
    # merge (:136-148) keeps keys present in all 3 sets: only pos 20 A>G
    assert [sorted(r) for r in rows] == [[0, 1, 2, 4, 5]]
