"""VCF 4.x text files (plain or .gz) as a `Variant` source.

The reference gets `Variant` / `Call` records from the Google Genomics API (rdd/VariantsRDD.scala:187-236,
`VariantsBuilder.build` at :115-160), which is retired.  This reader produces the same records from a VCF, field for
field, so everything downstream is the reference's own path: `extractCallInfo` (VariantsPca.scala:56-60),
`--min-allele-frequency` on INFO `AF` (:96-111), the 2-way join / N-way merge on the variant key (:115-148) when several
files are given, and the encode + Gram on the GPU.

Field mapping (API v1 `Variant` as the reference's builder reads it):
    contig          CHROM through `VariantsBuilder.normalize` ("chr17" -> "17"; X / Y / MT records are dropped, :103-135)
    start           POS - 1 (0-based, half-open like the API)
    end             start + len(REF)              id     "<stem>:<CHROM>:<POS>:<REF>:<ALT>" (the API's id is opaque)
    names           ID split on ';' (None for '.') referenceBases / alternateBases  REF / ALT split on ','
    info            INFO key -> list of strings (flags -> [])
    calls           one per sample column: callsetId "<stem>-<column index>", callsetName = the column header,
                    genotype = GT allele indices ('.' -> -1; '/' and '|' both separate alleles), phaseset "*" when '|'
A record without a GT field gets `calls = None`, which `extractCallInfo` turns into an empty row (:57).
"""
from __future__ import annotations

import gzip
import io
import os
import re
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

from .records import Call, Variant


def _open(path: str) -> io.TextIOBase:
    if path.endswith(".gz") or path.endswith(".bgz"):
        return io.TextIOWrapper(gzip.open(path, "rb"), encoding="utf-8")
    return open(path, "r", encoding="utf-8")


def dataset_stem(path: str) -> str:
    """The dataset name of a file: its basename up to the first '.', with '-' replaced (callset ids are
    "<dataset>-<index>" and VariantsPca.scala:235 takes `callsetId.split("-").head` as the dataset)."""
    return os.path.basename(path).split(".")[0].replace("-", "_") or "vcf"


def read_header(path: str) -> Tuple[List[Tuple[str, str]], List[str]]:
    """([(callset id, callset name)] in column order, meta lines)."""
    meta: List[str] = []
    stem = dataset_stem(path)
    with _open(path) as fh:
        for line in fh:
            if line.startswith("##"):
                meta.append(line.rstrip("\n"))
                continue
            if line.startswith("#CHROM"):
                cols = line.rstrip("\n").split("\t")
                samples = cols[9:] if len(cols) > 9 else []
                return [(f"{stem}-{i}", name) for i, name in enumerate(samples)], meta
            break
    raise ValueError(f"{path}: no #CHROM header line")


_REF_NAME = re.compile(r"([a-z]*)?([0-9]*)")


def normalize_contig(reference_name: str) -> Optional[str]:
    """`VariantsBuilder.normalize` (rdd/VariantsRDD.scala:103-110): optional lower-case prefix + digits -> the digits
    ("chr17" -> "17"); any other name (X, Y, MT, chrX, ...) -> None and `build` drops the record (:134-135)."""
    m = _REF_NAME.fullmatch(reference_name)
    return m.group(2) if m else None


def _parse_gt(gt: str) -> Tuple[Tuple[int, ...], str]:
    phased = "|" in gt
    alleles = tuple(-1 if a in (".", "") else int(a) for a in gt.replace("|", "/").split("/"))
    return alleles, ("*" if phased else "")


def parse_regions(references: Sequence[str]) -> List[Tuple[str, int, int]]:
    """`--references` values, "contig:start:end" (GenomicsConf.scala:47-52), several per value separated by ','."""
    out = []
    for ref in references:
        for item in ref.split(","):
            item = item.strip()
            if not item:
                continue
            contig, start, end = item.rsplit(":", 2)
            out.append((contig, int(start), int(end)))
    return out


def read_variants(path: str, regions: Optional[Sequence[Tuple[str, int, int]]] = None) -> Iterator[Variant]:
    """Stream the records of one file.  regions: keep records whose start lies in [start, end) of a listed contig."""
    stem = dataset_stem(path)
    callset_ids: List[str] = []
    names: List[str] = []
    with _open(path) as fh:
        for lineno, line in enumerate(fh, 1):
            if line.startswith("##") or not line.strip():
                continue
            f = line.rstrip("\n").split("\t")
            if line.startswith("#CHROM"):
                names = f[9:]
                callset_ids = [f"{stem}-{i}" for i in range(len(names))]
                continue
            if len(f) < 8:
                raise ValueError(f"{path}:{lineno}: a VCF record needs at least 8 tab-separated columns")
            chrom, pos, vid, ref, alt, _qual, _flt, info_s = f[:8]
            contig = normalize_contig(chrom)
            if contig is None:
                continue                                   # VariantsBuilder.build returns None (:134-135)
            start = int(pos) - 1
            if regions is not None and not any(normalize_contig(c) == contig and s <= start < e for c, s, e in regions):
                continue
            info: Dict[str, List[str]] = {}
            if info_s not in (".", ""):
                for kv in info_s.split(";"):
                    if not kv:
                        continue
                    k, sep, v = kv.partition("=")
                    info[k] = v.split(",") if sep else []
            alts = None if alt in (".", "") else alt.split(",")
            calls = None
            if len(f) > 9:
                keys = f[8].split(":")
                if "GT" in keys:
                    gi = keys.index("GT")
                    if len(f) - 9 != len(callset_ids):
                        raise ValueError(f"{path}:{lineno}: {len(f) - 9} sample columns, header has {len(callset_ids)}")
                    calls = []
                    for j, col in enumerate(f[9:]):
                        parts = col.split(":")
                        genotype, phaseset = _parse_gt(parts[gi] if gi < len(parts) else ".")
                        calls.append(Call(callset_ids[j], names[j], genotype, phaseset=phaseset))
            yield Variant(contig=contig, id=f"{stem}:{chrom}:{pos}:{ref}:{alt}",
                          names=None if vid in (".", "") else vid.split(";"), start=start, end=start + len(ref),
                          referenceBases=ref, alternateBases=alts, info=info, variantSetId=stem, calls=calls)


def write_vcf(path: str, samples: Sequence[str], records: Sequence[dict]) -> None:
    """Minimal writer for tests and examples.  records: dicts with chrom, pos, ref, alt (list), optional id, info
    (dict), gts (list of 'a/b' strings, one per sample)."""
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "wt", encoding="utf-8") as fh:
        fh.write("##fileformat=VCFv4.2\n")
        fh.write('##INFO=<ID=AF,Number=A,Type=Float,Description="Allele Frequency">\n')
        fh.write('##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n')
        fh.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(samples) + "\n")
        for r in records:
            info = r.get("info") or {}
            info_s = ";".join(k if v is None else f"{k}={','.join(str(x) for x in v)}" for k, v in info.items()) or "."
            fh.write("\t".join([str(r["chrom"]), str(r["pos"]), r.get("id", "."), r["ref"], ",".join(r["alt"]) or ".",
                                ".", "PASS", info_s, "GT"] + list(r["gts"])) + "\n")
