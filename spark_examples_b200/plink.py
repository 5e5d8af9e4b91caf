"""PLINK 1 binary filesets (.bed / .bim / .fam) as a variants source.

The reference streams `Variant` records from the Google Genomics API (rdd/VariantsRDD.scala:187-236), which is retired;
cohorts of the size this path is built for live on disk as PLINK filesets or VCF.  A .bed file is already the packed
wire format of SURVEY.md 8f-1: variant-major, two bits per sample, four samples per byte (low bits first):

    0b00 homozygous A1   0b01 missing   0b10 heterozygous   0b11 homozygous A2        (A1, A2 = .bim columns 5, 6)

`hasVariation` (VariantsPca.scala:58, `genotype.foldLeft(false)(_ || _ > 0)`) becomes "carries the counted allele":
codes {00, 10} when A1 is counted (PLINK's default: A1 is the minor / alternate allele), {10, 11} when A2 is.  A missing
call is a no-call (-1, -1) and has no variation, exactly like the Scala rule.  Rows go to the GPU as they are on disk
(`NativePca.accumulateBed`, N/4 bytes per variant); `decode_rows` is the host-side statement of the same rule used by
`CallsRdd.collect()` and by the tests.
"""
from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path
from typing import List, Sequence, Tuple

import numpy as np

BED_MAGIC = bytes([0x6C, 0x1B, 0x01])          # 0x01 = variant-major
COUNT_A1, COUNT_A2 = 1, 2


def _prefix(path: str) -> str:
    p = str(path)
    for ext in (".bed", ".bim", ".fam"):
        if p.endswith(ext):
            return p[: -len(ext)]
    return p


def read_fam(path: str) -> List[Tuple[str, str]]:
    """[(callset id, callset name)] in file order.  id = "FID-IID" so that `callsetId.split("-").head`
    (VariantsPca.scala:235) yields the family id as the dataset column; name = IID."""
    out = []
    with open(_prefix(path) + ".fam", "r", encoding="utf-8") as fh:
        for line in fh:
            f = line.split()
            if len(f) >= 2:
                out.append((f"{f[0]}-{f[1]}", f[1]))
    if len({c[0] for c in out}) != len(out):
        raise ValueError(f"{path}: duplicate FID-IID in .fam")
    return out


@dataclass(frozen=True)
class BimRecord:
    contig: str
    id: str
    position: int
    a1: str
    a2: str


def read_bim(path: str) -> List[BimRecord]:
    out = []
    with open(_prefix(path) + ".bim", "r", encoding="utf-8") as fh:
        for line in fh:
            f = line.split()
            if len(f) >= 6:
                out.append(BimRecord(f[0], f[1], int(f[3]), f[4], f[5]))
    return out


class BedFile:
    """Memory-mapped .bed: `rows(v0, v1)` is the (v1 - v0) x ceil(N / 4) uint8 block of variants [v0, v1)."""

    def __init__(self, path: str, n_samples: int | None = None, n_variants: int | None = None):
        self.prefix = _prefix(path)
        self.n_samples = n_samples if n_samples is not None else len(read_fam(self.prefix))
        self.stride = (self.n_samples + 3) // 4
        bed = Path(self.prefix + ".bed")
        size = bed.stat().st_size
        with open(bed, "rb") as fh:
            magic = fh.read(3)
        if magic[:2] != BED_MAGIC[:2]:
            raise ValueError(f"{bed}: not a PLINK .bed file")
        if magic[2:3] != BED_MAGIC[2:3]:
            raise ValueError(f"{bed}: sample-major .bed files are not supported (re-export with plink --make-bed)")
        if self.stride == 0 or (size - 3) % self.stride != 0:
            raise ValueError(f"{bed}: size {size} does not match {self.n_samples} samples")
        self.n_variants = (size - 3) // self.stride
        if n_variants is not None and n_variants != self.n_variants:
            raise ValueError(f"{bed}: {self.n_variants} variants on disk, {n_variants} in the .bim")
        self._map = np.memmap(bed, dtype=np.uint8, mode="r", offset=3, shape=(self.n_variants, self.stride)) \
            if self.n_variants else np.zeros((0, self.stride), np.uint8)

    def rows(self, v0: int, v1: int) -> np.ndarray:
        return np.ascontiguousarray(self._map[v0:v1])


def decode_rows(rows: np.ndarray, n_samples: int, counted: int = COUNT_A1) -> np.ndarray:
    """(nv, ceil(N/4)) uint8 -> (nv, N) bool `hasVariation` matrix (the rule in the module docstring)."""
    rows = np.asarray(rows, dtype=np.uint8)
    codes = np.stack([(rows >> s) & 3 for s in (0, 2, 4, 6)], axis=-1).reshape(rows.shape[0], -1)[:, :n_samples]
    if counted == COUNT_A1:
        return (codes == 0) | (codes == 2)
    if counted == COUNT_A2:
        return (codes == 2) | (codes == 3)
    raise ValueError("counted allele must be 1 (A1) or 2 (A2)")


def rows_to_calls(rows: np.ndarray, n_samples: int, counted: int = COUNT_A1):
    """The `RDD[Seq[Int]]` form (CSR offsets int64, idx int32) of a block of .bed rows; variants without any carrier are
    dropped, as VariantsPca.scala:166 does."""
    has = decode_rows(rows, n_samples, counted)
    has = has[has.any(axis=1)]
    counts = has.sum(axis=1)
    off = np.zeros(len(counts) + 1, np.int64)
    np.cumsum(counts, out=off[1:])
    idx = np.nonzero(has)[1].astype(np.int32)
    return off, idx


def write_fileset(prefix: str, dosage_a1: np.ndarray, fam: Sequence[Tuple[str, str]] | None = None,
                  contig: str = "17", start: int = 41196311) -> None:
    """Write a fileset from an (N samples) x (V variants) array of A1 allele counts in {0, 1, 2} (-1 = missing);
    fam = [(FID, IID)] (default ("synth", "S000000"), ...).  Used by the tests and to export the synthetic cohort."""
    d = np.asarray(dosage_a1)
    n, v = d.shape
    code = np.full((v, n), 1, np.uint8)                     # missing
    dt = d.T
    code[dt == 2] = 0
    code[dt == 1] = 2
    code[dt == 0] = 3
    pad = (-n) % 4
    if pad:
        code = np.concatenate([code, np.zeros((v, pad), np.uint8)], axis=1)     # PLINK pads with 0 bits
    c4 = code.reshape(v, -1, 4)
    packed = (c4[:, :, 0] | (c4[:, :, 1] << 2) | (c4[:, :, 2] << 4) | (c4[:, :, 3] << 6)).astype(np.uint8)
    prefix = _prefix(prefix)
    with open(prefix + ".bed", "wb") as fh:
        fh.write(BED_MAGIC)
        fh.write(packed.tobytes())
    with open(prefix + ".bim", "w", encoding="utf-8") as fh:
        for j in range(v):
            fh.write(f"{contig}\trs{j + 1}\t0\t{start + j}\tA\tG\n")
    with open(prefix + ".fam", "w", encoding="utf-8") as fh:
        for i in range(n):
            fid, iid = fam[i] if fam is not None else ("synth", f"S{i:06d}")
            fh.write(f"{fid} {iid} 0 0 0 -9\n")
