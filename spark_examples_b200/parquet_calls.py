"""Parquet files of calls rows -- `RDD[Seq[Int]]` (VariantsPca.scala:153-168) at rest.

One row per variant, column `carriers: list<int32>` = the callset indices with variation at that variant (a sample may
appear twice, as a-2 allows); optional pass-through columns (`contig`, `start`, ...) are ignored.  The callsets live in
the file metadata (key `callsets`, JSON list of [id, name] in index order).  A Parquet row group is one partition: its
list column is already the CSR pair (offsets, values) that `vpca_accumulate_calls` takes, so a partition goes from disk
to the GPU without a per-record Python loop -- the offline replacement for re-running the ingestion
(rdd/VariantsRDD.scala:187-236) on every job, next to `--input-path` (GenomicsConf.scala:41).
"""
from __future__ import annotations

import json
from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np


def _pa():
    import pyarrow as pa
    import pyarrow.parquet as pq
    return pa, pq


def write_calls(path: str, callsets: Sequence[Tuple[str, str]], offsets: np.ndarray, idx: np.ndarray,
                row_group_variants: int = 65536, **columns) -> None:
    """offsets (nv + 1, any integer type) / idx: CSR rows; columns: optional per-variant arrays stored alongside."""
    pa, pq = _pa()
    off = np.ascontiguousarray(offsets, dtype=np.int64)
    if off[0] != 0 or np.any(np.diff(off) < 0) or off[-1] != len(idx):
        raise ValueError("offsets must start at 0, be non-decreasing and end at len(idx)")
    if off[-1] >= 2 ** 31:
        raise ValueError("more than 2^31 - 1 calls in one file: split it")
    arr = pa.ListArray.from_arrays(pa.array(off.astype(np.int32)), pa.array(np.ascontiguousarray(idx, dtype=np.int32)))
    cols = {"carriers": arr}
    for name, values in columns.items():
        cols[name] = pa.array(values)
    table = pa.table(cols).replace_schema_metadata({"callsets": json.dumps([[c[0], c[1]] for c in callsets])})
    pq.write_table(table, path, row_group_size=max(1, int(row_group_variants)))


@dataclass
class ParquetSlice:
    """One row group of a calls file; `load()` is its CSR pair with the empty rows dropped (VariantsPca.scala:166)."""
    file: "CallsParquet"
    row_group: int
    nv: int

    def load(self):
        from .variants_common import CallsBatch
        off, idx = self.file.read_row_group(self.row_group)
        counts = np.diff(off)
        keep = counts > 0
        if not keep.all():
            new_off = np.zeros(int(keep.sum()) + 1, np.int64)
            np.cumsum(counts[keep], out=new_off[1:])
            off = new_off                                  # the values of empty rows occupy no space: idx is unchanged
        return CallsBatch(off, idx)


class CallsParquet:
    def __init__(self, path: str):
        _, pq = _pa()
        self.path = path
        self._file = pq.ParquetFile(path)
        meta = self._file.schema_arrow.metadata or {}
        if b"callsets" not in meta:
            raise ValueError(f"{path}: no 'callsets' entry in the file metadata (written by write_calls)")
        self.callsets: List[Tuple[str, str]] = [(c[0], c[1]) for c in json.loads(meta[b"callsets"].decode("utf-8"))]
        if "carriers" not in self._file.schema_arrow.names:
            raise ValueError(f"{path}: no 'carriers' column")
        md = self._file.metadata
        self.slices = [ParquetSlice(self, g, md.row_group(g).num_rows) for g in range(md.num_row_groups)]

    def read_row_group(self, g: int):
        col = self._file.read_row_group(g, columns=["carriers"]).column("carriers").combine_chunks()
        if hasattr(col, "chunks"):                         # ChunkedArray on some pyarrow versions
            col = col.chunk(0) if col.num_chunks else col
        off = col.offsets.to_numpy().astype(np.int64)
        values = col.values.to_numpy()
        idx = np.ascontiguousarray(values[off[0]:off[-1]], dtype=np.int32)
        return off - off[0], idx
