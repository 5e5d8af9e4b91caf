"""GPU tests of the multi-dataset keying: variant keys hashed on the device (Guava murmur3_128 known
answers + the host restatement), the 2-way join and N-way merge of VariantsPca.scala:115-148 against the dict-based host
implementation, and the joined rows fed to the encoder / Gram without leaving the device."""
import numpy as np
import pytest

import spark_examples_b200 as pkg
from spark_examples_b200 import native
from spark_examples_b200.variants_common import JoinedSlice
from spark_examples_b200.variants_pca import (VariantsPcaDriver, joined_rows_on_host, murmur3_128, variantKeyBytes)

pytestmark = pytest.mark.gpu


def _hex(h):
    return np.ascontiguousarray(h, dtype="<u8").tobytes().hex()


def test_hash_keys_known_answers_and_host_restatement(oracle):
    rng = np.random.default_rng(4)
    keys = [b"", b"hello", b"The quick brown fox jumps over the lazy dog"]
    keys += [bytes(rng.integers(0, 256, int(l), dtype=np.uint8)) for l in list(range(0, 70)) + [127, 128, 129, 1000]]
    v = pkg.Variant("17", start=41196311, end=41196312, referenceBases="A", alternateBases=["C", "T"])
    keys.append(variantKeyBytes(v))
    with native.NativePca(64) as nat:
        got = nat.hashKeys(keys)
        assert nat.hashKeys([]).shape == (0, 2)
    assert _hex(got[0]) == "0" * 32                                     # Guava: murmur3_128().hashBytes(new byte[0])
    assert _hex(got[1]) == "029bbd41b3a7d8cb191dae486a901e5b"
    assert _hex(got[2]) == "6c1b07bc7bbc4be347939ac4a93c437a"
    for k, h in zip(keys, got):
        assert _hex(h) == murmur3_128(k) == oracle.np_murmur3_128(k).hex()


def _random_slice(rng, n, rows_per, datasets, mode, universe):
    """`datasets` datasets of `rows_per` rows each over `universe` distinct variant keys (so keys repeat across and inside
    datasets), random carrier lists (some empty)."""
    keys, lens, idx = [], [], []
    for d in range(datasets):
        for _ in range(rows_per):
            pos = int(rng.integers(0, universe))
            keys.append(b"chr%d" % (pos % 3) + int(pos).to_bytes(8, "little") + int(pos + 1).to_bytes(8, "little") + b"A" + b"G" * (pos % 4))
            c = int(rng.integers(0, 12)) if rng.random() < 0.9 else 0
            lens.append(c)
            idx.extend(rng.integers(0, n, c).tolist())
    off = np.zeros(len(keys) + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    return JoinedSlice(mode, keys, off, np.asarray(idx, np.int32), rows_per, datasets)


def _drop_empty(off, idx):
    rows = [idx[off[i]:off[i + 1]].tolist() for i in range(len(off) - 1)]
    return [r for r in rows if r]


@pytest.mark.parametrize("mode,datasets", [(native.JOIN, 2), (native.MERGE, 2), (native.MERGE, 3), (native.MERGE, 5)])
@pytest.mark.parametrize("rows_per,universe", [(1, 1), (40, 25), (3000, 2500), (20000, 60000)])
def test_join_and_merge_rows_equal_the_host_implementation(mode, datasets, rows_per, universe, oracle):
    rng = np.random.default_rng(rows_per * 7 + datasets)
    n = 97
    p = _random_slice(rng, n, rows_per, datasets, mode, universe)
    want = joined_rows_on_host(p)
    # the oracle's restatement of VariantsPca.scala:115-148 on (key, calls) records, keys hashed by the oracle itself
    recs = [(oracle.np_murmur3_128(k).hex(), p.idx[p.offsets[i]:p.offsets[i + 1]].tolist()) for i, k in enumerate(p.keys)]
    if mode == native.JOIN:
        ref_rows = oracle.np_join_datasets(recs[:p.n_left], recs[p.n_left:])
    else:
        per = len(recs) // datasets
        ref_rows = oracle.np_merge_datasets([recs[d * per:(d + 1) * per] for d in range(datasets)], p.variant_set_count)
    ref_rows = [r for r in ref_rows if r]                                         # :166 drops variants without carriers
    with native.NativePca(n, max_multiplicity=8) as nat:
        rows, nnz = nat.joinRows(p.mode, p.keys, p.offsets, p.idx, p.n_left, p.variant_set_count)
        off, idx = nat.joinFetch(rows, nnz)
    assert off[0] == 0 and off[-1] == nnz and (np.diff(off) >= 0).all()
    assert _drop_empty(off, idx) == _drop_empty(want.offsets, want.idx)          # same rows in the same order
    assert _drop_empty(off, idx) == ref_rows                                      # and the oracle's rows


def test_joined_rows_feed_the_gram_without_leaving_the_device(oracle):
    rng = np.random.default_rng(9)
    n = 300
    p = _random_slice(rng, n, 5000, 2, native.JOIN, 3000)
    want = joined_rows_on_host(p)
    S_want = oracle.c_similarity(n, want.offsets, want.idx, 2)
    with native.NativePca(n, max_multiplicity=24) as nat:
        rows, nnz = nat.joinRows(p.mode, p.keys, p.offsets, p.idx, p.n_left, p.variant_set_count)
        st0 = nat.stats()
        nat.accumulateJoined(3)
        nat.abort(3)                                   # a failed task: its staged contribution is discarded ...
        nat.accumulateJoined(3)                        # ... and the retry counts once
        nat.commit(3)
        st1 = nat.stats()
        nat.finalizeGram()
        S = nat.getGram()
    assert np.array_equal(S, S_want)
    assert st1["h2d_bytes"] == st0["h2d_bytes"]        # nothing was uploaded again: the joined rows were already there
    assert rows >= len(want.offsets) - 1


def test_two_and_three_datasets_through_the_driver(oracle):
    """VariantsPcaDriver.getCallsRdd on 2 datasets (join, :159) and 3 (merge, :160) -> getSimilarityMatrix on the GPU."""
    rng = np.random.default_rng(31)
    na, nb, nv = 40, 30, 500
    cs_a = [(f"a-{i}", f"A{i:03d}") for i in range(na)]
    cs_b = [(f"b-{i}", f"B{i:03d}") for i in range(nb)]

    def dataset(callsets, positions):
        out = []
        for pos in positions:
            calls = [pkg.Call(cid, genotype=[0, 1] if rng.random() < 0.3 else [0, 0]) for cid, _ in callsets]
            out.append(pkg.Variant("chr2", start=int(pos), end=int(pos) + 1, referenceBases="C", alternateBases=["T"], calls=calls))
        return out

    pos_a = rng.choice(2 * nv, nv, replace=False)
    pos_b = rng.choice(2 * nv, nv, replace=False)
    for datasets in ([dataset(cs_a, pos_a), dataset(cs_b, pos_b)],
                     [dataset(cs_a, pos_a), dataset(cs_b, pos_b), dataset(cs_b, pos_b[: nv // 2])]):
        conf = pkg.PcaConf([])
        common = pkg.VariantsCommon(conf, callsets=cs_a + cs_b, datasets=datasets)
        d = VariantsPcaDriver(conf, common=common)
        rdd = d.getCallsRdd(d.getData)
        assert isinstance(rdd.partitions[0], JoinedSlice)
        rows = rdd.collect()                                            # host route (dicts): the expected rows
        assert len(rows) > 10
        off = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int64)
        S_want = oracle.c_similarity(na + nb, off, np.asarray([c for r in rows for c in r], np.int32), 2)
        sim = d.getSimilarityMatrix(rdd)
        assert np.array_equal(sim.toArray(), S_want)
        d.stop()
