"""One process, several GPU contexts (SURVEY 8b "process model": one driver JVM, tasks running concurrently,
VariantsPca.scala:38-50, :184-190): vpca_pool, vpca_gram_set_peers_local and band-only Grams.

On a 1-GPU box the contexts share device 0 -- the same kernels (owner-rows epilogue, commit into owners, flag barriers,
push of the row bands) run as on G devices, so the driver's 1-GPU round-end run exercises the fused reduce too; with
more devices the contexts are spread over them."""
import os
import subprocess
import threading
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
SEED = 20240901


def _devices(world):
    import torch
    nd = max(1, torch.cuda.device_count())
    return [g % nd for g in range(world)]


def _split_rows(off, idx, parts):
    """cut CSR rows into `parts` contiguous partitions (offsets rebased)"""
    nv = len(off) - 1
    out = []
    for p in range(parts):
        r0, r1 = nv * p // parts, nv * (p + 1) // parts
        out.append(((off[r0:r1 + 1] - off[r0]).astype(np.int64), idx[off[r0]:off[r1]]))
    return out


@pytest.mark.parametrize("world", [2, 4])
def test_pool_threads_match_oracle(oracle, world):
    """T task threads feed P partitions through the pool (int32 / uint16 / bitmap wire formats, one abort + retry, one
    poisoned partition); reduceAndFinalize = reduceByKey (:190); S bit-exact, PCs within 1e-6."""
    from spark_examples_b200 import native
    n, nv, P, T = 700, 6000, 16, 8
    off, idx = oracle.c_synth_calls(SEED, n, 0, nv)
    S_want = oracle.c_similarity(n, off, idx, 2)
    parts = _split_rows(off, idx, P)
    errors = []
    with native.NativePcaPool(n, world, devices=_devices(world), max_multiplicity=1, partitions_in_flight=T + 2,
                              chunk_variants=8192, chunk_nnz=1 << 22) as pool:
        assert pool.size == world
        for _pass in range(2):                                      # second pass: reset and the same analysis again
            pool.reset()
            nxt = iter(range(P))
            lock = threading.Lock()

            def task():
                while True:
                    with lock:
                        pid = next(nxt, None)
                    if pid is None:
                        return
                    o, ix = parts[pid]
                    try:
                        if pid % 5 == 1:                            # a task that fails half way and is retried
                            pool.accumulateCalls(pid, o, ix)
                            pool.abort(pid)
                        if pid % 7 == 3:                            # a corrupt batch poisons only its own partition
                            bad = ix.copy()
                            bad[len(bad) // 2] = n + 5
                            with pytest.raises(IndexError):
                                pool.accumulateCalls(pid, o, bad)
                        if pid % 3 == 0:
                            pool.accumulateCalls16(pid, o, ix)
                        elif pid % 3 == 1:
                            pool.accumulateCalls(pid, o, ix)
                        else:
                            bits = np.zeros((len(o) - 1, (n + 7) // 8), np.uint8)
                            rows = np.repeat(np.arange(len(o) - 1), np.diff(o))
                            np.bitwise_or.at(bits, (rows, ix // 8), (1 << (ix % 8)).astype(np.uint8))
                            pool.accumulateBits(pid, bits)
                        pool.commit(pid)
                    except Exception as exc:                        # surfaced after the join
                        errors.append((pid, exc))

            threads = [threading.Thread(target=task) for _ in range(T)]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
            assert not errors, errors
            pool.reduceAndFinalize()
            assert np.array_equal(pool.getGram(), S_want)
        vecs, evals, nz = pool.computePca(2)
        st = pool.stats()
    U, _ = oracle.compute_pca(S_want, 2)
    assert np.all(oracle.eigvec_rel_err(vecs, U) <= 1e-6)
    assert st["variants_accumulated"] == len(off) - 1


@pytest.mark.parametrize("mode", ["owner_rows", "replicate"])
def test_same_process_fused_reduce_of_resident_shards(oracle, mode):
    """The multi-GPU bench path inside ONE process: every context owns a variant shard that is resident in HBM, the Gram
    kernel's epilogue adds straight into the peers (owner of the row band / every rank), then barrier (+ push of the
    bands).  All contexts must end with the oracle's matrix of the whole cohort."""
    import torch
    from spark_examples_b200 import native
    world, n, per, P = 2, 1092, 4096, 2048
    devs = _devices(world)
    off, idx = oracle.c_synth_calls(SEED, n, 0, world * per)
    S_want = oracle.c_similarity(n, off, idx, 2)
    ctxs, bufs = [], []
    try:
        for r in range(world):
            ctxs.append(native.NativePca(n, device=devs[r], max_multiplicity=1))
        native.setPeersLocal(ctxs, mode)
        for r, c in enumerate(ctxs):
            with torch.cuda.device(devs[r]):
                buf = torch.empty(c.panelBytes(per, P), dtype=torch.uint8, device=f"cuda:{devs[r]}")
            bufs.append(buf)
            c.synthPanelsDevice(SEED, r * per, per, 0, buf.data_ptr(), P)
        for _pass in range(2):
            for c in ctxs:
                c.reset()
            for c in ctxs:
                c.synchronize()                      # all Grams are zero before any peer adds into them
            for r, c in enumerate(ctxs):
                c.accumulatePanels(bufs[r].data_ptr(), per, P)
            for c in ctxs:
                c.gatherGram()                       # enqueued on every stream before the host blocks on any
            for c in ctxs:
                c.finalizeGram()
            for c in ctxs:
                assert np.array_equal(c.getGram(), S_want)
    finally:
        for c in ctxs:
            c.synchronize()
        for c in ctxs:
            c.close()


def test_band_only_grams_hold_the_owner_rows(oracle):
    """Biobank form (BASELINE configs[3]; sizing note VariantsPca.scala:176-177): every context allocates ONLY the row band
    it owns; kernels flush to the owners; the bands ARE the result (no gather).  Small N here, the layout is the same."""
    import torch
    from spark_examples_b200 import native
    world, n, per, P = 4, 1400, 2048, 2048
    devs = _devices(world)
    bands = native.ownerRowBands(n, world)
    assert bands[0][0] == 0 and sum(b[1] for b in bands) == n and all(b[1] % 32 == 0 for b in bands[:-1])
    off, idx = oracle.c_synth_calls(SEED, n, 0, world * per)
    S_want = np.tril(oracle.c_similarity(n, off, idx, 2))
    ctxs, bufs = [], []
    try:
        for r in range(world):
            ctxs.append(native.NativePca(n, device=devs[r], max_multiplicity=1, gram_band=bands[r]))
        native.setPeersLocal(ctxs, "owner_rows")
        for r, c in enumerate(ctxs):
            buf = torch.empty(c.panelBytes(per, P), dtype=torch.uint8, device=f"cuda:{devs[r]}")
            bufs.append(buf)
            c.synthPanelsDevice(SEED, r * per, per, 0, buf.data_ptr(), P)
        for c in ctxs:
            c.reset()
        for c in ctxs:
            c.synchronize()
        for r, c in enumerate(ctxs):
            c.accumulatePanels(bufs[r].data_ptr(), per, P)
        for c in ctxs:
            c.gatherGram()                           # closing barrier only: bands stay where they are
        for c in ctxs:
            c.finalizeGram()
        for r, c in enumerate(ctxs):
            row0, rows = bands[r]
            got = np.tril(c.gramBand(row0, rows), k=row0)        # lower-triangle part of the band
            assert np.array_equal(got, S_want[row0:row0 + rows])
            with pytest.raises(native.VpcaError):
                c.getGram()                                       # a band is not the whole matrix
    finally:
        for c in ctxs:
            c.synchronize()
        for c in ctxs:
            c.close()


@pytest.mark.parametrize("dtype_name", ["i8", "e2m1"])
def test_owner_computes_bands_need_no_reduction(oracle, dtype_name):
    """The other biobank form (SURVEY 8e "shard output tiles across GPUs ... no reduction"): band-only contexts WITHOUT peers.
    Every context is fed ALL variants and its Gram kernel enumerates only the tiles of the rows it stores, so nothing is
    flushed to anybody and nothing is produced twice.  Two launches (two halves of the cohort) to cover accumulation."""
    import torch
    from spark_examples_b200 import native
    world, n, nv, P = 4, 1400, 8192, 2048
    dtype = {"i8": native.DTYPE_I8, "e2m1": native.DTYPE_E2M1}[dtype_name]
    devs = _devices(world)
    bands = native.ownerRowBands(n, world)
    off, idx = oracle.c_synth_calls(SEED, n, 0, nv)
    S_want = np.tril(oracle.c_similarity(n, off, idx, 2))
    ctxs = []
    try:
        for r in range(world):
            ctxs.append(native.NativePca(n, device=devs[r], dtype=dtype, max_multiplicity=1, gram_band=bands[r]))
        for r, c in enumerate(ctxs):
            with torch.cuda.device(devs[r]):
                buf = torch.empty(c.panelBytes(nv // 2, P), dtype=torch.uint8, device=f"cuda:{devs[r]}")
                for half in range(2):
                    c.synthPanelsDevice(SEED, half * (nv // 2), nv // 2, 0, buf.data_ptr(), P)
                    c.accumulatePanels(buf.data_ptr(), nv // 2, P)
                    c.synchronize()
            c.finalizeGram()
            row0, rows = bands[r]
            got = c.gramBand(row0, rows)
            assert np.array_equal(np.tril(got, k=row0), S_want[row0:row0 + rows])
            assert not np.triu(got, k=row0 + 1).any()            # nothing above the diagonal, nothing outside the band's tiles
            assert c.stats()["variants_accumulated"] == nv
    finally:
        for c in ctxs:
            c.synchronize()
        for c in ctxs:
            c.close()


def test_band_only_biobank_scale_n(oracle):
    """N = 70 000 across 4 band-only contexts (the whole matrix would be 19.6 GB per context; the bands sum to that once):
    64-bit addressing of the virtual Gram origin, spot-checked rows against an fp32 matmul."""
    import torch
    from spark_examples_b200 import native
    world, n, per, P = 4, 70_000, 1024, 1024
    free, _ = torch.cuda.mem_get_info()
    if free < 30 * 2 ** 30:
        pytest.skip("needs 30 GB of free HBM")
    devs = _devices(world)
    bands = native.ownerRowBands(n, world)
    ctxs, bufs = [], []
    try:
        for r in range(world):
            ctxs.append(native.NativePca(n, device=devs[r], max_multiplicity=1, gram_band=bands[r]))
        native.setPeersLocal(ctxs, "owner_rows")
        for r, c in enumerate(ctxs):
            buf = torch.empty(c.panelBytes(per, P), dtype=torch.uint8, device=f"cuda:{devs[r]}")
            bufs.append(buf)
            c.synthPanelsDevice(SEED, r * per, per, 0, buf.data_ptr(), P)
        for c in ctxs:
            c.reset()
        for c in ctxs:
            c.synchronize()
        for r, c in enumerate(ctxs):
            c.accumulatePanels(bufs[r].data_ptr(), per, P)
        for c in ctxs:
            c.gatherGram()
        for c in ctxs:
            c.synchronize()
        X = torch.cat([b.to("cuda:0").view(torch.int8).view(n, P) for b in bufs], dim=1).to(torch.float32)
        for r, c in enumerate(ctxs):
            row0, rows = bands[r]
            pick = sorted({row0, row0 + 1, row0 + rows // 2, row0 + rows - 1})
            for row in pick:
                got = c.gramBand(row, 1)[0]
                want = (X[row:row + 1] @ X[:row + 1].t()).to(torch.int32).cpu().numpy()[0]
                assert np.array_equal(got[:row + 1], want), (r, row)
    finally:
        for c in ctxs:
            c.synchronize()
        for c in ctxs:
            c.close()


def _run(cmd):
    env = dict(os.environ, OMP_NUM_THREADS="8")
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert proc.returncode == 0, f"{cmd}\nstdout: {proc.stdout}\nstderr: {proc.stderr}"
    return proc.stdout


def _harness(name):
    import __graft_entry__ as entry
    path = ROOT / "tests" / "_build" / name
    if not path.exists():
        entry.build()
    return str(path)


@pytest.mark.parametrize("contexts,threads", [(1, 8), (4, 8), (8, 12)])
def test_c_abi_from_many_threads_in_one_process(contexts, threads):
    """tests/abi_threads.c: pthreads drive accumulate / commit / abort (with retries) on `contexts` GPU contexts of ONE
    process through the C ABI only; the reduced Gram must equal the oracle's bit for bit on every context."""
    import json
    import torch
    nd = max(1, torch.cuda.device_count())
    out = _run([_harness("abi_threads"), str(contexts), str(threads), "640", "24", "900", "1", "20240901", str(nd)])
    rep = json.loads(out.strip().splitlines()[-1])
    assert rep["gram_bit_exact_vs_oracle"] is True
    assert rep["retries_after_abort"] >= 7 and rep["retries_after_bad_index"] >= 2
    assert rep["partitions_on_uint16_wire"] == 12
    assert rep["non_zero_rows"] == 640 and rep["eval0"] > rep["eval1"] > 0


def test_jni_shim_runs_against_a_mock_jvm():
    """spark_examples_b200/jvm/vpca_jni.c executed through a mock JNIEnv: NativePcaPool end to end with Java-array,
    pinned-direct-buffer and bitmap inputs, abort + retry, IndexOutOfBounds on a corrupt row; Gram bit-exact."""
    import json
    out = _run([_harness("jni_harness"), "gpu", "520", "3000", "1"])
    rep = json.loads(out.strip().splitlines()[-1])
    assert rep["failures"] == 0 and rep["gram_entries_differing"] == 0 and rep["non_zero_rows"] == 520
