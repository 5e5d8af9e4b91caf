"""Two-GPU tests (skipped on boxes with fewer than 2 devices): variant-sharded Gram reduced (a) by the host-driven
NCCL all-reduce, (b) by the fused peer-memory epilogue into every rank's Gram and (c) by the fused reduce-scatter into
row-band owners + all-gather, all equal to the oracle's Gram of the whole cohort."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, nv, out_dir):
    import torch
    import torch.distributed as dist
    from spark_examples_b200 import native
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)
        per = nv // world
        X = torch.empty((n, per), dtype=torch.int8, device="cuda")
        from oracle import oracle as _o
        calls = _o.c_synth_calls(20240901, n, rank * per, per)     # the same shard as index rows (test input only)
        # (a) NCCL
        S = torch.zeros((n, n), dtype=torch.int32, device="cuda")
        with native.NativePca(n, device=rank, stream=stream.cuda_stream, d_gram=S.data_ptr()) as nat:
            nat.synthDenseDevice(20240901, rank * per, per, 0, X.data_ptr(), per)
            nat.accumulateDenseDevice(X.data_ptr(), per, per)
            dist.all_reduce(S)
            nat.finalizeGram()
            np.save(os.path.join(out_dir, f"nccl_{rank}.npy"), nat.getGram())
        # (b) fused epilogue over peer memory, two passes to exercise reset + barrier
        with native.NativePca(n, device=rank, stream=stream.cuda_stream) as nat:
            handles = [None] * world
            dist.all_gather_object(handles, nat.exportIpcHandle())
            nat.setPeers(handles, rank)
            for _ in range(2):
                nat.reset()
                nat.peerBarrier()
                nat.accumulateDenseDevice(X.data_ptr(), per, per)
                nat.peerBarrier()
                nat.finalizeGram()
                G = nat.getGram()
            np.save(os.path.join(out_dir, f"fused_{rank}.npy"), G)
            dist.barrier()          # nobody unmaps a peer buffer while another rank may still touch it
        # (c) fused reduce-scatter by Gram row bands + all-gather over peer memory; the second pass goes through a
        #     staged partition (commit adds into the owners) to cover that route too
        with native.NativePca(n, device=rank, stream=stream.cuda_stream) as nat:
            handles = [None] * world
            dist.all_gather_object(handles, nat.exportIpcHandle())
            nat.setPeers(handles, rank, mode="owner_rows")
            for p in range(2):
                nat.reset()
                nat.peerBarrier()
                if p == 0:
                    nat.accumulateDenseDevice(X.data_ptr(), per, per)
                else:
                    off, idx = calls
                    nat.accumulateCalls(7 + rank, off, idx)
                    nat.commit(7 + rank)
                nat.gatherGram()
                nat.finalizeGram()
                np.save(os.path.join(out_dir, f"owner{p}_{rank}.npy"), nat.getGram())
            dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_two_gpu_nccl_and_fused_reduce(tmp_path, oracle):
    import torch.multiprocessing as mp
    n, nv, world = 777, 40_000, 2
    mp.spawn(_worker, args=(world, _free_port(), n, nv, str(tmp_path)), nprocs=world, join=True)
    want = oracle.np_similarity_dense(oracle.c_synth_dense(20240901, n, 0, nv))
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"nccl_{r}.npy"), want)
        assert np.array_equal(np.load(tmp_path / f"fused_{r}.npy"), want)
        assert np.array_equal(np.load(tmp_path / f"owner0_{r}.npy"), want)
        assert np.array_equal(np.load(tmp_path / f"owner1_{r}.npy"), want)
