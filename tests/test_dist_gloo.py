"""world_size-2 `gloo` test of the N > 1 host logic on CPU: partitions are dealt round-robin, every rank builds the
partial Gram of its own partitions (here with the oracle, the GPU is not available), one all-reduce sums them, and
every rank ends with the Gram of the whole cohort -- VariantsPca.scala:184-190."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, nv, per_part, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle
        from spark_examples_b200 import dist as vdist
        assert vdist.rank_world() == (rank, world)
        nparts = (nv + per_part - 1) // per_part
        mine = vdist.my_partitions(nparts, rank, world)
        S = np.zeros((n, n), np.int32)
        for p in mine:
            off, idx = oracle.c_synth_calls(20240901, n, p * per_part, min(per_part, nv - p * per_part))
            S += oracle.c_similarity(n, off, idx, 1)
        t = torch.from_numpy(S)
        vdist.allreduce_gram(t)
        np.save(os.path.join(out_dir, f"gram_rank{rank}.npy"), t.numpy())
        np.save(os.path.join(out_dir, f"parts_rank{rank}.npy"), np.array(mine))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_variant_sharding_and_allreduce(tmp_path, oracle):
    n, nv, per_part, world = 96, 1000, 128, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, nv, per_part, str(tmp_path)), nprocs=world, join=True)
    off, idx = oracle.c_synth_calls(20240901, n, 0, nv)
    want = oracle.c_similarity(n, off, idx, 2)
    parts = [np.load(tmp_path / f"parts_rank{r}.npy").tolist() for r in range(world)]
    assert sorted(parts[0] + parts[1]) == list(range(8)) and not set(parts[0]) & set(parts[1])
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"gram_rank{r}.npy"), want)


def test_partition_ownership_is_a_partition():
    from spark_examples_b200 import dist as vdist
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            seen += vdist.my_partitions(37, r, world)
        assert sorted(seen) == list(range(37))
    assert vdist.rank_world() == (0, 1)
