"""N>1 path on CPU: two gloo ranks shard a batch contiguously, reduce the time with MAX, and the
gathered per-shard results equal the unsharded ones (replicas are independent and deterministic)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from roomnet_b200.sharding import aggregate_throughput, reduce_max, split_contiguous


def test_split_contiguous_matches_scheduler_rule():
    assert split_contiguous(8192, 8) == [(i * 1024, (i + 1) * 1024) for i in range(8)]
    assert split_contiguous(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert split_contiguous(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    for n in (0, 1, 7, 64, 8191):
        for g in (1, 2, 3, 8):
            s = split_contiguous(n, g)
            assert s[0][0] == 0 and s[-1][1] == n and all(a[1] == b[0] for a, b in zip(s, s[1:]))
            assert max(e - b for b, e in s) - min(e - b for b, e in s) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle.roomnet_oracle import RoomNetOracle, synthetic_suite
    imgs = synthetic_suite(4)
    b, e = split_contiguous(len(imgs), world)[rank]
    orc = RoomNetOracle(dtype=np.float32, conv_backend="torch").load()
    local = orc.forward(orc.normalise(imgs[b:e]))["logits"].astype(np.float32)
    # "gathering only the logits to the host": every rank owns a disjoint slice of one output buffer
    gathered = [torch.zeros(e - b, 6) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(local))
    dist.barrier()
    t_max = reduce_max(0.5 + rank)  # slowest rank defines the step time
    if rank == 0:
        np.save(os.path.join(out_dir, "logits.npy"), torch.cat(gathered).numpy())
        np.save(os.path.join(out_dir, "tmax.npy"), np.array([t_max]))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding(tmp_path, golden):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    logits = np.load(tmp_path / "logits.npy")
    assert logits.shape == (4, 6)
    assert np.abs(logits - golden["logits"][:4]).max() < 2e-4
    assert np.load(tmp_path / "tmax.npy")[0] == 1.5
    assert aggregate_throughput(256, 2, 10, 1.5) == 2 * 256 * 10 / 1.5
