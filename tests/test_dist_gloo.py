"""CPU, world_size 2, gloo: the multi-GPU plumbing of the training steps -- the product's `shard_range` and
`all_reduce_sum_` (qbm_b200.dist) around a flat statistics buffer laid out like the trainers' error buffer
[statistics..., loss], followed by the update K9 applies (`param -= lr * err / global_batch`, restated here in torch
because K9 is CUDA) -- gives the single-process result on every rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _pack(tensors):
    """the trainers' flat float64 error buffer: every statistic row-major, in parameter order, the loss last"""
    return torch.cat([t.reshape(-1).to(torch.float64) for t in tensors])


def _sgd_apply(params, flat, lr, global_batch):
    """K9 (csrc/disc_qbm.cu qbm_sgd_apply) restated: param -= lr * (err / global_batch) over consecutive slices"""
    pos = 0
    for p in params:
        cnt = p.numel()
        p -= lr * (flat[pos:pos + cnt].reshape(p.shape).to(p.dtype) / global_batch)
        pos += cnt
    return pos


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import qbm_b200.dist as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        B, shapes = 7, [(5,), (3, 4), (4, 4), (1,)]
        # per-image statistics (clamped - unclamped) of a minibatch of B images, identical on both ranks
        per_image = [[torch.from_numpy(rng.standard_normal(s)) for s in shapes] for _ in range(B)]
        losses = torch.from_numpy(rng.random(B))
        params0 = [torch.from_numpy(rng.standard_normal(s)) for s in shapes]
        lo, hi = D.shard_range(B, world, rank)
        err = [sum(per_image[i][k] for i in range(lo, hi)) for k in range(len(shapes))]
        flat = _pack(err + [losses[lo:hi].sum()])
        D.all_reduce_sum_(flat, dist.group.WORLD)
        params = [p.clone() for p in params0]
        used = _sgd_apply(params, flat, 0.3, float(B))
        assert used == flat.numel() - 1
        # single-process reference
        ref = [p.clone() for p in params0]
        full = _pack([sum(per_image[i][k] for i in range(B)) for k in range(len(shapes))] + [losses.sum()])
        _sgd_apply(ref, D.all_reduce_sum_(full, None), 0.3, float(B))
        ok = all(torch.allclose(a, b, rtol=0, atol=1e-13) for a, b in zip(params, ref))
        ok = ok and abs(float(flat[-1]) - float(losses.sum())) < 1e-12
        # every rank ends with the same parameters
        chk = _pack(params).clone()
        gathered = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(gathered, chk)
        ok = ok and all(torch.equal(g, gathered[0]) for g in gathered)
        open(os.path.join(out_dir, f"rank{rank}.txt"), "w").write("ok" if ok else "mismatch")
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    import qbm_b200.dist as D
    for total in (0, 1, 7, 64, 100, 101):
        for world in (1, 2, 3, 8):
            blocks = [D.shard_range(total, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.shard_range(4, 2, 2)


def test_two_rank_gloo_step_equals_single_process(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert open(tmp_path / f"rank{r}.txt").read() == "ok"
