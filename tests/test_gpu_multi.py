"""GPU, 2 ranks over NCCL (skipped on a single-GPU box): data-parallel training steps equal the single-GPU step
on the whole minibatch, and read-sharded sampling equals one launch."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import faulthandler
    import sys
    faulthandler.dump_traceback_later(150, exit=True)       # a rank stuck in a collective: show where, do not hang the box
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    import qbm_b200
    from qbm_b200.dist import shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    res = {}
    try:
        pg = dist.group.WORLD
        # ---- Disc_QBM (C1 shapes): sharded minibatch + all-reduce == whole minibatch on one GPU ----
        g = np.load(os.path.join(G, "disc_qbm_loop_onehot.npz"))
        X, Y = g["X"], g["Y"]
        mk = lambda pg_: (np.random.seed(19), qbm_b200.DiscQBM(dim_input=16, num_classes=10, use_one_hot_encoding=True,
                          n_hidden_nodes=24, sample_count=60, anneal_steps=200, seed=19, stats_mode="loop", device=dev,
                          process_group=pg_))[1]
        single, dp = mk(None), mk(pg)
        single.train_for_one_iteration(X, Y, 0.1)
        lo, hi = shard_range(len(X), world, rank)
        dp.train_for_one_iteration(X[lo:hi], Y[lo:hi], 0.1, global_batch=len(X), first_image=lo)
        a, b = single.get_params(), dp.get_params()
        res["disc"] = max(float(np.abs(a[k] - b[k]).max()) for k in a)
        # ---- Conv_Deep_QBM ----
        g = np.load(os.path.join(G, "convdeep_binary.npz"))
        X, Y = g["X"], g["Y"]
        mk = lambda pg_: qbm_b200.ConvDeepQBM(100, 1, image_shape=(10, 10), kernel_size=3, pooling_size=2,
                                              sequential_layer_sizes=[12], hidden_bias_type="shared", anneal=200, seed=44,
                                              device=dev, process_group=pg_)
        single, dp = mk(None), mk(pg)
        l1 = single.train_one_iteration(X, Y, 40, 1.0, 0.05)
        lo, hi = shard_range(len(X), world, rank)
        l2 = dp.train_one_iteration(X[lo:hi], Y[lo:hi], 40, 1.0, 0.05, global_batch=len(X), first_image=lo)
        a, b = single.get_params(), dp.get_params()
        d = 0.0
        for k in a:
            for u, v in zip(a[k] if isinstance(a[k], list) else [a[k]], b[k] if isinstance(b[k], list) else [b[k]]):
                d = max(d, float(np.abs(u - v).max()))
        res["convdeep"] = max(d, abs(l1 - l2))
        # ---- ClassificationRBM discriminative step (float32) ----
        rng = np.random.default_rng(3)
        xb = (rng.random((64, 784)) < 0.3).astype(np.float32)
        yb = rng.integers(0, 10, 64)
        mk = lambda pg_: qbm_b200.B200ClassificationRBM(784, 500, 1, num_classes=10, learning_rate=0.05, seed=7, device=dev,
                                                        process_group=pg_)
        single, dp = mk(None), mk(pg)
        e1, _, _ = single.discriminative_training(xb, yb)
        lo, hi = shard_range(64, world, rank)
        e2, _, _ = dp.discriminative_training(xb[lo:hi], yb[lo:hi], global_batch=64)
        res["rbm"] = max(float((single.weights - dp.weights).abs().max()), float((single.class_weights - dp.class_weights).abs().max()),
                         float((single.hidden_bias - dp.hidden_bias).abs().max()), abs(float(e1) - float(e2)))
        # ---- sampler: reads sharded by global index == one launch ----
        Q = np.triu(rng.uniform(-1, 1, (150, 150)))
        full, _, _ = qbm_b200.sample_qubo_batch(Q, 64, 200, seed=9, initial_states_generator="philox", device=dev,
                                                return_energy=False)
        lo, hi = shard_range(64, world, rank)
        part, _, _ = qbm_b200.sample_qubo_batch(Q, hi - lo, 200, seed=9, initial_states_generator="philox", device=dev,
                                                return_energy=False, chain_offset=lo)
        res["sa"] = float(np.abs(full[0, lo:hi].astype(int) - part[0].astype(int)).max())
        # the sharded drop-in: every rank gets all reads, identical to the single-GPU call (uneven shards: 33 reads)
        one = qbm_b200.B200SASampler(num_sweeps=200, seed=9, device=dev).sample_Q(Q, 33)
        both = qbm_b200.B200SASampler(num_sweeps=200, seed=9, device=dev, process_group=pg).sample_Q(Q, 33)
        res["sa_gather"] = float(np.abs(one - both).max()) + (0.0 if both.shape == (33, 150) else 1.0)
        # K2 on the shards + all-gather of the energies: samples AND energies of the sharded call equal the single-GPU call
        s1, e1, _ = qbm_b200.sample_qubo_batch(Q, 33, 200, seed=9, device=dev, return_energy=True)
        s2, e2, _ = qbm_b200.sample_qubo_batch(Q, 33, 200, seed=9, device=dev, return_energy=True, process_group=pg)
        res["sa_energy"] = float(np.abs(s1.astype(int) - s2.astype(int)).max()) + float(np.abs(e1 - e2).max()) + \
            (0.0 if e2.shape == (1, 33) else 1.0)
        # the problem taken from rank 0 only (NCCL broadcast of Q): the other rank passes zeros of the right shape
        Qr = Q if rank == 0 else np.zeros_like(Q)
        s3, e3, _ = qbm_b200.sample_qubo_batch(Qr, 33, 200, seed=9, device=dev, return_energy=True, process_group=pg, src_rank=0)
        res["sa_bcast"] = float(np.abs(s1.astype(int) - s3.astype(int)).max()) + float(np.abs(e1 - e3).max())
        # seed=None: drawn on rank 0 and broadcast, so every rank returns the same (all) reads
        s4, _, _ = qbm_b200.sample_qubo_batch(Q, 10, 50, seed=None, device=dev, return_energy=False, process_group=pg)
        t4 = torch.from_numpy(s4.astype(np.int32)).to(dev)
        t4max = t4.clone(); dist.all_reduce(t4max, op=dist.ReduceOp.MAX)
        res["sa_seed"] = float((t4max - t4).abs().max())
        # ---- ClassificationRBM CD-1, data-parallel: rank r draws from Philox stream step * world + r; the all-reduced update
        # equals the sum of the two shard updates run separately (single-GPU steps with lr * B_local / B_global)
        mk = lambda pg_: qbm_b200.B200ClassificationRBM(784, 500, 1, num_classes=10, learning_rate=0.05, seed=7, device=dev,
                                                        process_group=pg_)
        dp, alone = mk(pg), mk(None)
        W0, U0, bv0 = dp.weights.clone(), dp.class_weights.clone(), dp.visible_bias.clone()
        lo, hi = shard_range(64, world, rank)
        dp.cd1_training(xb[lo:hi], yb[lo:hi], global_batch=64)
        alone._step = rank                                   # stream = step * 1 + 0 = rank, the stream this rank used above
        alone.learning_rate = 0.05 * (hi - lo) / 64.0
        alone.cd1_training(xb[lo:hi], yb[lo:hi])
        deltas = torch.cat([(alone.weights - W0).reshape(-1), (alone.class_weights - U0).reshape(-1),
                            (alone.visible_bias - bv0).reshape(-1)])
        dist.all_reduce(deltas, op=dist.ReduceOp.SUM)
        got = torch.cat([(dp.weights - W0).reshape(-1), (dp.class_weights - U0).reshape(-1), (dp.visible_bias - bv0).reshape(-1)])
        res["rbm_cd1"] = float((got - deltas).abs().max())
        res["rbm_cd1_wt"] = float((dp._Wt[:, :784] - dp._W[:, :500].t()).abs().max())
        # ---- several steps: from the second one on the data-parallel step (shard kernels + NCCL all-reduce + apply) is replayed
        # as one CUDA graph; it must follow the eager launches (use_graphs=False) and, for the discriminative step, the
        # single-GPU model on the whole minibatch
        mkg = lambda pg_, ug: qbm_b200.B200ClassificationRBM(784, 500, 1, num_classes=10, learning_rate=0.05, seed=7, device=dev,
                                                             process_group=pg_, use_graphs=ug)
        single, dpg, dpe = mkg(None, True), mkg(pg, True), mkg(pg, False)
        cg, ce = mkg(pg, True), mkg(pg, False)
        d_disc = d_cd1 = d_single = 0.0
        for step in range(4):
            xs = (rng.random((64, 784)) < 0.3).astype(np.float32)
            ys = rng.integers(0, 10, 64)
            l0, _, _ = single.discriminative_training(xs, ys)
            l1, _, p1 = dpg.discriminative_training(xs[lo:hi], ys[lo:hi], global_batch=64)
            l2, _, p2 = dpe.discriminative_training(xs[lo:hi], ys[lo:hi], global_batch=64)
            d_disc = max(d_disc, float((dpg.weights - dpe.weights).abs().max()), float((dpg.class_weights - dpe.class_weights).abs().max()),
                         abs(float(l1) - float(l2)), float((p1 - p2).abs().max()))
            d_single = max(d_single, float((dpg.weights - single.weights).abs().max()), abs(float(l1) - float(l0)))
            cg.cd1_training(xs[lo:hi], ys[lo:hi], global_batch=64)
            ce.cd1_training(xs[lo:hi], ys[lo:hi], global_batch=64)
            d_cd1 = max(d_cd1, float((cg.weights - ce.weights).abs().max()), float((cg.class_weights - ce.class_weights).abs().max()),
                        float((cg.visible_bias - ce.visible_bias).abs().max()))
        res["rbm_graph_disc"], res["rbm_graph_cd1"], res["rbm_graph_vs_single"] = d_disc, d_cd1, d_single
        res["rbm_graph_captured"] = 0.0 if any(isinstance(e, dict) for e in dpg._graphs.values()) and \
            any(isinstance(e, dict) for e in cg._graphs.values()) else 1.0
        for m in (single, dpg, dpe, cg, ce, dp, alone):
            m.release_graphs()                  # captured NCCL collectives must go before the communicator does
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), res, allow_pickle=True)
    finally:
        dist.destroy_process_group()


def test_two_gpu_data_parallel_equals_single_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        res = np.load(tmp_path / f"rank{r}.npy", allow_pickle=True).item()
        assert res["disc"] < 1e-12, res          # float64 statistics: only the all-reduce summation order differs
        assert res["convdeep"] < 1e-9, res
        assert res["sa_energy"] == 0.0 and res["sa_bcast"] == 0.0 and res["sa_seed"] == 0.0, res
        assert res["rbm_cd1"] < 2e-5 and res["rbm_cd1_wt"] == 0.0, res
        assert res["rbm"] < 2e-5, res            # float32 parameters, TF32 products
        assert res["sa"] == 0.0, res             # bit-identical reads
        assert res["sa_gather"] == 0.0, res      # sharded + all-gathered sample_Q == single-GPU sample_Q
        # graph replays (incl. the captured all-reduce) == eager launches; 4 data-parallel steps == 4 single-GPU steps
        assert res["rbm_graph_captured"] == 0.0 and res["rbm_graph_disc"] < 1e-6 and res["rbm_graph_cd1"] < 1e-6, res
        assert res["rbm_graph_vs_single"] < 1e-4, res
