"""Training-throughput legs of bench.py: "QBM train images/sec" (BASELINE.json configs 1, 3, 5) and the
ClassificationRBM steps (config 2), on synthetic images of the named shapes.

One *step* = one minibatch of `batch` images per GPU through the complete training step (both phases,
statistics, [all-reduce of the parameter-shaped statistics over ranks,] SGD update).  Weak scaling: the
per-GPU minibatch is fixed, the global minibatch is `world x batch`.  An *image* = one training image
processed through both phases and its share of the update.

GPU legs import only the product package; the CPU legs (`cpu_*`) are the oracle ports of the reference's
per-image loop (oracle/model_oracle.py + oracle/neal_sa.c) on a bounded sample -- a reported baseline.
"""
from __future__ import annotations

import os
import threading
import time

import numpy as np

CONFIGS = {
    # name: (description, default per-GPU batch)
    "c1": ("C1 Disc_QBM 28x28 -> 16-dim input, 10 one-hot labels, 24 hidden, 100 reads x 1000 sweeps (QUBO n = 24 / 34)", 73),
    "c3": ("C3 Conv_Deep_QBM 18x18, 3x3 kernel, pool 2 -> 64 pooled units, 128 sequential units, 1000 reads x 1000 sweeps "
           "(QUBO n = 192 / 193)", 8),
    "c5": ("C5 Disc_QBM CIFAR-10 shape 32x32x3 -> 128-dim input, 10 one-hot labels, 512 hidden, 100 reads x 1000 sweeps "
           "(QUBO n = 512 / 522)", 64),
    "c2": ("C2 ClassificationRBM 784+10 visible, 500 hidden, batch 256, binarized 28x28", 256),
}


# ------------------------------------------------------------------------------------------------
# synthetic data: class templates, pixel-on probability 0.2 + 0.6 * template
# ------------------------------------------------------------------------------------------------
def synthetic_images(num, shape, num_classes, seed=19, binarize=True):
    rng = np.random.default_rng(seed)
    npx = int(np.prod(shape))
    templates = rng.random((num_classes, npx)) < 0.5
    y = rng.integers(0, num_classes, num)
    p = 0.2 + 0.6 * templates[y]
    x = (rng.random((num, npx)) < p) if binarize else np.clip(p + 0.15 * rng.standard_normal((num, npx)), 0, 1)
    return x.astype(np.float32).reshape((num,) + tuple(shape)), y.astype(np.int64)


def project_inputs(x, dim, seed=19):
    """fixed random projection of flattened images to `dim` inputs, rescaled to [0, 1]."""
    rng = np.random.default_rng(seed + 1)
    flat = x.reshape(x.shape[0], -1).astype(np.float64)
    z = flat @ (rng.standard_normal((flat.shape[1], dim)) / np.sqrt(flat.shape[1]))
    z = (z - z.min(axis=0)) / np.maximum(z.max(axis=0) - z.min(axis=0), 1e-12)
    return z


def make_data(cfg, num, seed=19):
    if cfg == "c1":
        x, y = synthetic_images(num, (28, 28), 10, seed)
        return project_inputs(x, 16, seed), np.eye(10)[y], y
    if cfg == "c5":
        x, y = synthetic_images(num, (3, 32, 32), 10, seed)
        return project_inputs(x, 128, seed), np.eye(10)[y], y
    if cfg == "c3":
        x, y = synthetic_images(num, (18, 18), 2, seed, binarize=False)
        return x, y, y
    if cfg == "c2":
        x, y = synthetic_images(num, (28, 28), 10, seed)
        return x.reshape(num, 784), y, y
    raise ValueError(cfg)


# ------------------------------------------------------------------------------------------------
# GPU legs
# ------------------------------------------------------------------------------------------------
def make_model(cfg, qbm, dev, pg=None):
    if cfg == "c1":
        np.random.seed(19)
        return qbm.DiscQBM(dim_input=16, num_classes=10, use_one_hot_encoding=True, n_hidden_nodes=24, restricted=False,
                           sample_count=100, anneal_steps=1000, beta_eff=1.0, seed=19, stats_mode="loop", device=dev,
                           process_group=pg)
    if cfg == "c5":
        np.random.seed(19)
        return qbm.DiscQBM(dim_input=128, num_classes=10, use_one_hot_encoding=True, n_hidden_nodes=512, restricted=False,
                           sample_count=100, anneal_steps=1000, beta_eff=1.0, seed=19, stats_mode="loop", device=dev,
                           process_group=pg)
    if cfg == "c3":
        return qbm.ConvDeepQBM(num_visible_nodes=324, num_lable_nodes=1, image_shape=(18, 18), kernel_size=3, pooling_size=2,
                               pooling_type="deterministic", stride=1, sequential_layer_sizes=[128], is_restricted=False,
                               hidden_bias_type="shared", solver="SA", anneal=1000, seed=44, device=dev, process_group=pg)
    if cfg == "c2":
        return qbm.B200ClassificationRBM(784, 500, k=1, num_classes=10, learning_rate=0.05, seed=19, device=dev,
                                         process_group=pg)
    raise ValueError(cfg)


def _step_fn(cfg, model, mode):
    if cfg in ("c1", "c5"):
        return lambda X, Y, gb, off: model.train_for_one_iteration(X, Y, 0.05, global_batch=gb, first_image=off)
    if cfg == "c3":
        return lambda X, Y, gb, off: model.train_one_iteration(X, Y, 1000, 1.0, 0.01, one_hot=False, global_batch=gb,
                                                               first_image=off)
    if cfg == "c2":
        if mode == "cd1":
            return lambda X, Y, gb, off: model.cd1_training(X, Y, global_batch=gb)
        return lambda X, Y, gb, off: model.discriminative_training(X, Y, global_batch=gb)[0].item()
    raise ValueError(cfg)


LAST_INFO = {}        # facts about the last gpu_train_rate run that bench.py adds to the JSON line


def gpu_train_rate(cfg, qbm, torch, dev, world, rank, barrier, batch, steps, warmup, pg=None, mode="disc"):
    """(images/s device-resident inputs, images/s end to end from host buffers, ms/step) for one config."""
    model = make_model(cfg, qbm, dev, pg)
    nsteps = warmup + 2 * steps
    X, Y, _ = make_data(cfg, batch * nsteps, seed=19 + rank)
    if cfg == "c2":
        Y = Y.astype(np.int32)
    fn = _step_fn(cfg, model, mode)
    gb = batch * world
    sl = lambda a, i: a[i * batch:(i + 1) * batch]
    # device-resident inputs: staged before the timed region
    Xd = torch.from_numpy(np.ascontiguousarray(X)).to(dev)
    Yd = torch.from_numpy(np.ascontiguousarray(Y)).to(dev)
    for i in range(warmup):
        fn(sl(Xd, i), sl(Yd, i), gb, rank * batch)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(warmup, warmup + steps):
        fn(sl(Xd, i), sl(Yd, i), gb, rank * batch)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    # end to end: numpy minibatches on the host, H2D inside the step, loss / parameters read back
    barrier()
    t0 = time.perf_counter()
    for i in range(warmup + steps, warmup + 2 * steps):
        fn(sl(X, i), sl(Y, i), gb, rank * batch)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    h2d = int(sl(X, 0).nbytes + sl(Y, 0).nbytes)
    LAST_INFO.clear()
    if hasattr(model, "release_graphs"):
        if pg is not None:
            LAST_INFO["dp_reduce"] = "peer-memory (signal / wait / reduce in rank order / apply in one pass)" if model._peer else "nccl all-reduce + apply"
            LAST_INFO["peer_error"] = bool(model.peer_error())
        model.release_graphs()                  # captured NCCL collectives / peer mappings must not outlive the process group
    return ms, e2e_s, h2d


# ------------------------------------------------------------------------------------------------
# CPU legs (oracle ports of the reference's per-image loop), bounded samples, threads over images
# ------------------------------------------------------------------------------------------------
def _threads_over(items, work, threads):
    out = [None] * len(items)
    idx = list(range(len(items)))
    lock = threading.Lock()

    def run():
        while True:
            with lock:
                if not idx:
                    return
                i = idx.pop()
            out[i] = work(items[i])

    ths = [threading.Thread(target=run) for _ in range(min(threads, len(items)))]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    return out, time.perf_counter() - t0


def cpu_train_rate(cfg, images, threads):
    """images/s of the reference's per-image training loop restated on the CPU (oracle ports): for every
    image both QUBOs, two neal-restatement sampler calls and both statistics passes."""
    from oracle import oracle as O
    from oracle import model_oracle as M
    O.lib()
    X, Y, ylab = make_data(cfg, images, seed=7)
    if cfg in ("c1", "c5"):
        di, h = (16, 24) if cfg == "c1" else (128, 512)
        np.random.seed(19)
        W_hh = np.triu(np.random.uniform(-1, 1, (h, h)), k=1)
        np.random.seed(19)
        p = dict(W_vh=np.random.uniform(-1, 1, (10 + di, h)), W_vo=np.random.uniform(-1, 1, (di, 10)),
                 W_oo=np.triu(np.random.uniform(-1, 1, (10, 10)), k=1), b_h=np.random.uniform(-1, 1, h),
                 b_o=np.random.uniform(-1, 1, 10), W_hh=W_hh)

        def work(i):
            Sc = O.sample_Q_reference(M.disc_qubo(p, X[i], Y[i]), 100, 1000, seed=19)
            Su = O.sample_Q_reference(M.disc_qubo(p, X[i], None), 100, 1000, seed=19)
            c = M.disc_stats_loop(Sc, X[i], Y[i], 10, di, h)
            u = M.disc_stats_loop(Su, X[i], None, 10, di, h)
            return [a - b for a, b in zip(c, u)]
    elif cfg == "c3":
        rng = np.random.default_rng(44)
        p = dict(kernel=rng.uniform(-1, 1, (3, 3)), W_seq=[rng.uniform(-1, 1, (64, 128))],
                 W_intra=[np.triu(np.tile(rng.uniform(-1, 1, 128), (128, 1)))], W_hy=rng.uniform(-1, 1, (128, 1)),
                 W_oo=np.zeros((1, 1)), b_conv=rng.uniform(-1, 1, 1), b_seq=rng.uniform(-1, 1, 128), b_out=rng.uniform(-1, 1, 1))

        def work(i):
            flat, pooled, patches = M.convdeep_context(X[i], p["kernel"], 1, 2)
            lab = np.array([float(Y[i])])
            Sc = O.sample_Q_reference(M.convdeep_qubo(p, flat, pooled, lab), 1000, 1000, seed=44)
            Su = O.sample_Q_reference(M.convdeep_qubo(p, flat, pooled, None), 1000, 1000, seed=44)
            c = M.convdeep_stats(Sc, X[i], lab, 64, [128], 1, patches)
            u = M.convdeep_stats(Su, X[i], None, 64, [128], 1, patches)
            return c[0] - u[0]
    else:
        raise ValueError(cfg)
    _, dt = _threads_over(list(range(images)), work, threads)
    return images / dt, dt


def cpu_rbm_rate(steps, batch=256, mode="disc"):
    """images/s of the ClassificationRBM step restated in float32 numpy (oracle/model_oracle.py), all BLAS threads."""
    from oracle import model_oracle as M
    rng = np.random.default_rng(3)
    V, H, C = 784, 500, 10
    W = (rng.standard_normal((V, H)) * 0.1).astype(np.float32)
    U = np.zeros((C, H), np.float32); bv = np.full(V, 0.5, np.float32); bh = np.zeros(H, np.float32); bc = np.zeros(C, np.float32)
    X, Y, _ = make_data("c2", batch * steps, seed=5)
    t0 = time.perf_counter()
    for s in range(steps):
        x, y = X[s * batch:(s + 1) * batch], Y[s * batch:(s + 1) * batch]
        if mode == "disc":
            new, _, _, _ = M.rbm_discriminative_step(W, U, bv, bh, bc, x, y, np.float32(0.05))
            W, U, bv, bh, bc = (new[k].astype(np.float32) for k in ("W", "U", "b_v", "b_h", "b_c"))
        else:
            oh = np.eye(C, dtype=np.float32)[y]
            ph0 = M.rbm_sample_hidden(W, U, bh, x, oh)
            h0 = (rng.random(ph0.shape, dtype=np.float32) < ph0).astype(np.float32)
            v1 = (rng.random(x.shape, dtype=np.float32) < M.rbm_sample_visible(W, bv, h0)).astype(np.float32)
            pc = M.rbm_sample_class(U, bc, h0)
            y1 = (pc.cumsum(axis=1) > rng.random((batch, 1), dtype=np.float32)).argmax(axis=1)
            oh1 = np.eye(C, dtype=np.float32)[y1]
            ph1 = M.rbm_sample_hidden(W, U, bh, v1, oh1)
            sc = np.float32(0.05 / batch)
            W = W + sc * (x.T @ ph0 - v1.T @ ph1)
            U = U + sc * (oh.T @ ph0 - oh1.T @ ph1)
            bv = bv + sc * (x - v1).sum(axis=0); bh = bh + sc * (ph0 - ph1).sum(axis=0); bc = bc + sc * (oh - oh1).sum(axis=0)
    dt = time.perf_counter() - t0
    return steps * batch / dt, dt
