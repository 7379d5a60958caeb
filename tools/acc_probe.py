"""Scratch: test accuracy of the C1-shaped Disc_QBM on synthetic images with a matched-filter projection."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, qbm_b200
import bench_train as BT
def data(num, seed=19):
    x, y = BT.synthetic_images(num, (28, 28), 10, seed)
    rng = np.random.default_rng(seed)
    templates = rng.random((10, 784)) < 0.5          # the generator's own first draw
    P = np.concatenate([(templates - 0.5) / 784 ** 0.5, np.random.default_rng(seed + 1).standard_normal((6, 784)) / 784 ** 0.5])
    z = x.reshape(num, -1).astype(np.float64) @ P.T
    z = (z - z.min(axis=0)) / np.maximum(z.max(axis=0) - z.min(axis=0), 1e-12)
    return z, np.eye(10)[y], y
X, Yoh, y = data(292 + 400)
ntr = 292
Xtr, Ytr, Xte, yte = X[:ntr], Yoh[:ntr], X[ntr:], y[ntr:]
for lr in (0.2, 0.6, 1.5):
    np.random.seed(19)
    m = qbm_b200.DiscQBM(dim_input=16, num_classes=10, use_one_hot_encoding=True, n_hidden_nodes=24, restricted=False,
                         sample_count=100, anneal_steps=1000, beta_eff=1.0, seed=19, stats_mode="loop")
    t = time.time()
    for ep in range(24):
        for s in range(0, ntr, 73):
            m.train_for_one_iteration(Xtr[s:s + 73], Ytr[s:s + 73], lr)
        if ep in (3, 7, 15, 23):
            acc = float(np.mean(m.predict_batch(Xte) == yte))
            print(f"lr {lr} epoch {ep + 1}: test acc {acc:.3f} ({time.time() - t:.1f}s)", flush=True)
