"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import qbm_b200

dev = torch.device("cuda:0")
rng = np.random.default_rng(1)
for n, reads, flags in [(24, 20, 0), (150, 12, 0), (150, 20, 16), (150, 9, 32), (600, 6, 0), (600, 18, 16), (1100, 3, 0), (1100, 17, 16),
                        (1100, 5, 32)]:
    Q = np.triu(rng.uniform(-1, 1, (n, n)))
    h, J, _ = qbm_b200.ising.qubo_to_ising(Q)
    betas, spb = qbm_b200.ising.beta_schedule(qbm_b200.ising.default_beta_range(h, J), 30)
    Jd, hd, bd = (torch.from_numpy(a[0].astype(np.float32)).to(dev) for a in (J, h, betas))
    res = qbm_b200.sa_sample(Jd, hd, bd, spb, reads, 7, count=True, flags=flags)
    e = qbm_b200.qubo_energies(torch.from_numpy(Q).to(dev), res.states)
    qbm_b200.phase_stats(res.states)
    torch.cuda.synchronize()
    print("sa", n, reads, flags, float(e.mean()))
np.random.seed(19)
m = qbm_b200.DiscQBM(dim_input=8, num_classes=3, use_one_hot_encoding=True, n_hidden_nodes=5, sample_count=16, anneal_steps=50,
                     seed=19, stats_mode="loop", device=dev)
print("disc", m.train_for_one_iteration(rng.random((5, 8)), np.eye(3)[rng.integers(0, 3, 5)], 0.1)[1])
c = qbm_b200.ConvDeepQBM(100, 1, image_shape=(10, 10), kernel_size=3, pooling_size=2, sequential_layer_sizes=[6],
                         hidden_bias_type="shared", anneal=50, seed=4, device=dev)
print("convdeep", c.train_one_iteration(rng.random((3, 10, 10)).astype(np.float32), np.array([0, 1, 1]), 16, 1.0, 0.05))
r = qbm_b200.B200ClassificationRBM(70, 50, 1, num_classes=4, seed=3, device=dev)
xb = (rng.random((20, 70)) < 0.3).astype(np.float32)
yb = rng.integers(0, 4, 20)
for _ in range(3):                                   # the second step of a shape is captured, the third replayed as a CUDA graph
    print("rbm", float(r.discriminative_training(xb, yb)[0]))
for _ in range(3):
    r.cd1_training(xb, yb)
torch.cuda.synchronize()
print("done")
