"""Scratch timing probe for the SA kernel (CUDA events on the launching stream)."""
import argparse
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import qbm_b200


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2048)
    ap.add_argument("--reads", type=int, default=2368)
    ap.add_argument("--sweeps", type=int, default=1000)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--iters", type=int, default=2)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(19)
    Q = np.stack([np.triu(rng.uniform(-1, 1, (a.n, a.n))) for _ in range(a.batch)])
    h, J, _ = qbm_b200.ising.qubo_to_ising(Q)
    betas, spb = qbm_b200.ising.beta_schedule(qbm_b200.ising.default_beta_range(h, J), a.sweeps)
    Jd = torch.from_numpy(J.astype(np.float32)).to(dev)
    hd = torch.from_numpy(h.astype(np.float32)).to(dev)
    bd = torch.from_numpy(betas.astype(np.float32)).to(dev)
    for it in range(a.iters + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        res = qbm_b200.sa_sample(Jd, hd, bd, spb, a.reads, 19 + it, count=True, flags=a.flags)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        acc, prop = [int(x) for x in res.accepted.cpu().numpy()]
        print(f"n={a.n} batch={a.batch} reads={a.reads} sweeps={a.sweeps} flags={a.flags}: {ms:.1f} ms  "
              f"{prop / ms / 1e6:.3f} G spin-updates/s  accepted {acc / prop:.4f}  flips/s {acc / ms / 1e6:.3f} G  "
              f"row-bytes {acc * 4 * a.n / ms / 1e9:.2f} TB/s", flush=True)
    Qd = torch.from_numpy(Q).to(dev)
    e = qbm_b200.qubo_energies(Qd, res.states)
    print("mean energy", e.mean().item(), "min", e.min().item())


if __name__ == "__main__":
    main()
