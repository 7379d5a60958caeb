"""A/B timing of qbm_sa_sample under different flag words (kernel selection bits, see include/qbm_b200.h): checks that
every flag word returns the states of flags = 0, then prints G spin-updates/s (CUDA events, best of `iters`)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import qbm_b200


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="24,34,100,192,193,256,384,512,522,640,768,1024,2048")
    ap.add_argument("--flags", default="0")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--sweeps", type=int, default=1000)
    ap.add_argument("--reads", type=int, default=0, help="reads per problem for n > 768 (default 2368)")
    ap.add_argument("--small", default="64x200", help="problems x reads for n <= 768")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    flags = [int(x, 0) for x in a.flags.split(",")]
    for n in [int(x) for x in a.sizes.split(",")]:
        batch, reads = tuple(int(x) for x in a.small.split("x")) if n <= 768 else (1, a.reads or 2368)
        rng = np.random.default_rng(19)
        Q = np.stack([np.triu(rng.uniform(-1, 1, (n, n))) for _ in range(batch)])
        h, J, _ = qbm_b200.ising.qubo_to_ising(Q)
        betas, spb = qbm_b200.ising.beta_schedule(qbm_b200.ising.default_beta_range(h, J), a.sweeps)
        Jd = torch.from_numpy(J.astype(np.float32)).to(dev)
        hd = torch.from_numpy(h.astype(np.float32)).to(dev)
        bd = torch.from_numpy(betas.astype(np.float32)).to(dev)
        base, line = None, []
        for f in flags:
            best = 1e30
            for _ in range(a.iters):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                res = qbm_b200.sa_sample(Jd, hd, bd, spb, reads, 19, flags=f)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            if base is None:
                base = res.states.clone()
            ok = torch.equal(base, res.states)
            rate = batch * reads * n * betas.shape[-1] * spb / best / 1e6
            line.append(f"flags={f} {rate:.2f}{'' if ok else ' MISMATCH'}")
        print(f"n={n} ({batch}x{reads}): " + " | ".join(line), flush=True)


if __name__ == "__main__":
    main()
