"""Time the first S sweeps of the default 1000-sweep schedule (the hot phase, where nearly every proposal is accepted) for
different kernels (flag words: 64 = warp-per-chain kernel alone, 16 = chain-tile kernel, 32 = chains-per-warp kernel, 0 = the
default, i.e. the two-phase schedule at n > 1792): G spin-updates/s, accepted fraction and clocks per accepted flip per SM."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import qbm_b200


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2048)
    ap.add_argument("--reads", type=int, default=2368)
    ap.add_argument("--cuts", default="50,100,150,200,300,1000")
    ap.add_argument("--flags", default="64,16,0")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(19)
    Q = np.triu(rng.uniform(-1, 1, (a.n, a.n)))[None]
    h, J, _ = qbm_b200.ising.qubo_to_ising(Q)
    betas, spb = qbm_b200.ising.beta_schedule(qbm_b200.ising.default_beta_range(h, J), 1000)
    Jd = torch.from_numpy(J.astype(np.float32)).to(dev)
    hd = torch.from_numpy(h.astype(np.float32)).to(dev)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    for S in [int(x) for x in a.cuts.split(",")]:
        bd = torch.from_numpy(np.ascontiguousarray(betas[:, :S]).astype(np.float32)).to(dev)
        line = []
        for f in [int(x, 0) for x in a.flags.split(",")]:
            best, res = 1e30, None
            for _ in range(2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                res = qbm_b200.sa_sample(Jd, hd, bd, spb, a.reads, 19, count=True, flags=f)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            acc, prop = [int(x) for x in res.accepted.cpu().numpy()]
            clk = best * 1e-3 * 1.965e9 * sms / acc
            line.append(f"flags={f}: {best:7.1f} ms {prop / best / 1e6:6.2f} G/s acc {acc / prop:.3f} {clk:6.1f} clk/flip/SM")
        print(f"first {S:4d} sweeps: " + " | ".join(line), flush=True)


if __name__ == "__main__":
    main()
