"""Where the warp roles of the chain-tile kernel wait: builds an instrumented copy of the library (-DQBM_TILE_PROF, clock()
around every hand-off) next to the shipped one, runs the first S sweeps of the C4 schedule and prints the share of the time
each role spends on each wait.  Run on the GPU box: python tools/probe_tile_prof.py [--n 2048] [--cuts 100]"""
import argparse
import ctypes
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("_b", os.path.join(ROOT, "qbm-image-classification_b200", "build.py"))
_b = importlib.util.module_from_spec(spec)
spec.loader.exec_module(_b)
_prof = os.path.join(ROOT, "qbm-image-classification_b200", "libqbm_b200_prof.so")
if not os.path.exists(_prof) or os.path.getmtime(_prof) < os.path.getmtime(os.path.join(_b.CSRC, "sa_tile.cu")):
    _b.build(variant="prof", extra=("-DQBM_TILE_PROF",))
os.environ["QBM_B200_LIB"] = _prof

import numpy as np
import torch
import qbm_b200


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2048)
    ap.add_argument("--reads", type=int, default=2368)
    ap.add_argument("--cuts", default="100")
    ap.add_argument("--start", type=int, default=0, help="first sweep of the window (the kernel runs sweeps 0..cut; the counters cover all of them)")
    ap.add_argument("--flags", type=lambda x: int(x, 0), default=16)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    L = qbm_b200._lib.load()
    L.qbm_debug_tile_prof.restype = ctypes.c_int
    L.qbm_debug_tile_prof.argtypes = [ctypes.c_void_p, ctypes.c_int]
    rng = np.random.default_rng(19)
    Q = np.triu(rng.uniform(-1, 1, (a.n, a.n)))[None]
    h, J, _ = qbm_b200.ising.qubo_to_ising(Q)
    betas, spb = qbm_b200.ising.beta_schedule(qbm_b200.ising.default_beta_range(h, J), 1000)
    Jd = torch.from_numpy(J.astype(np.float32)).to(dev)
    hd = torch.from_numpy(h.astype(np.float32)).to(dev)
    for S in [int(x) for x in a.cuts.split(",")]:
        bd = torch.from_numpy(np.ascontiguousarray(betas[:, :S]).astype(np.float32)).to(dev)
        qbm_b200.sa_sample(Jd, hd, bd, spb, a.reads, 19, count=True, flags=a.flags)       # warm-up
        L.qbm_debug_tile_prof(None, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        qbm_b200.sa_sample(Jd, hd, bd, spb, a.reads, 19, count=True, flags=a.flags)
        e1.record()
        torch.cuda.synchronize()
        out = (ctypes.c_ulonglong * 32)()
        L.qbm_debug_tile_prof(out, 1)
        v = [float(x) for x in out]
        pc = lambda x, tot: f"{100 * x / max(tot, 1):5.1f}%"
        print(f"n={a.n} first {S} sweeps: {e0.elapsed_time(e1):.1f} ms")
        print(f"  appliers (all warps): ring-full {pc(v[1], v[0])}  record-full {pc(v[2], v[0])}  export-buffer {pc(v[3], v[0])}")
        print(f"  applier warp 0:       ring-full {pc(v[5], v[4])}  record-full {pc(v[6], v[4])}  export-buffer {pc(v[7], v[4])}")
        print(f"  scanner 0: fields-exported {pc(v[9], v[8])}  bounds {pc(v[10], v[8])}  record-buffer {pc(v[11], v[8])}  "
              f"cp.async {pc(v[12], v[8])}  | catch-up {pc(v[13], v[8])}  scan {pc(v[14], v[8])}")
        print(f"  producer:  record {pc(v[17], v[16])}  slot-empty {pc(v[18], v[16])}", flush=True)


if __name__ == "__main__":
    main()
