"""Run a few training steps of one benchmark config (for ncu launch lists): python tools/profile_train.py c1 [steps]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qbm_b200
import bench_train as BT

cfg = sys.argv[1] if len(sys.argv) > 1 else "c1"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mode = sys.argv[3] if len(sys.argv) > 3 else "disc"
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
ms, e2e, h2d = BT.gpu_train_rate(cfg, qbm_b200, torch, dev, 1, 0, torch.cuda.synchronize, BT.CONFIGS[cfg][1], steps, 2, mode=mode)
print(f"{cfg}: {ms / steps:.3f} ms/step device-resident, {1e3 * e2e / steps:.3f} ms/step end to end")
