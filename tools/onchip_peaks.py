"""Measure the on-chip peaks of the GPU (FFMA2 / FFMA TFLOP/s, shared-memory and L1 TB/s) through
qbm_probe_onchip_peaks and print them as one JSON line (bench.py does the same inside its run)."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import qbm_b200  # noqa: E402


def measure(device="cuda:0"):
    L = qbm_b200._lib.load()
    torch.cuda.set_device(device)
    scratch = torch.empty((1 << 20) + 64, dtype=torch.uint8, device=device)
    out = (ctypes.c_double * 6)()
    qbm_b200._lib.check(L.qbm_probe_onchip_peaks(out, scratch.data_ptr(), scratch.numel(),
                                                 torch.cuda.current_stream().cuda_stream))
    return {"fp32_ffma2_tflops": out[0], "fp32_ffma_tflops": out[1], "smem_lds128_tbs": out[2], "l1_ldg128_tbs": out[3],
            "fp32_ffma2_3reg_8warps_cmajor_tflops": out[4], "fp32_ffma2_3reg_8warps_rmajor_tflops": out[5]}


if __name__ == "__main__":
    print(json.dumps(measure()))
