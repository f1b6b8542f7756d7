"""Measure the FP64 roofline denominators on the GPU box.

cuBLAS DGEMM (torch.matmul) is used ONLY as the measuring stick for the FP64 tensor peak; it is
not on any product path.  Writes profiles/peaks_fp64.json (copied from gpurun_out/)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(iters):
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best


def main():
    out = {"gpu": torch.cuda.get_device_name(0)}
    for n in (4096, 8192):
        a = torch.randn(n, n, dtype=torch.float64, device="cuda")
        b = torch.randn(n, n, dtype=torch.float64, device="cuda")
        ms = timed(lambda: torch.matmul(a, b), 5)
        out["cublas_dgemm_%d_tflops" % n] = 2.0 * n ** 3 / ms / 1e9
    # sustained: back-to-back for ~3 s
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    t0 = time.time()
    cnt = 0
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    while time.time() - t0 < 3.0:
        torch.matmul(a, b)
        cnt += 1
        if cnt % 4 == 0:
            torch.cuda.synchronize()
    e.record()
    torch.cuda.synchronize()
    out["cublas_dgemm_8192_sustained_tflops"] = cnt * 2.0 * n ** 3 / s.elapsed_time(e) / 1e9
    # cuSOLVER potrf as a second yardstick for the Cholesky
    n = 16384
    x = torch.randn(n, n, dtype=torch.float64, device="cuda")
    spd = x @ x.T / n + torch.eye(n, dtype=torch.float64, device="cuda")
    ms = timed(lambda: torch.linalg.cholesky(spd), 2, warm=1)
    out["cusolver_potrf_16384_ms"] = ms
    out["cusolver_potrf_16384_tflops"] = n ** 3 / 3.0 / ms / 1e9
    del x, spd
    # our own DMMA GEMM shapes through the library: lauum = N^3/3 flops
    from multifidelity_datafusion_gps_b200 import ops
    for n in (4096, 8192, 16384):
        W = torch.tril(torch.randn(n, n, dtype=torch.float64, device="cuda"))
        ms = timed(lambda: ops.lauum(W), 3, warm=1)
        out["mfgp_lauum_%d_ms" % n] = ms
        out["mfgp_lauum_%d_tflops" % n] = n ** 3 / 3.0 / ms / 1e9
        del W
    # HBM copy
    src = torch.empty(1 << 28, dtype=torch.float64, device="cuda")
    dst = torch.empty_like(src)
    ms = timed(lambda: dst.copy_(src), 5)
    out["hbm_copy_gbs"] = 2 * src.numel() * 8 / ms / 1e6
    print(json.dumps(out, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/peaks_fp64.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
