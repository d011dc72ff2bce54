"""Runs the UMMA probe / tcgen05 GEMM over layout variants on the GPU and prints max errors (diagnostic)."""
import itertools
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from las_pytorch_b200 import _cabi  # noqa: E402


def main():
    lib = _cabi.load_library()
    dev = torch.device("cuda")
    st = _cabi.current_stream_ptr()
    g = torch.Generator(device="cpu").manual_seed(0)
    for N, K in [(16, 64), (16, 256), (64, 128), (128, 64)]:
        a = torch.randn(128, K, generator=g).to(dev).to(torch.bfloat16)
        b = torch.randn(N, K, generator=g).to(dev).to(torch.bfloat16)
        ref = a.float() @ b.float().t()
        for a_sw, b_sw, variant in [(0, 0, 0), (1, 1, 0), (0, 0, 2), (0, 1, 2)]:
            if variant and a_sw and b_sw:
                continue
            d = torch.full((128, N), float("nan"), device=dev)
            rc = lib.las_debug_umma_probe(_cabi.ptr(a), _cabi.ptr(b), _cabi.ptr(d), N, K, a_sw, b_sw, variant, st)
            torch.cuda.synchronize()
            err = float((d - ref).abs().max()) if rc == 0 else float("nan")
            print(f"probe N={N:3d} K={K:3d} a_sw128={a_sw} b_sw128={b_sw} variant={variant} rc={rc} max_err={err:.4g} "
                  f"(ref max {float(ref.abs().max()):.3g})", flush=True)
    for M, N, K in [(128, 256, 64), (300, 128, 80), (1000, 2048, 1024), (389, 280, 96), (25600, 2048, 1024)]:
        a = torch.randn(M, K, generator=g).to(dev).to(torch.bfloat16)
        w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev).to(torch.bfloat16)
        bias = torch.randn(N, generator=g).to(dev)
        c = torch.full((M, N), float("nan"), device=dev)
        rc = lib.las_debug_gemm_bf16(_cabi.ptr(a), _cabi.ptr(w), _cabi.ptr(bias), _cabi.ptr(c), M, N, K, st)
        torch.cuda.synchronize()
        ref = a.float() @ w.float().t() + bias
        err = float((c - ref).abs().max()) if rc == 0 else float("nan")
        msg = "" if rc == 0 else lib.las_last_error().decode()
        t = ""
        if rc == 0 and M >= 1000:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(3):
                lib.las_debug_gemm_bf16(_cabi.ptr(a), _cabi.ptr(w), _cabi.ptr(bias), _cabi.ptr(c), M, N, K, st)
            e0.record()
            for _ in range(10):
                lib.las_debug_gemm_bf16(_cabi.ptr(a), _cabi.ptr(w), _cabi.ptr(bias), _cabi.ptr(c), M, N, K, st)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            t = f" {ms * 1e3:.1f} us  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s"
        print(f"gemm M={M} N={N} K={K} rc={rc} max_err={err:.4g}{t} {msg}", flush=True)


if __name__ == "__main__":
    main()
