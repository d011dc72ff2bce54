"""A/B timing of the tcgen05 input-projection GEMM epilogues at the c3 shapes: TMA-store (0) vs row-per-thread stores (1).

    python tools/gemm_ab.py          # las_debug_set_option(8, v) selects the epilogue
"""
import os
import sys

import torch

ROOT = os.environ.get("LAS_ROOT") or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from las_pytorch_b200 import _cabi  # noqa: E402

SHAPES = [("L0", 51200, 2048, 80), ("L1", 25600, 2048, 1024), ("L2", 12800, 2048, 1024), ("psi", 12800, 64, 512)]


def main():
    lib = _cabi.load_library()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name, M, N, K in SHAPES:
        g = torch.Generator().manual_seed(M + N + K)
        a = torch.randn(M, K, generator=g).cuda().to(torch.bfloat16)
        w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda().to(torch.bfloat16)
        bias = torch.randn(N, generator=g).cuda()
        c = torch.empty(M, N, device="cuda")
        st = _cabi.current_stream_ptr()
        for direct in (1, 0, 1, 0):
            lib.las_debug_set_option(8, direct)
            ts = []
            for it in range(8):
                flush.fill_(it)  # L2 flush between timed launches
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _cabi.check(lib.las_debug_gemm_bf16(_cabi.ptr(a), _cabi.ptr(w), _cabi.ptr(bias), _cabi.ptr(c), M, N, K, st))
                e1.record()
                torch.cuda.synchronize()
                if it >= 2:
                    ts.append(e0.elapsed_time(e1))
            t = sorted(ts)[len(ts) // 2]
            flops, byts = 2.0 * M * N * K, 2.0 * M * K + 2.0 * N * K + 4.0 * M * N
            print(f"{name:4s} M={M} N={N} K={K} epilogue={'direct' if direct else 'tma   '}: {t * 1e3:8.1f} us  {flops / t / 1e9:7.1f} TFLOP/s  {byts / t / 1e6:7.1f} GB/s")
        lib.las_debug_set_option(8, 0)


if __name__ == "__main__":
    main()
