#!/usr/bin/env python
"""K4 micro-benchmark: the tcgen05 TF32 GEMM on the ultra-wide (config 4) and Reddit (config 3)
layer shapes, CUDA-event timed, with cuBLAS (TF32 and fp32) beside it.  Prints one JSON line per
shape.  TF32 tensor peak is taken as half the measured dense bf16 peak (MEASURED_PEAKS.json)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gist_b200 import ops  # noqa: E402


def timeit(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    bf16 = json.load(open(pk))['bf16_tflops'] if os.path.exists(pk) else 1590.0
    peak = bf16 / 2
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')   # > 126 MB L2
    n = 2280
    shapes = [
        # (name, M, N, K, a_mn, b_mn)
        ('cfg4 mid fwd  y=zW^T', n, 4096, 8192, False, False),
        ('cfg4 mid dz   dyW', n, 8192, 4096, False, True),
        ('cfg4 mid dW   dy^Tz', 4096, 8192, n, True, True),
        ('cfg4 l0  fwd', n, 4096, 200, False, False),
        ('cfg4 out fwd', n, 47, 8192, False, False),
        ('cfg3 l0 fwd', 2586, 256, 1204, False, False),
        ('cfg3 l0 dW', 256, 1204, 2586, True, True),
        ('cfg3 l1 dz', 2586, 512, 256, False, True),
        ('cfg3 l1 dW', 256, 512, 2586, True, True),
        ('square 8192', 8192, 8192, 8192, False, False),
    ]
    only = sys.argv[1] if len(sys.argv) > 1 else ''
    for name, M, N, K, a_mn, b_mn in shapes:
        if only and only not in name:
            continue
        A = torch.randn((K, M) if a_mn else (M, K), device='cuda')
        B = torch.randn((K, N) if b_mn else (N, K), device='cuda')
        out = torch.empty(M, N, device='cuda')
        Al = A.t() if a_mn else A
        Bl = B.t() if b_mn else B
        res = {'shape': name, 'M': M, 'N': N, 'K': K, 'a_mn': a_mn, 'b_mn': b_mn}
        fl = 2.0 * M * N * K
        variants = {'auto': 0}
        if M * N >= 1 << 22:
            variants.update({'bn128': 8, 'bn256': 16})
        for vn, f in variants.items():
            t = timeit(lambda: ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, out=out, flags=f), 10, flush)
            res['gist_%s_us' % vn] = round(t * 1e3, 1)
            res['gist_%s_tflops' % vn] = round(fl / t / 1e9, 1)
        res['frac_of_tf32_peak'] = round(res['gist_auto_tflops'] / peak, 3)
        # 3xTF32 (fp32-accurate): operands pre-split; the split launches are timed separately
        A_lo, B_lo = ops.split_tf32(A), ops.split_tf32(B)
        v3 = {'auto': 0, 'kc2': 2 << 8, 'kc8': 8 << 8, 'kc16': 16 << 8}
        if M * N >= 1 << 22:
            v3.update({'bn64': 4, 'bn128': 8})
        for vn, f in v3.items():
            t = timeit(lambda: ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, out=out, flags=f, A_lo=A_lo, B_lo=B_lo),
                       10, flush)
            res['gist3x_%s_us' % vn] = round(t * 1e3, 1)
            res['gist3x_%s_tflops' % vn] = round(fl / t / 1e9, 1)          # useful (fp32-equivalent) FLOP
        res['gist3x_mma_frac_of_tf32_peak'] = round(3 * res['gist3x_auto_tflops'] / peak, 3)
        t = timeit(lambda: (ops.split_tf32(A), ops.split_tf32(B)), 10, flush)
        res['split_both_us'] = round(t * 1e3, 1)
        ref = Al[:256].double() @ Bl.double().t()
        ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, out=out, A_lo=A_lo, B_lo=B_lo)
        res['rel_err3x_vs_fp64'] = float(((out[:256].double() - ref).norm() / ref.norm()).item())
        torch.backends.cuda.matmul.allow_tf32 = True
        t = timeit(lambda: torch.matmul(Al, Bl.t(), out=out), 10, flush)
        res['cublas_tf32_us'], res['cublas_tf32_tflops'] = round(t * 1e3, 1), round(fl / t / 1e9, 1)
        torch.backends.cuda.matmul.allow_tf32 = False
        t = timeit(lambda: torch.matmul(Al, Bl.t(), out=out), 5, flush)
        res['cublas_fp32_us'], res['cublas_fp32_tflops'] = round(t * 1e3, 1), round(fl / t / 1e9, 1)
        torch.matmul(Al, Bl.t(), out=out)
        res['rel_err_cublas_fp32_vs_fp64'] = float(((out[:256].double() - ref).norm() / ref.norm()).item())
        ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, out=out)
        res['rel_err_vs_fp64'] = float(((out[:256].double() - ref).norm() / ref.norm()).item())
        res['tf32_peak_tflops'] = peak
        print(json.dumps(res), flush=True)


if __name__ == '__main__':
    main()
