#!/usr/bin/env python
"""Where does the time of a SMALL K4 launch go?  For every GEMM of the Reddit-shape training step
(and a few config-4 ones) this runs the kernel warm (operands L2-resident, as inside the step) with
the per-CTA phase trace on (gist_gemm_set_trace) and prints, per shape: the launch's plan, the
median / max over CTAs of each phase in microseconds, the spread of CTA start times, and the
per-launch time of 20 back-to-back launches replayed from a CUDA graph.  Diagnostic only."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gist_b200 import _lib, ops  # noqa: E402


def graph_time(fn, reps=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
    return best


def main():
    lib = _lib.load()
    slots = lib.gist_gemm_trace_slots()
    nsm = torch.cuda.get_device_properties(0).multi_processor_count
    trace = torch.zeros(nsm * slots, dtype=torch.int64, device='cuda')
    n = 2590
    shapes = [
        # name, M, N, K, a_mn, b_mn, x3
        ('r3 L0 fwd', n, 256, 1204, False, False, True),
        ('r3 L1 fwd', n, 256, 512, False, False, True),
        ('r3 L2 fwd', n, 41, 512, False, False, True),
        ('r3 dz2', n, 512, 41, False, True, True),
        ('r3 dz1', n, 512, 256, False, True, True),
        ('r3 dW2', 41, 512, n, True, True, True),
        ('r3 dW1', 256, 512, n, True, True, True),
        ('r3 dW0', 256, 1204, n, True, True, True),
        ('m8 L0 fwd', n, 32, 1204, False, False, True),
        ('m8 L1 fwd', n, 32, 64, False, False, True),
        ('tiny K=32', n, 128, 32, False, False, True),
        ('tiny K=32 1x', n, 128, 32, False, False, False),
        ('r3 L1 fwd 1x', n, 256, 512, False, False, False),
        ('cfg4 mid fwd', 2280, 4096, 8192, False, False, True),
    ]
    only = sys.argv[1] if len(sys.argv) > 1 else ''
    ldk = lambda c: (c + 3) // 4 * 4  # noqa: E731
    for name, M, N, K, a_mn, b_mn, x3 in shapes:
        if only and only not in name:
            continue
        A = torch.randn((K, ldk(M)) if a_mn else (M, ldk(K)), device='cuda')[:, :(M if a_mn else K)]
        B = torch.randn((K, ldk(N)) if b_mn else (N, ldk(K)), device='cuda')[:, :(N if b_mn else K)]
        out = torch.empty(M, ldk(N), device='cuda')[:, :N]
        A_lo, B_lo = (ops.split_tf32(A), ops.split_tf32(B)) if x3 else (None, None)
        for flags, tag in ((0, 'auto'), (2, 'nosplit')):
            tn, sp, kbps = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
            lib.gist_gemm_plan(M, N, K, flags, 1 if x3 else 0, ctypes.byref(tn), ctypes.byref(sp), ctypes.byref(kbps))
            fn = lambda: ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, out=out, flags=flags, A_lo=A_lo, B_lo=B_lo)  # noqa: E731
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            lib.gist_gemm_set_trace(ctypes.c_void_p(trace.data_ptr()), nsm)
            trace.zero_()
            torch.cuda.synchronize()
            fn()
            torch.cuda.synchronize()
            lib.gist_gemm_set_trace(None, 0)
            t = trace.view(nsm, slots).cpu()
            tiles = ((M + 127) // 128) * ((N + tn.value - 1) // tn.value) * sp.value
            ncta = min(tiles, nsm)
            t = t[:ncta].double()
            wall = (t[:, 9] - t[:, 8])                                   # ns per CTA
            cyc = (t[:, 7] - t[:, 0]).clamp_min(1)
            ghz = float((cyc / wall.clamp_min(1)).median())
            us = lambda a, b: ((t[:, b] - t[:, a]) / (ghz * 1e3))       # noqa: E731
            ph = {'setup': us(0, 1), 'first_tma_issue': us(1, 2), 'first_stage_land': us(2, 3),
                  'mainloop': us(3, 4), 'mma_tail_to_epi': us(4, 5), 'epilogue': us(5, 6), 'exit': us(6, 7),
                  'cta_total': us(0, 7)}
            start_spread = float((t[:, 8].max() - t[:, 8].min()) / 1e3)
            span = float((t[:, 9].max() - t[:, 8].min()) / 1e3)
            per_launch = graph_time(fn)
            print(json.dumps({'shape': name, 'M': M, 'N': N, 'K': K, 'x3': x3, 'variant': tag, 'tile_n': tn.value,
                              'splits': sp.value, 'kb_per_split': kbps.value, 'ctas': ncta, 'sm_ghz': round(ghz, 3),
                              'phase_us_median': {k: round(float(v.median()), 2) for k, v in ph.items()},
                              'phase_us_max': {k: round(float(v.max()), 2) for k, v in ph.items()},
                              'cta_start_spread_us': round(start_spread, 2), 'kernel_span_us': round(span, 2),
                              'graph_us_per_call': round(per_launch, 2), 'launches_per_call': 2 if sp.value > 1 else 1}),
                  flush=True)
            if sp.value == 1:
                break


if __name__ == '__main__':
    main()
