"""Micro-benchmark of the fused DepthConvBlock tail (dcvc_op_dcb_tail) against the per-op launches it replaces, both as
CUDA graphs of n back-to-back launches (warm L2).  DCVC_B200_GEMM_DBG=1 (no MMA) | 2 (no epilogue bodies) | 3 (loads only).
  python tools/dcb_tail_micro.py H W C inner inner_next"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dcvc_b200 import ops  # noqa: E402

H, W, C, inner, inner_n = [int(v) for v in sys.argv[1:6]] if len(sys.argv) > 5 else (136, 240, 384, 384, 384)
g = torch.Generator().manual_seed(0)
rnd = lambda *s, sc=0.5: (torch.randn(*s, generator=g) * sc).half().cuda()  # noqa: E731
t2, x = rnd(H, W, inner), rnd(H, W, C)
w3, wf0, wf2 = rnd(C, inner, sc=inner ** -0.5), rnd(4 * inner, C, sc=C ** -0.5), rnd(C, inner, sc=inner ** -0.5)
w0n = rnd(inner_n, C, sc=C ** -0.5) if inner_n else None
b3, bf0, bf2 = rnd(C, sc=0.1), rnd(4 * inner, sc=0.1), rnd(C, sc=0.1)
b0n = rnd(inner_n, sc=0.1) if inner_n else None
y = torch.zeros(H, W, C, dtype=torch.float16, device="cuda")
o = torch.zeros_like(y)
t1 = torch.zeros(H, W, inner, dtype=torch.float16, device="cuda")
t1n = torch.zeros(H, W, inner_n, dtype=torch.float16, device="cuda") if inner_n else None


def fused():
    assert ops.dcb_tail(t2, x, y, w3, b3, wf0, bf0, wf2, bf2, t1n=t1n, w0n=w0n, b0n=b0n)


def per_op():
    ops.gemm(ops.GEMM_PW, t2, w3, C, o, bias=b3, res1=x)
    ops.gemm(ops.GEMM_PW, o, wf0, 4 * inner, t1, bias=bf0, act=ops.ACT_WSILU, chunk_add=True)
    ops.gemm(ops.GEMM_PW, t1, wf2, C, y, bias=bf2, res1=o)
    if inner_n:
        ops.gemm(ops.GEMM_PW, y, w0n, inner_n, t1n, bias=b0n, act=ops.ACT_WSILU)


def time_graph(fn, n=10, reps=5):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            for _ in range(n):
                fn()
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            graph.replay()
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (n * reps)


fl = 2.0 * H * W * (C * inner + 4 * inner * C + C * inner + (inner_n * C if inner_n else 0))
which = sys.argv[6] if len(sys.argv) > 6 else "both"
tag = f"M={H*W} C={C} inner={inner} next={inner_n} dbg={os.environ.get('DCVC_B200_GEMM_DBG', '0')}"
if which in ("both", "fused"):
    us = time_graph(fused)
    print(f"{tag} fused : {us:.1f} us  {fl / us / 1e6:.0f} TFLOP/s")
if which in ("both", "perop"):
    us = time_graph(per_op)
    print(f"{tag} per-op: {us:.1f} us  {fl / us / 1e6:.0f} TFLOP/s")
