"""Micro-benchmark of one pw_gemm shape through the op-level C ABI: CUDA-event time of N back-to-back launches
(warm L2).  Env switches are read by gemm_plan(): DCVC_B200_GEMM_MODE=stream|resident, DCVC_B200_GEMM_BN=64..256,
DCVC_B200_GEMM_DBG=1 (no MMA) | 2 (no epilogue) | 3 (loads only)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dcvc_b200 import ops  # noqa: E402

H, W, K, N = [int(v) for v in sys.argv[1:5]] if len(sys.argv) > 4 else (136, 240, 384, 384)
act = int(sys.argv[5]) if len(sys.argv) > 5 else 0
chunk = int(sys.argv[6]) if len(sys.argv) > 6 else 0
res = int(sys.argv[7]) if len(sys.argv) > 7 else 0
g = torch.Generator().manual_seed(0)
x = (torch.randn(H, W, K, generator=g) * 0.5).half().cuda()
w = (torch.randn(N, K, 1, 1, generator=g) * K ** -0.5)
wp = ops.pack_weight(ops.GEMM_PW, w)
b = torch.zeros(N).half().cuda()
Co = N // 4 if chunk else N
out = torch.zeros(H, W, Co, dtype=torch.float16, device="cuda")
r1 = (torch.randn(H, W, Co, generator=g)).half().cuda() if res else None
s = torch.cuda.Stream()
n = 20
with torch.cuda.stream(s):
    for _ in range(3):
        ops.gemm(ops.GEMM_PW, x, wp, N, out, bias=b, act=act, chunk_add=bool(chunk), res1=r1)
    torch.cuda.synchronize()
    # n launches in one CUDA graph: the host cost of planning + launching (~10 us through ctypes) stays out of the
    # measurement, exactly as in the codec (whole segments are graphs)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s):
        for _ in range(n):
            ops.gemm(ops.GEMM_PW, x, wp, N, out, bias=b, act=act, chunk_add=bool(chunk), res1=r1)
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / (n * reps)
fl = 2.0 * H * W * K * N
print(f"M={H*W} K={K} N={N} act={act} chunk={chunk} res={res} ares={os.environ.get('DCVC_B200_GEMM_ARES','1')} "
      f"pair={os.environ.get('DCVC_B200_GEMM_PAIR','1')} bn={os.environ.get('DCVC_B200_GEMM_BN','auto')} "
      f"dbg={os.environ.get('DCVC_B200_GEMM_DBG','0')}: {us:.1f} us  {fl/us/1e6:.0f} TFLOP/s")
