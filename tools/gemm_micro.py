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
with torch.cuda.stream(s):
    for _ in range(5):
        ops.gemm(ops.GEMM_PW, x, wp, N, out, bias=b, act=act, chunk_add=bool(chunk), res1=r1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for _ in range(n):
        ops.gemm(ops.GEMM_PW, x, wp, N, out, bias=b, act=act, chunk_add=bool(chunk), res1=r1)
    e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / n
fl = 2.0 * H * W * K * N
print(f"M={H*W} K={K} N={N} act={act} chunk={chunk} res={res} mode={os.environ.get('DCVC_B200_GEMM_MODE','auto')} "
      f"bn={os.environ.get('DCVC_B200_GEMM_BN','auto')} dbg={os.environ.get('DCVC_B200_GEMM_DBG','0')}: "
      f"{us:.1f} us  {fl/us/1e6:.0f} TFLOP/s")
