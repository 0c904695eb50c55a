"""Timeline of CTA 0 of one fused dcb_tail launch (globaltimer marks, DCVC_B200_GEMM_TRACE): per chunk, when the MMA thread
could start it / got its first weights / finished issuing it, and when epilogue warp 2 started waiting, got the
accumulator, handed it back, finished the body and signalled.
  python tools/dcb_tail_trace.py H W C inner inner_next"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
H, W, C, inner, inner_n = [int(v) for v in sys.argv[1:6]] if len(sys.argv) > 5 else (136, 240, 384, 384, 384)
trace = torch.zeros(2048, dtype=torch.int64, device="cuda")
from dcvc_b200 import ops  # noqa: E402

g = torch.Generator().manual_seed(0)
rnd = lambda *s, sc=0.5: (torch.randn(*s, generator=g) * sc).half().cuda()  # noqa: E731
t2, x = rnd(H, W, inner), rnd(H, W, C)
w3, wf0, wf2 = rnd(C, inner, sc=inner ** -0.5), rnd(4 * inner, C, sc=C ** -0.5), rnd(C, inner, sc=inner ** -0.5)
w0n = rnd(inner_n, C, sc=C ** -0.5) if inner_n else None
b3, bf0, bf2 = rnd(C, sc=0.1), rnd(4 * inner, sc=0.1), rnd(C, sc=0.1)
b0n = rnd(inner_n, sc=0.1) if inner_n else None
y = torch.zeros(H, W, C, dtype=torch.float16, device="cuda")
t1n = torch.zeros(H, W, inner_n, dtype=torch.float16, device="cuda") if inner_n else None


def run():
    assert ops.dcb_tail(t2, x, y, w3, b3, wf0, bf0, wf2, bf2, t1n=t1n, w0n=w0n, b0n=b0n)


s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(20):     # warm clocks and caches
        run()
    torch.cuda.synchronize()
    os.environ["DCVC_B200_GEMM_TRACE"] = hex(trace.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        run()
    e0.record()
    run()
    e1.record()
torch.cuda.synchronize()
t = trace.cpu().numpy().astype(np.int64)
nz = t[t > 0]
t0 = nz.min()
us = lambda v: (v - t0) / 1000.0 if v > 0 else float("nan")  # noqa: E731
nch = [C // 128, 4 * inner // 128, C // 128, inner_n // 128]
per_tile = sum(nch)
print(f"M={H*W} C={C} inner={inner} next={inner_n}: event time {e0.elapsed_time(e1) * 1e3:.1f} us; span of marks {(nz.max() - t0) / 1000.0:.1f} us")
print("chunk ph |  mma: free   w0-land  issued | epi: start  wait    acc     handed  body    signalled | epi total  mma issue")
c = 0
tile = 0
while True:
    for ph in range(4):
        for n in range(nch[ph]):
            m = [us(t[512 + 4 * c + k]) for k in range(3)]
            e = [us(t[8 * c + k]) for k in (0, 5, 1, 2, 3, 4)]
            if np.isnan(m[0]) and np.isnan(e[0]):
                break
            print(f"{c:5d} {ph + 1}  | {m[0]:9.2f} {m[1]:9.2f} {m[2]:8.2f} | {e[0]:9.2f} {e[1]:7.2f} {e[2]:7.2f} {e[3]:7.2f} {e[4]:7.2f} {e[5]:9.2f} |"
                  f" {e[5] - e[2]:8.2f} {m[2] - m[0]:9.2f}")
            c += 1
    tile += 1
    if tile >= 2 or c >= 60:
        break
st = np.array([us(v) for v in t[1024:1024 + 2 * 126]])
st = st[~np.isnan(st)]
if len(st) > 2:
    d = np.diff(st)
    print(f"producer: {len(st)} stages issued between {st[0]:.2f} and {st[-1]:.2f} us; gap median {np.median(d):.3f} us, p90 {np.percentile(d, 90):.3f}, max {d.max():.3f}")
