"""Per-CTA timeline of one pw_gemm launch (globaltimer marks written by the kernel when DCVC_B200_GEMM_TRACE is set)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
H, W, K, N = [int(v) for v in sys.argv[1:5]]
act = int(sys.argv[5]) if len(sys.argv) > 5 else 0
chunk = int(sys.argv[6]) if len(sys.argv) > 6 else 0
res = int(sys.argv[7]) if len(sys.argv) > 7 else 0
trace = torch.zeros(148 * 64, dtype=torch.int64, device="cuda")
os.environ["DCVC_B200_GEMM_TRACE"] = hex(trace.data_ptr())
from dcvc_b200 import ops  # noqa: E402

g = torch.Generator().manual_seed(0)
x = (torch.randn(H, W, K, generator=g) * 0.5).half().cuda()
w = (torch.randn(N, K, 1, 1, generator=g) * K ** -0.5)
wp = ops.pack_weight(ops.GEMM_PW, w)
b = torch.zeros(N).half().cuda()
Co = N // 4 if chunk else N
out = torch.zeros(H, W, Co, dtype=torch.float16, device="cuda")
r1 = torch.randn(H, W, Co, generator=g).half().cuda() if res else None
s = torch.cuda.Stream()
names = ["entry", "setup", "ld0_req", "ldN_req", "ld0_land", "mma_t0", "mma_last", "acc0", "epi0", "acc1", "epi1",
         "accL", "epiL", "drain0", "drain1"]
with torch.cuda.stream(s):
    for _ in range(5):
        ops.gemm(ops.GEMM_PW, x, wp, N, out, bias=b, act=act, chunk_add=bool(chunk), res1=r1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.gemm(ops.GEMM_PW, x, wp, N, out, bias=b, act=act, chunk_add=bool(chunk), res1=r1)
    e1.record()
torch.cuda.synchronize()
t = trace.cpu().numpy().reshape(148, 64).astype(np.int64)
used = t[:, 0] > 0
t = t[used]
t0 = t[:, 0].min()
rel = (t - t0) / 1000.0
rel[t == 0] = np.nan
print(f"M={H*W} K={K} N={N} act={act} chunk={chunk} res={res}: event time {e0.elapsed_time(e1)*1e3:.1f} us, CTAs {used.sum()}")
print("mark      " + " ".join(f"{n:>8}" for n in names))
for lab, fn in (("min", np.nanmin), ("median", np.nanmedian), ("max", np.nanmax)):
    print(f"{lab:<9} " + " ".join(f"{fn(rel[:, i]):8.2f}" for i in range(len(names))))
for c in (0, 1, 73, 147):
    if c < rel.shape[0]:
        print(f"cta{c:<6} " + " ".join(f"{rel[c, i]:8.2f}" for i in range(len(names))))

if os.environ.get("DCVC_B200_GEMM_ARES", "0") == "1":
    # A-resident kernel: per-phase SM-clock marks of the first tile's chunks, epilogue warp (q=0, h=0)
    raw = trace.cpu().numpy().reshape(148, 64).astype(np.int64)[used]
    ph = ["ldtm_issue", "ldtm_done", "math_done", "res_q_done", "sts_done", "store_issued"]
    for a in range(4):
        c = raw[:, 16 + a * 8:16 + a * 8 + 6]
        ok = (c > 0).all(axis=1)
        if not ok.any():
            continue
        d = np.diff(c[ok], axis=1)
        print(f"chunk {a}: " + "  ".join(f"{ph[k]}->{ph[k+1]} {np.median(d[:, k]):.0f}" for k in range(5)) +
              (f"  | chunk period {np.median(raw[ok, 16 + (a + 1) * 8] - raw[ok, 16 + a * 8]):.0f} clk" if a < 3 and (raw[ok, 16 + (a + 1) * 8] > 0).all() else ""))
    sys.exit(0)
# per-stage TMA timeline of the first 24 k-block loads in SM clocks (cheap clock64 marks kept in smem): request
# (producer thread) and landing as seen by the MMA thread, relative to the "setup done" mark
raw = trace.cpu().numpy().reshape(148, 64).astype(np.int64)[used]
c0 = raw[:, 15] & 0xffffffff
clk = ((raw[:, 16:64] & 0xffffffff) - c0[:, None]) & 0xffffffff
req, land = clk[:, :24].astype(float), clk[:, 24:].astype(float)
print("stage      " + " ".join(f"{i:>6}" for i in range(24)))
print("req  clk   " + " ".join(f"{np.median(req[:, i]):6.0f}" for i in range(24)))
print("land clk   " + " ".join(f"{np.median(land[:, i]):6.0f}" for i in range(24)))
print("lat  med   " + " ".join(f"{np.median(land[:, i] - req[:, i]):6.0f}" for i in range(24)))
