"""Run-to-run determinism + accuracy stress of pw_gemm / dw3x3 at codec-sized shapes (diagnostic, not a test)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dcvc_b200 import ops  # noqa: E402

torch.manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def ref_pw(x, w, b, act, chunk, r1, r2):
    y = x.float().reshape(-1, x.shape[-1]) @ w.float().reshape(w.shape[0], -1).t().cuda() + b.float()
    if act:
        y = y * torch.sigmoid(4.0 * y)
    if chunk:
        y = y.reshape(y.shape[0], -1, 4).sum(-1)
    y = y.reshape(x.shape[0], x.shape[1], -1)
    if r1 is not None:
        y = y + r1.float()
    if r2 is not None:
        y = y + r2.float()
    return y


def run(H, W, K, N, act=0, chunk=0, nres=0, in_pitch=None, out_pitch=None, reps=6):
    g = torch.Generator().manual_seed(1)
    ip = in_pitch or K
    xb = (torch.randn(H, W, ip, generator=g) * 0.5).half().cuda()
    x = xb[:, :, :K]
    w = (torch.randn(N, K, 1, 1, generator=g) * K ** -0.5).half()
    wp = ops.pack_weight(ops.GEMM_PW, w)
    b = (torch.randn(N, generator=g) * 0.1).half().cuda()
    Co = N // 4 if chunk else N
    op = out_pitch or Co
    r1 = torch.randn(H, W, Co, generator=g).half().cuda() if nres > 0 else None
    r2 = torch.randn(H, W, Co, generator=g).half().cuda() if nres > 1 else None
    outs = []
    for i in range(reps):
        ob = torch.full((H, W, op), 7.0, dtype=torch.float16, device="cuda")
        o = ob[:, :, :Co]
        flush.fill_(i)
        ops.gemm(ops.GEMM_PW, x, wp, N, o, bias=b, act=act, chunk_add=bool(chunk), res1=r1, res2=r2)
        torch.cuda.synchronize()
        outs.append(o.clone())
    ref = ref_pw(x, w, b, act, chunk, r1, r2)
    err = (outs[0].float() - ref).abs().max().item()
    nd = [int((o != outs[0]).sum().item()) for o in outs[1:]]
    print(f"pw H={H} W={W} K={K} N={N} act={act} chunk={chunk} res={nres} ip={ip} op={op}: max_err={err:.4f} "
          f"run-to-run diffs={nd}", flush=True)


def run_dw(H, W, Cc, reps=4):
    g = torch.Generator().manual_seed(2)
    x = (torch.randn(H, W, Cc, generator=g)).half().cuda()
    w = (torch.randn(9, Cc, generator=g) * 0.3).half().cuda()
    outs = []
    for i in range(reps):
        o = torch.zeros(H, W, Cc, dtype=torch.float16, device="cuda")
        flush.fill_(i)
        ops.dw3x3(x, w, o)
        torch.cuda.synchronize()
        outs.append(o.clone())
    ref = torch.nn.functional.conv2d(x.float().permute(2, 0, 1)[None], w.float().t().reshape(Cc, 1, 3, 3), padding=1,
                                     groups=Cc)[0].permute(1, 2, 0)
    err = (outs[0].float() - ref).abs().max().item()
    nd = [int((o != outs[0]).sum().item()) for o in outs[1:]]
    print(f"dw H={H} W={W} C={Cc}: max_err={err:.4f} run-to-run diffs={nd}", flush=True)


H8, W8, H16, W16 = 135, 240, 68, 120
run_dw(H8, W8, 512)
run_dw(H16, W16, 256)
for (K, N, act, chunk, nres, ip, op) in [
    (1024, 512, 0, 0, 0, 1024, None),   # adaptor on cat buffers
    (2048, 512, 0, 0, 0, 2048, None),
    (512, 512, 1, 0, 0, None, None),    # conv0 + wsilu
    (512, 512, 0, 0, 1, None, None),    # conv3 + residual
    (512, 2048, 1, 1, 0, None, None),   # ffn expand, chunk-add
    (512, 512, 0, 0, 1, None, 1024),    # ffn out into a cat slice
    (512, 512, 0, 0, 2, None, None),
    (512, 256, 0, 0, 0, 1024, None),
    (256, 256, 0, 0, 1, None, 2048),
]:
    run(H8, W8, K, N, act, chunk, nres, ip, op)
for (K, N, act, chunk, nres, ip, op) in [
    (768, 256, 0, 0, 0, 768, None),
    (256, 256, 1, 0, 0, None, None),
    (256, 1024, 1, 1, 0, None, None),
    (256, 256, 0, 0, 1, None, 512),
    (512, 256, 0, 0, 0, 512, None),
    (256, 768, 0, 0, 0, None, None),
]:
    run(H16, W16, K, N, act, chunk, nres, ip, op)
print("done")
