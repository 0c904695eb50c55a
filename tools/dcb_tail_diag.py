"""Diagnosis: fused dcb_tail vs the per-op kernels at one shape; prints which 128-row tiles / channels disagree.
  python tools/dcb_tail_diag.py H W C inner inner_next [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_dcb_tail_gpu as T  # noqa: E402

H, W, C, inner, inner_n = [int(v) for v in sys.argv[1:6]]
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
inplace = len(sys.argv) > 7 and sys.argv[7] == "inplace"
gen = torch.Generator().manual_seed(1)
d = T._make(gen, H, W, C, inner, inner_n)
y_p, t_p = T._per_op(d, H, W, C, inner, inner_n, False, None)
torch.cuda.synchronize()
for r in range(reps):
    y_f, t_f = T._fused(d, H, W, C, inner, inner_n, False, None, y_out=T._nhwc(d["x"]).clone() if inplace else None,
                        x_is_y=inplace)
    for name, a, b in (("y", y_f, y_p), ("t1n", t_f, t_p)):
        if a is None:
            continue
        bad = (a != b).reshape(H * W, -1)
        nbad = int(bad.sum())
        rows = bad.any(dim=1).nonzero().flatten()
        msg = f"rep {r} {name}: {nbad} mismatching elements"
        if nbad:
            tiles = sorted(set((rows // 128).tolist()))
            cols = bad.any(dim=0).nonzero().flatten()
            msg += f", {len(rows)} rows in {len(tiles)} 128-row tiles (first {tiles[:12]}), channels {int(cols.min())}..{int(cols.max())} ({len(cols)} distinct)"
            msg += f", max abs diff {float((a.float() - b.float()).abs().max()):.4f}"
        print(msg)
