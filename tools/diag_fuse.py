"""Codec-level diagnosis of the fused DepthConvBlock tail: Intra encode + decode at HxW under several switch settings;
every variant must reproduce the per-op build's stream and reconstructions bit for bit.
  python tools/diag_fuse.py H W [qp]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util_frames import synth_frame  # noqa: E402
from dcvc_b200.model import DMCI  # noqa: E402

h, w = int(sys.argv[1]), int(sys.argv[2])
qp = int(sys.argv[3]) if len(sys.argv) > 3 else 40
x = synth_frame(h, w, 77).half().cuda().contiguous(memory_format=torch.channels_last)
VARIANTS = [("per-op", {"DCVC_B200_FUSE_TAIL": "0"}),
            ("fused", {}),
            ("fused again", {}),
            ("fused, no graphs", {"DCVC_B200_GRAPHS": "0"}),
            ("fused, no PDL", {"DCVC_B200_PDL": "0"}),
            ("fused, 8 pairs", {"DCVC_B200_DT_MAXPAIRS": "8"}),
            ("fused, any width", {"DCVC_B200_FUSE_TAIL": "16384"})]
base = None
for name, env in VARIANTS:
    for k in ("DCVC_B200_FUSE_TAIL", "DCVC_B200_GRAPHS", "DCVC_B200_PDL", "DCVC_B200_DT_MAXPAIRS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    m = DMCI.synthetic(0)
    m.update(0.15)
    m = m.half().to("cuda")
    pad_r, pad_b = m.get_padding_size(h, w, 16)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        res = []
        for rep in range(2):
            enc = m.compress(x, qp, pad_b, pad_r)
            xe = enc["x_hat"].clone()
            dec = m.decompress(enc["bit_stream"], {"height": h, "width": w}, qp, enc["ec_parallel"])
            xd = dec["x_hat"].clone()
            torch.cuda.synchronize()
            res.append((enc["bit_stream"], xe, xd))
    if base is None:
        base = res[0]
    msg = []
    for rep, (bs, xe, xd) in enumerate(res):
        bad_e = int((xe != base[1]).sum())
        bad_d = int((xd != base[2]).sum())
        msg.append(f"rep{rep}: stream {'same' if bs == base[0] else f'DIFF ({len(bs)} vs {len(base[0])} B)'}, "
                   f"x_hat enc {bad_e} dec {bad_d} differing elements, enc==dec {bool(torch.equal(xe, xd))}")
    print(f"[{name}] " + " | ".join(msg), flush=True)
    del m
