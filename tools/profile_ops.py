"""In-situ (warm L2, no ncu) per-launch timing of one 1080p decode: CUDA events around every launch,
dumped as CSV (kind, op index in segment, us, algorithmic bytes, flops)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ops.csv"
from util_frames import synth_frame  # noqa: E402
from dcvc_b200.model import DMCI  # noqa: E402

h, w = 1080, 1920
m = DMCI.synthetic(0)
m.update(0.15)
m = m.half().to("cuda")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    x = synth_frame(h, w, 1234).half().cuda().contiguous(memory_format=torch.channels_last)
    pr, pb = m.get_padding_size(h, w, 16)
    enc = m.compress(x, 32, pb, pr)
    for _ in range(3):
        m.decompress(enc["bit_stream"], {"height": h, "width": w}, 32, enc["ec_parallel"])
    torch.cuda.synchronize()
    m.proxy.profile_enable(True)
    m.decompress(enc["bit_stream"], {"height": h, "width": w}, 32, enc["ec_parallel"])  # warm (profile mode)
    if os.path.exists(out):
        os.remove(out)
    os.environ["DCVC_B200_PROFILE_CSV"] = out
    m.decompress(enc["bit_stream"], {"height": h, "width": w}, 32, enc["ec_parallel"])
    torch.cuda.synchronize()
print("wrote", out)
