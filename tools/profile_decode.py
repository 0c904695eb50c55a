"""One 1080p encode + N decodes through the public API — the command profiled under ncu
(see /opt/skills/guides/B200_PROFILING.md).  DCVC_B200_GRAPHS=0 makes every kernel a plain launch."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util_frames import synth_frame  # noqa: E402
from dcvc_b200.model import DMCI  # noqa: E402

h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1080, 1920)
n_dec = int(sys.argv[3]) if len(sys.argv) > 3 else 1
m = DMCI.synthetic(0)
m.update(0.15)
m = m.half().to("cuda")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    x = synth_frame(h, w, 1234).half().cuda().contiguous(memory_format=torch.channels_last)
    pr, pb = m.get_padding_size(h, w, 16)
    enc = m.compress(x, 32, pb, pr)
    for _ in range(n_dec):
        m.decompress(enc["bit_stream"], {"height": h, "width": w}, 32, enc["ec_parallel"])
torch.cuda.synchronize()
print("done", len(enc["bit_stream"]))
