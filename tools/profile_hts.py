"""One intra frame + three HT-S chunk encodes + decodes at HxW through the public API — the command profiled under ncu for
the 4K DRAM-traffic capture (BASELINE.json configs[4])."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util_frames import synth_frame  # noqa: E402
from dcvc_b200.model import DMC, DMCI  # noqa: E402

h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2160, 3840)
i_net = DMCI.synthetic(0); i_net.update(0.15); i_net = i_net.half().to("cuda")
p_net = DMC.synthetic(1); p_net.update(0.15); p_net = p_net.half().to("cuda")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    x0 = synth_frame(h, w, 4000).half().cuda().contiguous(memory_format=torch.channels_last)
    pr, pb = i_net.get_padding_size(h, w, 16)
    e = i_net.compress(x0, 32, pb, pr)
    p_net.add_ref_feature_from_frame(e["x_hat"])
    chunks = [synth_frame(h, w, 4100 + c, channels=24).half().cuda().contiguous(memory_format=torch.channels_last) for c in range(2)]
    encs = [p_net.compress(c, 32, 0, pb, pr) for c in chunks]
    d = i_net.decompress(e["bit_stream"], {"height": h, "width": w}, 32, e["ec_parallel"])
    p_net.add_ref_feature_from_frame(d["x_hat"], False)
    p_net.decompress(encs[0]["bit_stream"], {"height": h, "width": w}, 32, encs[0]["ec_parallel"], 0)   # warm: plans + graphs
    torch.cuda.synchronize()
    torch.cuda.profiler.start()      # ncu --profile-from-start off: only the second chunk decode is captured
    p_net.decompress(encs[1]["bit_stream"], {"height": h, "width": w}, 32, encs[1]["ec_parallel"], 0)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
torch.cuda.synchronize()
print("done")
