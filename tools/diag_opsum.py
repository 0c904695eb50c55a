"""Find the first op whose arena checksum differs between two identical HT-S chunk encodes (diagnostic)."""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from util_frames import synth_frame
from dcvc_b200.model import DMC, DMCI

h, w = int(sys.argv[1]), int(sys.argv[2])
i_net = DMCI.synthetic(0); i_net.update(0.15); i_net = i_net.half().to("cuda")
p_net = DMC.synthetic(1); p_net.update(0.15); p_net = p_net.half().to("cuda")
f0 = synth_frame(h, w, 300).half().cuda().contiguous(memory_format=torch.channels_last)
f1 = synth_frame(h, w, 301, channels=24).half().cuda().contiguous(memory_format=torch.channels_last)
pad_r, pad_b = i_net.get_padding_size(h, w, 16)
enc = i_net.compress(f0, 30, pad_b, pad_r)
xh = enc["x_hat"].clone()
logs = []
for r in range(3):
    path = f"/tmp/opsum_{r}.txt"
    if os.path.exists(path):
        os.remove(path)
    os.environ["DCVC_B200_OPSUM"] = path
    p_net.clear_dpb(); p_net.add_ref_feature_from_frame(xh)
    e = p_net.compress(f1, 25, 0, pad_b, pad_r)
    torch.cuda.synchronize()
    del os.environ["DCVC_B200_OPSUM"]
    logs.append(open(path).read().splitlines())
print("ops logged:", [len(l) for l in logs])
for r in (1, 2):
    a, b = logs[0], logs[r]
    first = next((i for i, (x, y) in enumerate(zip(a, b)) if x != y), None)
    print(f"run0 vs run{r}: first differing op line = {first}")
    if first is not None:
        for i in range(max(0, first - 3), min(len(a), first + 4)):
            print("   ", i, a[i], "|", b[i].split()[-1])
