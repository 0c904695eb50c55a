"""Diagnostic: per-chunk encoder/decoder state comparison of the HT-S codec (not a test)."""
import sys
import numpy as np
import torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from util_frames import synth_frame
from dcvc_b200.model import DMC, DMCI

h, w, n_chunks = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
i_net = DMCI.synthetic(0); i_net.update(0.15); i_net = i_net.half().to("cuda")
p_net = DMC.synthetic(1); p_net.update(0.15); p_net = p_net.half().to("cuda")
frames = [synth_frame(h, w, 300)] + [synth_frame(h, w, 301 + c, channels=24) for c in range(n_chunks)]
pad_r, pad_b = i_net.get_padding_size(h, w, 16)
sps = {"height": h, "width": w}
TAPS = ["y", "y_hat", "z_i8", "cat_fam", "cat_enc", "common", "feature_i"]


def fetch():
    torch.cuda.synchronize()
    out = {}
    for t in TAPS:
        try:
            out[t] = p_net.proxy.debug_fetch(t, np.float16).copy()
        except Exception as e:  # noqa
            out[t] = None
    return out


def encode():
    x0 = frames[0].half().cuda().contiguous(memory_format=torch.channels_last)
    enc = i_net.compress(x0, 30, pad_b, pad_r)
    p_net.clear_dpb(); p_net.add_ref_feature_from_frame(enc["x_hat"])
    st = [fetch()]
    streams = [(enc["bit_stream"], enc["ec_parallel"])]
    for c in range(n_chunks):
        x = frames[1 + c].half().cuda().contiguous(memory_format=torch.channels_last)
        e = p_net.compress(x, 25, 1 if c == 1 else 0, pad_b, pad_r)
        streams.append((e["bit_stream"], e["ec_parallel"]))
        st.append(fetch())
    return streams, st, enc["x_hat"].clone()


def decode(streams):
    d = i_net.decompress(streams[0][0], sps, 30, streams[0][1])
    p_net.clear_dpb(); p_net.add_ref_feature_from_frame(d["x_hat"], False)
    st = [fetch()]
    for c in range(n_chunks):
        p_net.decompress(streams[1 + c][0], sps, 25, streams[1 + c][1], 1 if c == 1 else 0)
        st.append(fetch())
    return st, d["x_hat"].clone()


def cmp(tag, a, b):
    for i, (sa, sb) in enumerate(zip(a, b)):
        for t in TAPS:
            if sa[t] is None or sb[t] is None:
                continue
            ua, ub = sa[t].view(np.uint16).ravel(), sb[t].view(np.uint16).ravel()
            n = min(ua.size, ub.size)
            bad = np.nonzero(ua[:n] != ub[:n])[0]
            if bad.size and t in ("cat_fam", "cat_enc"):
                pitch = 1024 if t == "cat_fam" else 2048
                ch = bad % pitch
                hist = np.bincount(ch // 512, minlength=pitch // 512)
                print(f"   {t} mismatches per 512-channel slice: {hist.tolist()}")
            if bad.size:
                print(f"{tag} state {i} tap {t}: {bad.size}/{n} differ, first idx {bad[:6]}, last {bad[-3:]}, "
                      f"vals {sa[t].ravel()[bad[:4]]} vs {sb[t].ravel()[bad[:4]]}")


s1, e1, xi1 = encode()
s2, e2, xi2 = encode()
print("streams equal across two encodes:", [bytes(a[0]) == bytes(b[0]) for a, b in zip(s1, s2)])
cmp("enc-vs-enc", e1, e2)
d1, xd1 = decode(s1)
d2, xd2 = decode(s1)
cmp("dec-vs-dec", d1, d2)
print("intra x_hat enc==dec:", torch.equal(xi1, xd1))
cmp("enc-vs-dec", e1, d1)
print("done")
