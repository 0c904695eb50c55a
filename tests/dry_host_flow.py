"""Driver of tests/test_host_dry_run.py — runs INSIDE a subprocess whose CUDA runtime is tests/cpp/cuda_dry_shim.cpp.

    python dry_host_flow.py <libdcvc_dry.so> <libdryshim.so> <plan|check> <intra|hts|ld|htl> HxW [HxW ...]

plan  : kernels do nothing (zero-filled "device" buffers).  Walks set_param / finalize, the plan at the given picture
        sizes, compress and decompress through the C ABI: the host logic must run to completion.
check : DRY_SHIM_EMULATE=1 — every launch runs the host restatement of its kernel (tests/cpp/cuda_emu_kernels.cpp), so the
        library's own operand wiring / weight packing / tensor maps produce real numbers, which are compared with the
        CPU oracle of that codec: stream size, reconstruction PSNR, and encoder-vs-decoder state identity."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from dcvc_b200 import _lib  # noqa: E402
from dcvc_b200.spec import dmci_spec, htl_spec, hts_spec, ld_spec, synth_state_dict  # noqa: E402
from util_frames import psnr, synth_frame  # noqa: E402

SKIP = 0.15


def load(path):
    lib = C.CDLL(path)
    for name, (res, args) in _lib.SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


def cdf_tables(sd):
    from dcvc_b200.entropy import bit_estimator_cdf_tables, gaussian_cdf_tables
    zc, zl = bit_estimator_cdf_tables(sd["bit_estimator_z.h"], sd["bit_estimator_z.b"], sd["bit_estimator_z.a"])
    yc, yl = gaussian_cdf_tables()
    return {"bit_estimator_z.quantized_cdf": zc, "bit_estimator_z.cdf_length": zl,
            "gaussian_encoder.quantized_cdf": yc, "gaussian_encoder.cdf_length": yl}


class Handle:
    """one codec handle fed through the C ABI with host arrays standing in for device tensors"""

    def __init__(self, lib, kind, spec, seed):
        self.lib, self.h, self.keep = lib, C.c_void_p(), []
        if lib.dcvc_create(kind, 0, C.byref(self.h)):
            raise RuntimeError("dcvc_create: " + lib.dcvc_last_error().decode())
        sd = synth_state_dict(spec, seed)
        for name, t in sd.items():
            a = np.ascontiguousarray(t.numpy().astype(np.float16))
            self.keep.append(a)
            shape = (C.c_int64 * max(1, a.ndim))(*a.shape)
            # on_device = 1: with the shim a "device" pointer is host memory
            self.chk(lib.dcvc_set_param(self.h, name.encode(), C.c_void_p(a.ctypes.data), _lib.DTYPE_F16, a.ndim, shape, 1), name)
        for name, a in cdf_tables(sd).items():
            a = np.ascontiguousarray(np.asarray(a), dtype=np.int32)
            self.keep.append(a)
            shape = (C.c_int64 * a.ndim)(*a.shape)
            self.chk(lib.dcvc_set_param(self.h, name.encode(), C.c_void_p(a.ctypes.data), _lib.DTYPE_I32, a.ndim, shape, 0), name)
        self.chk(lib.dcvc_finalize_params(self.h, C.c_float(SKIP)), "finalize_params")

    def chk(self, rc, what):
        if rc:
            raise RuntimeError(f"{what}: " + self.lib.dcvc_codec_error(self.h).decode(errors="replace"))

    def fetch(self, name, dtype, nbytes):
        buf = np.zeros(nbytes, dtype=np.uint8)
        n = C.c_int64()
        self.chk(self.lib.dcvc_debug_fetch(self.h, name.encode(), buf.ctypes.data, nbytes, C.byref(n)), "debug_fetch " + name)
        return buf[: n.value].view(dtype).copy()

    def profiled(self, fn):
        """algorithmic bytes / FLOPs the codec books for the launches of fn() (Segment annotations: bench.py's roofline
        numerators); kinds 0 = pw_gemm (all instantiations), 1 = dw3x3, 2 = elementwise / entropy, 3 = fused dcb_tail"""
        self.lib.dcvc_profile_enable(self.h, 1)
        out = fn()
        acc = {"bytes": 0.0, "flops": 0.0, "launches": 0}
        for kind in range(4):
            ms, n, b, f = C.c_double(), C.c_int64(), C.c_double(), C.c_double()
            self.lib.dcvc_profile_get(self.h, kind, C.byref(ms), C.byref(n), C.byref(b), C.byref(f))
            acc["bytes"] += b.value; acc["flops"] += f.value; acc["launches"] += n.value
        self.lib.dcvc_profile_enable(self.h, 0)
        return out, acc

    def close(self):
        self.lib.dcvc_destroy(self.h)


def f16(x):
    """[1,C,H,W] float tensor -> contiguous fp16 numpy NCHW"""
    return np.ascontiguousarray(x.numpy().astype(np.float16))


def nhwc_to_tensor(a):
    return torch.from_numpy(a.astype(np.float32)).permute(2, 0, 1).unsqueeze(0).contiguous()


def pads(H, W):
    return (H + 15) // 16 * 16, (W + 15) // 16 * 16


# ---------------------------------------------------------------------------------------------------- intra
def intra_compress(hd, x, qp):
    _, _, H, W = x.shape
    Hp, Wp = pads(H, W)
    xh = np.zeros((Hp, Wp, 3), dtype=np.float16)
    bs, n, ec = C.c_void_p(), C.c_int32(), C.c_int32()
    hd.chk(hd.lib.dcvc_compress(hd.h, x.ctypes.data, H, W, H * W, W, 1, qp, Hp - H, Wp - W, None, C.byref(bs), C.byref(n),
                                C.byref(ec), xh.ctypes.data), "compress")
    return bytes(C.string_at(bs.value, n.value)), ec.value, xh


def intra_decompress(hd, stream, qp, H, W, ec):
    Hp, Wp = pads(H, W)
    xh = np.zeros((Hp, Wp, 3), dtype=np.float16)
    hd.chk(hd.lib.dcvc_decompress(hd.h, stream, len(stream), qp, H, W, ec, None, xh.ctypes.data), "decompress")
    return xh


def run_intra(lib, mode, sizes):
    hd = Handle(lib, _lib.KIND_INTRA, dmci_spec(), 0)
    out = []
    for (H, W) in sizes:
        Hp, Wp = pads(H, W)
        xt = synth_frame(H, W, 1234) if mode == "check" else torch.zeros(1, 3, H, W)
        stream, ec, xh_enc = intra_compress(hd, f16(xt), 32)
        xh_dec, acc = hd.profiled(lambda: intra_decompress(hd, stream, 32, H, W, ec))
        rec = {"size": [H, W], "bytes": len(stream), "ec": ec, "decode_alg_gb": acc["bytes"] / 1e9,
               "decode_gmac": acc["flops"] / 2e9, "decode_launches": acc["launches"]}
        if mode == "check":
            from oracle.dmci_oracle import DmciOracle
            assert np.array_equal(xh_enc.view(np.uint16), xh_dec.view(np.uint16)), "decoder reconstruction differs from the encoder's"
            o = DmciOracle(synth_state_dict(dmci_spec(), 0), skip_thres=SKIP, emulate_fp16=True, threads=int(os.environ.get("DCVC_DRY_THREADS", "8")))
            ref = o.compress(xt, 32, Hp - H, Wp - W)
            rec["ref_bytes"] = len(ref["bit_stream"])
            rec["psnr"] = psnr(nhwc_to_tensor(xh_enc)[:, :, :H, :W], xt)
            rec["ref_psnr"] = psnr(ref["x_hat"][:, :, :H, :W], xt)
            totals = hd.fetch("totals", np.int32, 16)
            rec["symbols"] = [int(v) for v in totals]
            rec["ref_symbols"] = [len(s) for s in ref["symbols"]]
        rec["arena_overflow_blocks"] = int(hd.fetch("arena_overflow_blocks", np.int32, 4)[0])
        out.append(rec)
    hd.close()
    return out


# ---------------------------------------------------------------------------------------------------- video kinds
def video_add_ref(hd, frame, apply):
    _, _, Hp, Wp = frame.shape
    hd.chk(hd.lib.dcvc_add_ref_feature_from_frame(hd.h, frame.ctypes.data, Hp, Wp, Hp * Wp, Wp, 1, 1 if apply else 0, None), "add_ref")


def video_compress(hd, x, qp, reset):
    _, _, H, W = x.shape
    Hp, Wp = pads(H, W)
    bs, n, ec = C.c_void_p(), C.c_int32(), C.c_int32()
    hd.chk(hd.lib.dcvc_compress_chunk(hd.h, x.ctypes.data, H, W, H * W, W, 1, qp, 1 if reset else 0, Hp - H, Wp - W, None,
                                      C.byref(bs), C.byref(n), C.byref(ec)), "compress_chunk")
    return bytes(C.string_at(bs.value, n.value)), ec.value


def video_decompress(hd, stream, qp, H, W, ec, reset, frames):
    Hp, Wp = pads(H, W)
    outs = [np.zeros((Hp, Wp, 3), dtype=np.float16) for _ in range(frames)]
    ptrs = (C.c_void_p * frames)(*[o.ctypes.data for o in outs])
    hd.chk(hd.lib.dcvc_decompress_chunk(hd.h, stream, len(stream), qp, H, W, ec, 1 if reset else 0, None, ptrs), "decompress_chunk")
    return outs


def run_video(lib, mode, which, sizes):
    kind, spec, seed, frames, oracle_cls, fam_c = {
        "hts": (_lib.KIND_HTS, hts_spec(), 1, 8, ("oracle.hts_oracle", "HtsOracle"), 1024),
        "ld": (_lib.KIND_LD, ld_spec(), 2, 1, ("oracle.ld_oracle", "LdOracle"), 512),
        "htl": (_lib.KIND_HTL, htl_spec(), 3, 8, ("oracle.htl_oracle", "HtlOracle"), 1024),
    }[which]
    hd = Handle(lib, kind, spec, seed)
    out = []
    resets = (False, True, False)
    for (H, W) in sizes:
        Hp, Wp = pads(H, W)
        if mode == "check":
            ref_frame = torch.nn.functional.pad(synth_frame(H, W, 700), (0, Wp - W, 0, Hp - H), mode="replicate")
            chunks = [synth_frame(H, W, 701 + c, channels=3 * frames) for c in range(len(resets))]
        else:
            ref_frame = torch.zeros(1, 3, Hp, Wp)
            chunks = [torch.zeros(1, 3 * frames, H, W) for _ in resets]
        video_add_ref(hd, f16(ref_frame), True)
        streams = [video_compress(hd, f16(x), 25, r) for x, r in zip(chunks, resets)]
        p8 = (Hp // 8) * (Wp // 8)
        enc_state = hd.fetch("cat_fam", np.uint16, p8 * fam_c * 2) if mode == "check" else None
        video_add_ref(hd, f16(ref_frame), False)
        recon = [video_decompress(hd, s, 25, H, W, ec, r, frames) for (s, ec), r in zip(streams[:-1], resets[:-1])]
        last, acc = hd.profiled(lambda: video_decompress(hd, streams[-1][0], 25, H, W, streams[-1][1], resets[-1], frames))
        recon.append(last)
        rec = {"size": [H, W], "bytes": [len(s[0]) for s in streams], "decode_alg_gb": acc["bytes"] / 1e9,
               "decode_gmac": acc["flops"] / 2e9, "decode_launches": acc["launches"]}
        if mode == "check":
            dec_state = hd.fetch("cat_fam", np.uint16, p8 * fam_c * 2)
            half = fam_c // 2
            e, d = enc_state.reshape(-1, fam_c)[:, half:], dec_state.reshape(-1, fam_c)[:, half:]
            assert np.array_equal(e, d), "decoder feature_p differs from the encoder's"
            mod = __import__(oracle_cls[0], fromlist=[oracle_cls[1]])
            oe = getattr(mod, oracle_cls[1])(synth_state_dict(spec, seed), SKIP, True, threads=int(os.environ.get("DCVC_DRY_THREADS", "8")))
            od = getattr(mod, oracle_cls[1])(synth_state_dict(spec, seed), SKIP, True, threads=int(os.environ.get("DCVC_DRY_THREADS", "8")))
            oe.add_ref_feature_from_frame(ref_frame, True)
            od.add_ref_feature_from_frame(ref_frame, False)
            rec["ref_bytes"], rec["psnr"], rec["ref_psnr"] = [], [], []
            for c, (x, r) in enumerate(zip(chunks, resets)):
                en = oe.compress(x, 25, r, Hp - H, Wp - W)
                de = od.decompress(en["bit_stream"], 25, H, W, en["ec_parallel"], r)
                rec["ref_bytes"].append(len(en["bit_stream"]))
                ref_hats = de["x_hat"] if frames > 1 else [de["x_hat"]]
                for f in range(frames):
                    src = x[:, 3 * f:3 * f + 3]
                    rec["psnr"].append(psnr(nhwc_to_tensor(recon[c][f])[:, :, :H, :W], src))
                    rec["ref_psnr"].append(psnr(ref_hats[f][:, :, :H, :W], src))
        rec["arena_overflow_blocks"] = int(hd.fetch("arena_overflow_blocks", np.int32, 4)[0])
        out.append(rec)
    hd.close()
    return out


def main():
    lib = load(sys.argv[1])
    shim = C.CDLL(sys.argv[2])
    for f in ("dry_shim_launches", "dry_shim_graph_launches", "dry_shim_tensor_maps", "dry_shim_bytes", "dry_shim_capture_forks"):
        getattr(shim, f).restype = C.c_longlong
    mode, which = sys.argv[3], sys.argv[4]
    if (mode == "check") != (os.environ.get("DRY_SHIM_EMULATE") == "1"):
        raise SystemExit("check mode needs DRY_SHIM_EMULATE=1 (and plan mode must not have it)")
    sizes = [tuple(int(v) for v in s.split("x")) for s in sys.argv[5:]]
    res = run_intra(lib, mode, sizes) if which == "intra" else run_video(lib, mode, which, sizes)
    print(json.dumps({"codec": which, "mode": mode, "runs": res, "launches": shim.dry_shim_launches(),
                      "graph_launches": shim.dry_shim_graph_launches(), "tensor_maps": shim.dry_shim_tensor_maps(),
                      "device_bytes": shim.dry_shim_bytes(), "capture_forks": shim.dry_shim_capture_forks()}))


if __name__ == "__main__":
    main()
