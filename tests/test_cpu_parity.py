"""CPU-side checks (no GPU): the oracle against the golden vectors minted from the reference,
the host-side mirrors (spec, CDF tables, rANS coder) against the same goldens, and the C-ABI
library loading with every symbol include/dcvc_b200.h declares."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def lib():
    from dcvc_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from dcvc_b200.build import build
        build()
    return _lib.load()


def test_abi_exports_every_declared_symbol(lib):
    from dcvc_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "dcvc_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(dcvc_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) > 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in dcvc_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.dcvc_abi_version() == 1
    assert b"sm_100a" in lib.dcvc_build_info()


def test_spec_matches_reference_layout():
    from dcvc_b200.spec import dmci_spec
    gold = json.load(open(os.path.join(GOLD, "dmci_state_dict_layout.json")))
    mine = {k: list(v) for k, v in dmci_spec().items()}
    assert mine == gold


def test_gaussian_cdf_tables(lib):
    from dcvc_b200.entropy import gaussian_cdf_tables
    g = np.load(os.path.join(GOLD, "gaussian_cdf.npz"))
    cdf, length = gaussian_cdf_tables()
    assert np.array_equal(cdf, g["quantized_cdf"])
    assert np.array_equal(length, g["cdf_length"])


def test_bit_estimator_cdf_tables(lib):
    from dcvc_b200.entropy import bit_estimator_cdf_tables
    g = np.load(os.path.join(GOLD, "bitest_cdf.npz"))
    cdf, length = bit_estimator_cdf_tables(torch.from_numpy(g["h"]), torch.from_numpy(g["b"]), torch.from_numpy(g["a"]))
    assert np.array_equal(cdf, g["quantized_cdf"])
    assert np.array_equal(length, g["cdf_length"])


def test_scale_lut_product_equals_oracle(lib):
    from dcvc_b200 import ops
    from oracle import ops_ref
    lut = ops.scale_index_lut()
    assert np.array_equal(lut, ops_ref.scale_index_lut())
    # monotone in the scale, spans the table
    vals = np.arange(0x2000, 0x5000, dtype=np.uint16)
    assert np.all(np.diff(lut[vals].astype(int)) >= 0) and lut[vals].max() == 127 and lut[vals].min() == 0


class _Rans:
    def __init__(self, lib, zc, zl, yc, yl):
        self.lib = lib
        self.h = C.c_void_p()
        assert lib.dcvc_rans_create(C.byref(self.h)) == 0
        for idx, (c, l) in enumerate(((zc, zl), (yc, yl))):
            c = np.ascontiguousarray(c, dtype=np.int32)
            l = np.ascontiguousarray(l, dtype=np.int32)
            assert lib.dcvc_rans_set_cdf(self.h, c.ctypes.data, l.ctypes.data, c.shape[0], c.shape[1], idx) == 0

    def encode(self, ys, z, qp, n_par):
        lib = self.lib
        lib.dcvc_rans_enc_reset(self.h)
        keep = []
        for k in (3, 2, 1, 0):
            a = np.ascontiguousarray(ys[k], dtype=np.int16)
            keep.append(a)
            lib.dcvc_rans_enc_y(self.h, a.ctypes.data, a.size)
        zz = np.ascontiguousarray(z, dtype=np.int8)
        lib.dcvc_rans_enc_z(self.h, zz.ctypes.data, zz.size, qp * 128, 128)
        data, size = C.c_void_p(), C.c_int32()
        assert lib.dcvc_rans_enc_finish(self.h, n_par, C.byref(data), C.byref(size)) == 0
        return np.ctypeslib.as_array(C.cast(data, C.POINTER(C.c_uint8)), (size.value,)).copy()

    def decode(self, stream, ys, n_z, qp, n_par):
        lib = self.lib
        s = np.ascontiguousarray(stream, dtype=np.uint8)
        assert lib.dcvc_rans_dec_set_stream(self.h, s.ctypes.data, s.size, n_par) == 0
        z = np.zeros(n_z, dtype=np.int8)
        assert lib.dcvc_rans_dec_z(self.h, z.ctypes.data, n_z, qp * 128, 128) == 0
        outs = []
        for k in range(4):
            rows = np.ascontiguousarray((ys[k] & 0xff).astype(np.uint8))
            o = np.zeros(rows.size, dtype=np.int8)
            assert lib.dcvc_rans_dec_y(self.h, o.ctypes.data, rows.ctypes.data, rows.size) == 0
            outs.append(o)
        return z, outs

    def __del__(self):
        self.lib.dcvc_rans_destroy(self.h)


@pytest.fixture(scope="module")
def tables(lib):
    from dcvc_b200.entropy import bit_estimator_cdf_tables, gaussian_cdf_tables
    from dcvc_b200.spec import dmci_spec, synth_state_dict
    sd = synth_state_dict({k: v for k, v in dmci_spec().items() if k.startswith("bit_estimator_z.")}, 0)
    # NB: the generator consumes random numbers in spec order, bit_estimator_z.* come first
    zc, zl = bit_estimator_cdf_tables(sd["bit_estimator_z.h"], sd["bit_estimator_z.b"], sd["bit_estimator_z.a"])
    yc, yl = gaussian_cdf_tables()
    return zc, zl, yc, yl


def test_rans_streams_bit_identical_to_reference(lib, tables):
    """product coder vs byte streams produced by the reference's own coder (golden)"""
    g = np.load(os.path.join(GOLD, "rans_streams.npz"))
    r = _Rans(lib, *tables)
    keys = sorted({k.rsplit("_", 1)[0] for k in g.files})
    assert len(keys) >= 8
    for key in keys:
        n_par = int(key.split("_")[0][1:])
        ys = [g[f"{key}_y{k}"] for k in range(4)]
        z, qp, ref_stream = g[f"{key}_z"], int(g[f"{key}_qp"]), g[f"{key}_stream"]
        mine = r.encode(ys, z, qp, n_par)
        assert mine.size == ref_stream.size and np.array_equal(mine, ref_stream), f"stream differs for {key}"
        dz, dys = r.decode(ref_stream, ys, z.size, qp, n_par)
        assert np.array_equal(dz, z)
        for k in range(4):
            assert np.array_equal(dys[k], (ys[k] >> 8).astype(np.int8)), f"decode differs {key} step {k}"


def test_rans_live_against_reference_build(lib, tables):
    """same check against the reference coder built into oracle/_ref, on fresh random symbols"""
    from oracle.build_ref import import_ref_shim
    ref = import_ref_shim()
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference)")
    zc, zl, yc, yl = tables
    enc, dec = ref.RansEncoder(), ref.RansDecoder()
    for c in (enc, dec):
        c.set_cdf(zc, zl, 0)
        c.set_cdf(yc, yl, 1)
    r = _Rans(lib, *tables)
    rng = np.random.default_rng(3)
    for n_par in (1, 2, 3, 4, 6, 7, 8):
        n = int(rng.integers(1000, 60000))
        sym = np.clip(np.rint(rng.standard_normal(n) * rng.choice([0.5, 3, 30], n)), -128, 127).astype(np.int32)
        ys = [((sym[: n // (k + 1)] << 8) + rng.integers(0, 128, n // (k + 1))).astype(np.int16) for k in range(4)]
        z = np.clip(np.rint(rng.standard_normal(128 * 9) * 5), -64, 63).astype(np.int8)
        qp = int(rng.integers(0, 64))
        enc.reset()
        enc.set_entropy_coder_parallel(n_par)
        for k in (3, 2, 1, 0):
            enc.encode_y(ys[k])
        enc.encode_z(z, qp * 128, 128)
        enc.flush()
        ref_stream = np.asarray(enc.get_encoded_stream()).copy()
        assert np.array_equal(r.encode(ys, z, qp, n_par), ref_stream)
        # reference decoder on the product's stream
        dec.set_entropy_coder_parallel(n_par)
        dec.set_stream(ref_stream)
        dec.decode_z(z.size, qp * 128, 128)
        assert np.array_equal(dec.get_decoded(z.size), z)
        dz, dys = r.decode(ref_stream, ys, z.size, qp, n_par)
        assert np.array_equal(dz, z)
        for k in range(4):   # includes escape-coded magnitudes (sigma 30 on narrow rows)
            assert np.array_equal(dys[k], (ys[k] >> 8).astype(np.int8)), (n_par, k)


def test_rans_edge_cases_against_reference_build(lib, tables):
    """the corners of the symbol format, byte for byte against the reference coder: empty steps (an all-skipped
    picture still carries z), a single symbol, fewer symbols than streams, the clamp limits -128 / 127 on the
    narrowest and the widest table rows (escape-coded magnitudes), z at its limits -64 / 63 on every qp row"""
    from oracle.build_ref import import_ref_shim
    ref = import_ref_shim()
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference)")
    zc, zl, yc, yl = tables
    enc, dec = ref.RansEncoder(), ref.RansDecoder()
    for c in (enc, dec):
        c.set_cdf(zc, zl, 0)
        c.set_cdf(yc, yl, 1)
    r = _Rans(lib, *tables)
    rng = np.random.default_rng(11)

    def pack(sym, rows):
        return ((np.asarray(sym, dtype=np.int32) << 8) + np.asarray(rows, dtype=np.int32)).astype(np.int16)

    e = np.zeros(0, dtype=np.int16)
    limits = np.array([-128, 127] * 64, dtype=np.int32)
    cases = {
        "all steps empty": [e, e, e, e],
        "only step 2": [e, e, pack(rng.integers(-3, 4, 777), rng.integers(0, 128, 777)), e],
        "single symbol": [pack([5], [17]), e, e, e],
        "fewer symbols than streams": [pack([1, -1, 0], [0, 64, 127]), pack([2], [3]), e, pack([-7, 7], [100, 5])],
        "clamp limits on every row": [pack(limits, np.arange(128)), pack(limits[::-1], np.arange(128)),
                                      pack(np.full(128, -128), np.zeros(128, dtype=np.int32)),
                                      pack(np.full(128, 127), np.full(128, 127))],
    }
    z_cases = [np.zeros(128, dtype=np.int8), np.tile(np.array([-64, 63], dtype=np.int8), 64 * 3),
               rng.integers(-64, 64, 128 * 5).astype(np.int8)]
    n_checked = 0
    for name, ys in cases.items():
        for z in z_cases:
            for n_par in (1, 2, 5, 8):
                for qp in (0, 63):
                    enc.reset()
                    enc.set_entropy_coder_parallel(n_par)
                    for k in (3, 2, 1, 0):
                        enc.encode_y(ys[k])
                    enc.encode_z(z, qp * 128, 128)
                    enc.flush()
                    ref_stream = np.asarray(enc.get_encoded_stream()).copy()
                    mine = r.encode(ys, z, qp, n_par)
                    assert np.array_equal(mine, ref_stream), (name, n_par, qp)
                    dz, dys = r.decode(ref_stream, ys, z.size, qp, n_par)
                    assert np.array_equal(dz, z), (name, n_par, qp)
                    for k in range(4):
                        assert np.array_equal(dys[k], (ys[k] >> 8).astype(np.int8)), (name, n_par, qp, k)
                    n_checked += 1
    assert n_checked == len(cases) * len(z_cases) * 4 * 2


def test_rans_decoder_survives_truncated_and_garbage_streams(lib, tables):
    """a damaged stream must decode to *something* without reading outside the stream buffer: the decoder runs
    unchecked blocks only while a zero-padded margin remains and falls back to a bounds-checked reader behind it
    (csrc/rans_host.cpp decode_run); reads behind the end return zero, as in the reference (rans.cpp byte reader)"""
    r = _Rans(lib, *tables)
    rng = np.random.default_rng(5)
    n = 40000
    sym = np.clip(np.rint(rng.standard_normal(n) * 2), -128, 127).astype(np.int32)
    ys = [((sym << 8) + rng.integers(0, 128, n)).astype(np.int16) for _ in range(4)]
    z = np.zeros(128 * 4, dtype=np.int8)
    for n_par in (1, 2, 5, 8):
        good = r.encode(ys, z, 10, n_par)
        header = 0 if n_par <= 2 else 4 * (n_par // 2 - 1 + n_par % 2)
        def damaged(stream):
            st = np.ascontiguousarray(stream, dtype=np.uint8)
            if lib.dcvc_rans_dec_set_stream(r.h, st.ctypes.data, st.size, n_par) != 0:
                return                                                   # rejected loudly (group offsets out of range)
            zz = np.zeros(z.size, dtype=np.int8)
            assert lib.dcvc_rans_dec_z(r.h, zz.ctypes.data, z.size, 10 * 128, 128) == 0
            for k in range(4):
                rows = np.ascontiguousarray((ys[k] & 0xff).astype(np.uint8))
                o = np.zeros(n, dtype=np.int8)
                assert lib.dcvc_rans_dec_y(r.h, o.ctypes.data, rows.ctypes.data, n) == 0

        for cut in (good.size // 2, good.size - 3, header + 16, header + 4):
            damaged(good[:cut])
        junk = rng.integers(0, 256, good.size, dtype=np.uint8)
        junk[:header] = good[:header]                                    # keep the group offsets valid
        damaged(junk)
        # and the intact stream still round-trips afterwards on the same handle
        dz, dys = r.decode(good, ys, z.size, 10, n_par)
        assert all(np.array_equal(dys[k], (ys[k] >> 8).astype(np.int8)) for k in range(4))


def test_pmf_to_quantized_cdf_against_reference(lib):
    from oracle.build_ref import import_ref_shim
    from dcvc_b200.entropy import pmf_to_quantized_cdf
    ref = import_ref_shim()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(9)
    for _ in range(200):
        n = int(rng.integers(2, 19))
        p = rng.random(n).astype(np.float32) ** 6
        p[rng.integers(0, n)] += 1e-9
        p = (p / p.sum()).astype(np.float32)
        assert list(ref.pmf_to_quantized_cdf(p.tolist())) == pmf_to_quantized_cdf(p).tolist()


@pytest.mark.parametrize("name", ["dmci_forward_64x64_qp0", "dmci_forward_64x64_qp32", "dmci_forward_128x64_qp63"])
def test_oracle_forward_pinned_to_reference(name):
    """oracle.forward_one_frame (fp32) vs outputs of the reference's own nn.Modules"""
    from dcvc_b200.spec import dmci_spec, synth_state_dict
    from oracle.dmci_oracle import DmciOracle
    g = np.load(os.path.join(GOLD, name + ".npz"))
    qp = int(name.split("qp")[1])
    o = DmciOracle(synth_state_dict(dmci_spec(), 0), emulate_fp16=False, threads=8)
    x = torch.from_numpy(g["x"])
    y = o.encoder(torch.nn.functional.pixel_unshuffle(x, 8), qp)
    assert np.allclose(y.numpy(), g["y"], atol=2e-4, rtol=1e-4)
    assert np.allclose(o.hyper_enc(y).numpy(), g["z"], atol=2e-4, rtol=1e-4)
    res = o.forward_one_frame(x, qp)
    # quantisation can flip on fp32 reassociation noise at exact ties; bound the damage
    diff = np.abs(res["x_hat"].numpy() - g["x_hat"])
    assert np.mean(diff > 1e-3) < 2e-3, f"x_hat mismatch fraction {np.mean(diff > 1e-3)}"


def test_oracle_compress_decompress_consistency():
    """proxy restatement: decoder reproduces the encoder's reconstruction bit for bit, through the
    reference's own rANS coder"""
    from oracle.build_ref import import_ref_shim
    if import_ref_shim() is None:
        pytest.skip("oracle/_ref not built")
    from dcvc_b200.spec import dmci_spec, synth_state_dict
    from oracle.dmci_oracle import DmciOracle
    g = np.load(os.path.join(GOLD, "dmci_forward_64x64_qp32.npz"))
    x = torch.from_numpy(g["x"])[:, :, :56, :60].contiguous()   # ragged: exercises padding
    o = DmciOracle(synth_state_dict(dmci_spec(), 0), skip_thres=0.15, emulate_fp16=True, threads=8)
    enc = o.compress(x, 40, 8, 4)
    dec = o.decompress(enc["bit_stream"], 40, 56, 60, enc["ec_parallel"])
    assert np.array_equal(enc["y_hat"].view(np.uint16), dec["y_hat"].view(np.uint16))
    assert torch.equal(enc["x_hat"], dec["x_hat"])
    assert enc["x_hat"].shape == (1, 3, 64, 64)
    total = sum(len(s) for s in enc["symbols"])
    assert 0 < total < 4 * 4 * 256 * 4 and len(enc["bit_stream"]) > 8


def test_hts_spec_matches_reference_layout():
    from dcvc_b200.spec import hts_spec
    gold = json.load(open(os.path.join(GOLD, "hts_state_dict_layout.json")))
    assert {k: list(v) for k, v in hts_spec().items()} == gold


def test_hts_oracle_forward_pinned_to_reference():
    """3 chunks with carried state (reset on the 2nd) vs the reference's DMC(HTS).forward_one_frame"""
    from dcvc_b200.spec import hts_spec, synth_state_dict
    from oracle.hts_oracle import HtsOracle
    g = np.load(os.path.join(GOLD, "hts_forward_64x64.npz"))
    o = HtsOracle(synth_state_dict(hts_spec(), 1), emulate_fp16=False, threads=8)
    o.feature_p = torch.nn.functional.pixel_unshuffle(torch.from_numpy(g["ref_frame"]), 8)
    for c, reset in enumerate([False, True, False]):
        res = o.forward_one_frame(torch.from_numpy(g[f"x{c}"]), 20 + c, reset_feature_memory=reset)
        x_hat = torch.cat(res["x_hat"], 1).numpy()
        d = np.abs(x_hat - g[f"x_hat{c}"])
        assert np.mean(d > 1e-3) < 5e-3, (c, np.mean(d > 1e-3), d.max())
        df = np.abs(o.feature_p.numpy() - g[f"ref_feature{c}"])
        assert np.mean(df > 2e-3) < 5e-3, (c, np.mean(df > 2e-3), df.max())


def test_hts_oracle_compress_decompress_consistency():
    from oracle.build_ref import import_ref_shim
    if import_ref_shim() is None:
        pytest.skip("oracle/_ref not built")
    from dcvc_b200.spec import hts_spec, synth_state_dict
    from oracle.hts_oracle import HtsOracle
    g = np.load(os.path.join(GOLD, "hts_forward_64x64.npz"))
    sd = synth_state_dict(hts_spec(), 1)
    enc, dec = HtsOracle(sd, 0.15, True, threads=8), HtsOracle(sd, 0.15, True, threads=8)
    ref = torch.from_numpy(g["ref_frame"])
    enc.add_ref_feature_from_frame(ref, True)        # encoder side of test_video.py:228-229
    dec.add_ref_feature_from_frame(ref, False)       # decoder side of test_video.py:314-315
    for c, reset in enumerate([False, True, False]):
        x = torch.from_numpy(g[f"x{c}"])[:, :, :56, :60].contiguous()
        e = enc.compress(x, 30 + c, reset, 8, 4)
        d = dec.decompress(e["bit_stream"], 30 + c, 56, 60, e["ec_parallel"], reset)
        assert np.array_equal(e["y_hat"].view(np.uint16), d["y_hat"].view(np.uint16)), f"chunk {c}"
        assert torch.equal(enc.feature_p, dec.feature_p), f"decoder state drifted at chunk {c}"
        assert len(d["x_hat"]) == 8 and d["x_hat"][0].shape == (1, 3, 64, 64)


def test_frame_io_oracle_matches_reference_expressions():
    """oracle/ops_ref.py frame IO restatements vs the reference driver's own expressions
    (test_video.py:74-76,115-122,352-361; transforms.py:69-90) evaluated with torch CPU half arithmetic"""
    import scipy.ndimage
    import torch.nn.functional as F
    from oracle import ops_ref
    rng = np.random.default_rng(5)
    y = rng.integers(0, 256, (34, 50), dtype=np.uint8)
    u = rng.integers(0, 256, (17, 25), dtype=np.uint8)
    v = rng.integers(0, 256, (17, 25), dtype=np.uint8)
    uv = scipy.ndimage.zoom(np.stack([u, v]).astype(np.float32), (1, 2, 2), order=0)       # ycbcr420_to_444_np
    x = torch.from_numpy(np.concatenate([y[None].astype(np.float32), uv], axis=0)).unsqueeze(0)
    x = x.half()
    x = x / 255.0
    x = x - 0.5
    assert torch.equal(ops_ref.yuv420_to_frame(y, u, v), x)
    g = torch.Generator().manual_seed(2)
    x_hat = (torch.rand(1, 3, 48, 64, generator=g) * 1.1 - 0.55).half()
    xs = x_hat[:, :, :34, :50] + 0.5
    y_rec, uv_rec = xs[:, :1], F.avg_pool2d(xs[:, 1:], kernel_size=2, stride=2)           # yuv_444_to_420
    y_rec = torch.clamp(y_rec * 255, 0, 255).round().byte().squeeze(0).numpy()
    uv_rec = torch.clamp(uv_rec * 255, 0, 255).byte().squeeze(0).numpy()
    ry, ru, rv = ops_ref.frame_to_yuv420(x_hat, 34, 50)
    assert np.array_equal(ry, y_rec[0]) and np.array_equal(ru, uv_rec[0]) and np.array_equal(rv, uv_rec[1])


def test_ld_spec_matches_reference_layout():
    """dcvc_b200.spec.ld_spec vs the state_dict of the reference's low-delay DMC (fixture minted by importing it)"""
    import json
    from dcvc_b200.spec import ld_spec
    ref = json.load(open(os.path.join(GOLD, "ld_state_dict_layout.json")))
    mine = {k: list(v) for k, v in ld_spec().items()}
    assert mine == ref


def test_ld_oracle_forward_pinned_to_reference():
    """oracle/ld_oracle.py (forward_one_frame + forward_prior_2x + state handling) vs the reference's own modules on
    a 4-frame sequence with a feature-memory reset (tests/golden/make_golden.py, seed-2 synthetic checkpoint)"""
    from dcvc_b200.spec import ld_spec, synth_state_dict
    from oracle.ld_oracle import LdOracle
    g = np.load(os.path.join(GOLD, "ld_forward_64x64.npz"))
    o = LdOracle(synth_state_dict(ld_spec(), 2), emulate_fp16=False)
    o.clear_dpb()
    o.feature_p = torch.nn.functional.pixel_unshuffle(torch.from_numpy(g["ref_frame"]), 8)
    for c, reset in enumerate([False, True, False, False]):
        r = o.forward_one_frame(torch.from_numpy(g[f"x{c}"]), int(g[f"qp{c}"]), reset_feature_memory=reset)
        assert (r["x_hat"] - torch.from_numpy(g[f"x_hat{c}"])).abs().max().item() < 2e-5
        assert (o.feature_p - torch.from_numpy(g[f"ref_feature{c}"])).abs().max().item() < 2e-5


def test_ld_oracle_compress_decompress_consistency():
    """proxy-control-flow restatement of the low-delay codec (dmc_ld_proxy.cpp:407-593): the decoder reproduces the
    encoder's latents and recurrent state over a 3-frame sequence with a feature-memory reset and ragged padding"""
    from oracle.build_ref import import_ref_shim
    if import_ref_shim() is None:
        pytest.skip("oracle/_ref not built")
    from dcvc_b200.spec import ld_spec, synth_state_dict
    from oracle.ld_oracle import LdOracle
    g = np.load(os.path.join(GOLD, "ld_forward_64x64.npz"))
    sd = synth_state_dict(ld_spec(), 2)
    enc, dec = LdOracle(sd, 0.15, True, threads=8), LdOracle(sd, 0.15, True, threads=8)
    ref = torch.from_numpy(g["ref_frame"])
    enc.add_ref_feature_from_frame(ref, True)        # encoder side of test_video.py:228-229
    dec.add_ref_feature_from_frame(ref, False)       # decoder side of test_video.py:314-315
    for c, reset in enumerate([False, True, False]):
        x = torch.from_numpy(g[f"x{c}"])[:, :, :56, :60].contiguous()
        e = enc.compress(x, 30 + c, reset, 8, 4)
        d = dec.decompress(e["bit_stream"], 30 + c, 56, 60, e["ec_parallel"], reset)
        assert np.array_equal(e["y_hat"].view(np.uint16), d["y_hat"].view(np.uint16)), f"frame {c}"
        assert torch.equal(enc.feature_p, dec.feature_p), f"decoder state drifted at frame {c}"
        assert d["x_hat"].shape == (1, 3, 64, 64)


def test_htl_spec_matches_reference_layout():
    """dcvc_b200.spec.htl_spec vs the state_dict of the reference's DMC(ModelStructure.HTL)"""
    import json
    from dcvc_b200.spec import htl_spec
    ref = json.load(open(os.path.join(GOLD, "htl_state_dict_layout.json")))
    assert {k: list(v) for k, v in htl_spec().items()} == ref


def test_htl_oracle_forward_pinned_to_reference():
    """oracle/htl_oracle.py (the non-HTS branches of video_model_ht.py + forward_prior_4x with scale updates) vs the
    reference's own modules on a 3-chunk sequence with a feature-memory reset (seed-3 synthetic checkpoint)"""
    from dcvc_b200.spec import htl_spec, synth_state_dict
    from oracle.htl_oracle import HtlOracle
    g = np.load(os.path.join(GOLD, "htl_forward_64x64.npz"))
    o = HtlOracle(synth_state_dict(htl_spec(), 3), emulate_fp16=False, threads=8)
    o.clear_dpb()
    o.feature_p = torch.nn.functional.pixel_unshuffle(torch.from_numpy(g["ref_frame"]), 8)
    for c, reset in enumerate([False, True, False]):
        r = o.forward_one_frame(torch.from_numpy(g[f"x{c}"]), int(g[f"qp{c}"]), reset_feature_memory=reset)
        assert (torch.cat(r["x_hat"], 1) - torch.from_numpy(g[f"x_hat{c}"])).abs().max().item() < 5e-5
        assert (o.feature_p - torch.from_numpy(g[f"ref_feature{c}"])).abs().max().item() < 5e-5


def test_htl_oracle_compress_decompress_consistency():
    """proxy-control-flow restatement of the HT-L chunk codec (dmc_htl_proxy.cpp:583-915): four symbol runs per chunk
    with scale updates; the decoder reproduces the encoder's latents and recurrent state (reset + ragged padding)"""
    from oracle.build_ref import import_ref_shim
    if import_ref_shim() is None:
        pytest.skip("oracle/_ref not built")
    from dcvc_b200.spec import htl_spec, synth_state_dict
    from oracle.htl_oracle import HtlOracle
    g = np.load(os.path.join(GOLD, "htl_forward_64x64.npz"))
    sd = synth_state_dict(htl_spec(), 3)
    enc, dec = HtlOracle(sd, 0.15, True, threads=8), HtlOracle(sd, 0.15, True, threads=8)
    ref = torch.from_numpy(g["ref_frame"])
    enc.add_ref_feature_from_frame(ref, True)
    dec.add_ref_feature_from_frame(ref, False)
    for c, reset in enumerate([False, True]):
        x = torch.from_numpy(g[f"x{c}"])[:, :, :56, :60].contiguous()
        e = enc.compress(x, 30 + c, reset, 8, 4)
        d = dec.decompress(e["bit_stream"], 30 + c, 56, 60, e["ec_parallel"], reset)
        assert np.array_equal(e["y_hat"].view(np.uint16), d["y_hat"].view(np.uint16)), f"chunk {c}"
        assert torch.equal(enc.feature_p, dec.feature_p), f"decoder state drifted at chunk {c}"
        assert len(d["x_hat"]) == 8 and d["x_hat"][0].shape == (1, 3, 64, 64)


@pytest.mark.parametrize("kind_name", ["TCONV2X2", "CONV3X3_PS2"])
def test_pack_weight_of_pixel_shuffle_kinds_as_a_plain_gemm(lib, kind_name):
    """host half of the pixel-shuffle GEMM kinds, no device: the packed operand [phase*Cout + co][tap][c], fed to a
    plain numpy GEMM over shifted (zero-padded) views and scattered phase-major, must equal conv + pixel_shuffle(2)
    (layers.py:92-103 SubpelConv2x); guards the layout contract between dcvc_pack_weight, the tap table of gemm_plan
    and the phase-major bias of CodecBase::load_conv."""
    import torch.nn.functional as F

    from dcvc_b200 import _lib
    kind = getattr(_lib, "GEMM_" + kind_name)
    k = 3 if kind_name == "CONV3X3_PS2" else 1
    cin, cout, H, W = 8, 6, 5, 7
    g = torch.Generator().manual_seed(11)
    w = (torch.randn(4 * cout, cin, k, k, generator=g) * 0.2).half()
    b = (torch.randn(4 * cout, generator=g) * 0.2).half()
    x = torch.randn(1, cin, H, W, generator=g).half()
    want = F.pixel_shuffle(F.conv2d(x.float(), w.float(), b.float(), padding=k // 2), 2)
    packed = torch.empty(w.numel(), dtype=torch.float16)
    assert lib.dcvc_pack_weight(kind, w.contiguous().data_ptr(), 4 * cout, cin, k, k, packed.data_ptr()) == 0
    B = packed.float().view(4 * cout, k * k, cin)                       # [n][tap][c]
    xp = F.pad(x.float(), (k // 2,) * 4)[0]                              # zero padding = TMA out-of-bounds fill
    acc = torch.zeros(H, W, 4 * cout)
    for t in range(k * k):
        dy, dx = (t // 3 - 1, t % 3 - 1) if k == 3 else (0, 0)          # gemm_plan's tap table
        a = xp[:, k // 2 + dy:k // 2 + dy + H, k // 2 + dx:k // 2 + dx + W].permute(1, 2, 0)   # [H][W][c]
        acc += a @ B[:, t, :].T
    bias_pm = b.float().view(cout, 4).T.reshape(-1)                      # load_conv: column ph*Cout+co <- channel co*4+ph
    acc += bias_pm
    out = torch.zeros(1, cout, 2 * H, 2 * W)
    for ph in range(4):                                                  # epilogue: phase ph -> pixel (2y + ph//2, 2x + ph%2)
        out[0, :, ph // 2::2, ph % 2::2] = acc[:, :, ph * cout:(ph + 1) * cout].permute(2, 0, 1)
    assert torch.allclose(out, want, atol=1e-5)
