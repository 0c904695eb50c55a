"""Host half of the codecs without a GPU (CPU test tier).

The shipped library cannot run here (no device, no CPU path — by design).  For the test, the SAME object files are
relinked against the shared CUDA runtime into a temporary directory and run in a subprocess whose CUDA runtime entry
points are replaced by tests/cpp/cuda_dry_shim.cpp (LD_PRELOAD):

* plan mode  — launches are no-ops: parameter loading / weight repacking, arena sizing, every GEMM plan (geometry
  checks, tile plans, tensor maps checked against the driver's argument rules), segment building and the compress /
  decompress control flow of all four codec kinds must run to completion at 64x64, a ragged size and 1080p;
* check mode — every launch runs a host restatement of WHAT that kernel computes, fed from the kernel's own argument
  block (tests/cpp/cuda_emu_kernels.cpp), so the library's operand wiring, packing and tensor-map geometry produce real
  numbers; they are compared with the CPU oracle of each codec with the tolerances of the GPU parity tests
  (bytes within 2 %, PSNR within 0.1 dB) plus encoder/decoder state identity.

* the `-m gpu` test files themselves, unmodified, in the same emulated process (tests/dry_pytest_runner.py +
  tests/dry_torch_patch.py answer torch's `cuda` idioms with host memory): model mirrors -> inference_extensions_cuda
  -> proxies -> C ABI -> emulated kernels.  By default the op-level file and one sequence-driver case; with
  DCVC_B200_DRY_FULL=1 also the small cases of every codec file (~10 more minutes).

This is test infrastructure: nothing under dcvc_b200/ can reach the shim, and the kernels themselves (tcgen05 / TMA
code) are only exercised by the `-m gpu` tests.  HT-L (experimental, never run on a device) is covered here too."""
import json
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_LIB = "/usr/local/cuda/lib64"


@pytest.fixture(scope="session")
def dry(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    gxx = shutil.which("g++")
    if not (os.path.exists(nvcc) and gxx and os.path.exists(os.path.join(CUDA_LIB, "libcudart.so"))):
        pytest.skip("needs nvcc, g++ and the shared CUDA runtime")
    sys.path.insert(0, ROOT)
    from dcvc_b200 import build as b
    b.build()                                            # no-op when the objects are current
    objs = [os.path.join(b.OBJ_DIR, f) for f in sorted(os.listdir(b.OBJ_DIR)) if f.endswith(".o")]
    d = tmp_path_factory.mktemp("dry")
    lib, shim = str(d / "libdcvc_dry.so"), str(d / "libdryshim.so")
    subprocess.run([nvcc, "-shared", "-cudart", "shared", "-o", lib] + objs +
                   ["-lpthread", "-gencode", "arch=compute_100a,code=sm_100a"], check=True, capture_output=True)
    subprocess.run([gxx, "-O2", "-shared", "-fPIC", "-std=c++17", "-I/usr/local/cuda/include",
                    os.path.join(ROOT, "tests/cpp/cuda_dry_shim.cpp"), os.path.join(ROOT, "tests/cpp/cuda_emu_kernels.cpp"),
                    "-o", shim], check=True, capture_output=True)
    return lib, shim


def _run(dry, mode, codec, sizes, timeout=900):
    lib, shim = dry
    env = dict(os.environ)
    env.update({"LD_PRELOAD": shim, "LD_LIBRARY_PATH": CUDA_LIB + ":" + env.get("LD_LIBRARY_PATH", ""),
                "DCVC_B200_EXPERIMENTAL_HTL": "1", "DCVC_B200_RANS_SPIN_US": "0"})
    env.pop("DRY_SHIM_EMULATE", None)
    if mode == "check":
        env["DRY_SHIM_EMULATE"] = "1"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests/dry_host_flow.py"), lib, shim, mode, codec] + sizes,
                       env=env, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])


# decode-side algorithmic GB and GMAC per picture (Intra, LD) / per 8-frame chunk (HT-S, HT-L) at 1088x1920, measured on
# the reference's own modules with forward hooks: the numerators of bench.py's roofline
REFERENCE_DECODE_WORK = {"intra": (5.89, 701.2), "hts": (14.09, 1483.4), "ld": (2.80, 147.8), "htl": (17.33, 2001.8)}


@pytest.mark.parametrize("codec", ["intra", "hts", "ld", "htl"])
def test_host_flow_plans_and_runs_at_all_sizes(dry, codec):
    res = _run(dry, "plan", codec, ["64x64", "200x328", "1080x1920", "2160x3840", "1096x1928"])
    assert len(res["runs"]) == 5 and res["launches"] > 100 and res["tensor_maps"] > 100
    for run in res["runs"]:
        sizes = run["bytes"] if isinstance(run["bytes"], list) else [run["bytes"]]
        assert all(s > 4 for s in sizes)            # an all-skipped picture still carries z
        # the arena estimate of the plan covered every buffer (HT-S at 4K once did not: hyper-prior padding buffer)
        assert run["arena_overflow_blocks"] == 0, (codec, run["size"])
    # the launches of one 1080p decode book exactly the reference network's work: no op missing, none counted twice
    gb, gmac = REFERENCE_DECODE_WORK[codec]
    full = res["runs"][2]
    assert abs(full["decode_alg_gb"] - gb) <= 0.005 * gb + 0.005, (codec, full["decode_alg_gb"], gb)
    assert abs(full["decode_gmac"] - gmac) <= 0.002 * gmac, (codec, full["decode_gmac"], gmac)


@pytest.mark.parametrize("codec,sizes", [("intra", ["64x64", "72x104"]), ("hts", ["72x104"]), ("ld", ["72x104"]),
                                         ("htl", ["72x104"])])
def test_emulated_codec_matches_its_oracle(dry, codec, sizes):
    res = _run(dry, "check", codec, sizes, timeout=1800)
    for run in res["runs"]:
        as_list = lambda v: v if isinstance(v, list) else [v]  # noqa: E731
        for n, n_ref in zip(as_list(run["bytes"]), as_list(run["ref_bytes"])):
            assert abs(n - n_ref) <= 0.02 * n_ref + 8, (codec, run["size"], run["bytes"], run["ref_bytes"])
        for p, p_ref in zip(as_list(run["psnr"]), as_list(run["ref_psnr"])):
            assert abs(p - p_ref) <= 0.1, (codec, run["size"], p, p_ref)
        if "symbols" in run:
            for n, n_ref in zip(run["symbols"], run["ref_symbols"]):
                assert abs(n - n_ref) <= 0.01 * n_ref + 4


def _pytest_under_emulation(dry, args, timeout=3000):
    lib, shim = dry
    env = dict(os.environ)
    env.update({"LD_PRELOAD": shim, "LD_LIBRARY_PATH": CUDA_LIB + ":" + env.get("LD_LIBRARY_PATH", ""),
                "DCVC_B200_EXPERIMENTAL_HTL": "1", "DCVC_B200_RANS_SPIN_US": "0", "DRY_SHIM_EMULATE": "1"})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests/dry_pytest_runner.py"), lib] + args + ["-q", "-x"],
                       env=env, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    tail = r.stdout[-3000:] + r.stderr[-2000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and " failed" not in r.stdout and " error" not in r.stdout.lower().replace("errors=", ""), tail


FULL = os.environ.get("DCVC_B200_DRY_FULL") == "1"
GPU_FILES_UNDER_EMULATION = [
    ("ops", ["tests/test_ops_gpu.py"], True),
    ("sequence-ld", ["tests/test_sequence_gpu.py", "-k", "ld"], True),
    ("sequence-hts", ["tests/test_sequence_gpu.py", "-k", "hts"], False),
    ("htl", ["tests/test_htl_gpu.py", "-k", "64-64 or oracle or bit_identical"], False),
    ("hts", ["tests/test_hts_gpu.py", "-k", "64-64 or oracle or bit_identical"], False),
    ("ld", ["tests/test_ld_gpu.py", "-k", "64-64 or oracle or bit_identical"], False),
    ("intra", ["tests/test_codec_gpu.py", "-k", "bit_exact and (256-256-32 or 64-64-0 or 200-328-63) or oracle"], False),
]


@pytest.mark.parametrize("name,args,default", GPU_FILES_UNDER_EMULATION, ids=[g[0] for g in GPU_FILES_UNDER_EMULATION])
def test_gpu_test_files_run_under_emulation(dry, name, args, default):
    if not default and not FULL:
        pytest.skip("set DCVC_B200_DRY_FULL=1 for the long emulated runs")
    _pytest_under_emulation(dry, args)


def test_smoke_entry_under_emulation(dry):
    """__graft_entry__.smoke() (the first thing the GPU box runs) end to end on the emulated runtime"""
    lib, shim = dry
    env = dict(os.environ)
    env.update({"LD_PRELOAD": shim, "LD_LIBRARY_PATH": CUDA_LIB + ":" + env.get("LD_LIBRARY_PATH", ""),
                "DCVC_B200_RANS_SPIN_US": "0", "DRY_SHIM_EMULATE": "1"})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests/dry_smoke_runner.py"), lib], env=env, capture_output=True,
                       text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "smoke-under-emulation ok" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
