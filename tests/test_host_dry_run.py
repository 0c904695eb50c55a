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
from concurrent.futures import ThreadPoolExecutor

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
    return _Dry(lib, shim)


class _Dry:
    """The emulated runs are independent subprocesses (each mostly one busy thread): every default job of this file is
    started when the first test asks for one and they run side by side; a test then only waits for its own job.
    Serially the file took ~7 minutes of the CPU suite."""

    def __init__(self, lib, shim):
        self.lib, self.shim = lib, shim
        self.pool = ThreadPoolExecutor(max_workers=max(2, min(8, (os.cpu_count() or 2) - 1)))
        self.jobs = {}

    def env(self, emulate, htl=True):
        env = dict(os.environ)
        env.update({"LD_PRELOAD": self.shim, "LD_LIBRARY_PATH": CUDA_LIB + ":" + env.get("LD_LIBRARY_PATH", ""),
                    "DCVC_B200_RANS_SPIN_US": "0", "OMP_NUM_THREADS": "2", "MKL_NUM_THREADS": "2", "DCVC_DRY_THREADS": "2",
                    "DCVC_B200_FUSE_TAIL": "0"})   # the shim emulates the per-op kernels only
        env.pop("DRY_SHIM_EMULATE", None)
        if emulate:
            env["DRY_SHIM_EMULATE"] = "1"
        return env

    def submit(self, key, cmd, env, timeout):
        if key not in self.jobs:
            self.jobs[key] = self.pool.submit(subprocess.run, cmd, env=env, capture_output=True, text=True,
                                              timeout=timeout, cwd=ROOT)

    def start_defaults(self):
        for codec, sizes in PLAN_JOBS:
            self.submit_flow("plan", codec, sizes)
        for codec, sizes in CHECK_JOBS:
            self.submit_flow("check", codec, sizes)
        self.submit_flow("check", "hts", ["72x104"], {"DCVC_B200_HEAD_LANES": "4"}, tag="lanes4")
        self.submit_flow("check", "hts", ["72x104"], {"DCVC_B200_HEAD_LANES": "1"}, tag="lanes1")
        self.submit_flow("plan", "hts", ["1080x1920", "2160x3840"], {"DCVC_B200_HEAD_LANES": "4"}, tag="lanes4")
        self.submit_flow("check", "hts", ["64x64"], {"DCVC_B200_HEAD_LANES": "2", "DCVC_B200_TEST_ALIAS_LANE_SCRATCH": "1"},
                         tag="lanes-racy")
        for name, args, default in GPU_FILES_UNDER_EMULATION:
            if default or FULL:
                self.submit_pytest(name, args)
        self.submit_smoke()
        self.submit_bench()

    def submit_bench(self):
        env = self.env(True, htl=False)
        env["DCVC_B200_BENCH_TEST_SIZE"] = "64x64"
        self.submit(("bench",), [sys.executable, os.path.join(ROOT, "tests/dry_bench_runner.py"), self.lib, "--steps", "1",
                                 "--warmup", "3", "--no-cpu-baseline", "--hts-size", "64x128"], env, 1200)

    def submit_flow(self, mode, codec, sizes, extra_env=None, tag=None):
        env = self.env(mode == "check")
        env.update(extra_env or {})
        self.submit((mode, codec) if tag is None else (mode, codec, tag),
                    [sys.executable, os.path.join(ROOT, "tests/dry_host_flow.py"), self.lib, self.shim, mode, codec] + sizes,
                    env, 1800)

    def submit_pytest(self, name, args):
        self.submit(("pytest", name), [sys.executable, os.path.join(ROOT, "tests/dry_pytest_runner.py"), self.lib] + args +
                    ["-q", "-x"], self.env(True), 3000)

    def submit_smoke(self):
        self.submit(("smoke",), [sys.executable, os.path.join(ROOT, "tests/dry_smoke_runner.py"), self.lib],
                    self.env(True, htl=False), 900)

    def result(self, key):
        self.start_defaults()
        return self.jobs[key].result()


def _run(dry, mode, codec, sizes):
    dry.submit_flow(mode, codec, sizes)
    r = dry.result((mode, codec))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])


# decode-side algorithmic GB and GMAC per picture (Intra, LD) / per 8-frame chunk (HT-S, HT-L) at 1088x1920, measured on
# the reference's own modules with forward hooks: the numerators of bench.py's roofline
REFERENCE_DECODE_WORK = {"intra": (5.89, 701.2), "hts": (14.09, 1483.4), "ld": (2.80, 147.8), "htl": (17.33, 2001.8)}


PLAN_SIZES = ["64x64", "200x328", "1080x1920", "2160x3840", "1096x1928"]
PLAN_JOBS = [(c, PLAN_SIZES) for c in ("intra", "hts", "ld", "htl")]
CHECK_JOBS = [("intra", ["64x64", "72x104"]), ("hts", ["72x104"]), ("ld", ["72x104"]), ("htl", ["72x104"])]


@pytest.mark.parametrize("codec", ["intra", "hts", "ld", "htl"])
def test_host_flow_plans_and_runs_at_all_sizes(dry, codec):
    res = _run(dry, "plan", codec, PLAN_SIZES)
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


@pytest.mark.parametrize("codec,sizes", CHECK_JOBS)
def test_emulated_codec_matches_its_oracle(dry, codec, sizes):
    res = _run(dry, "check", codec, sizes)
    for run in res["runs"]:
        as_list = lambda v: v if isinstance(v, list) else [v]  # noqa: E731
        for n, n_ref in zip(as_list(run["bytes"]), as_list(run["ref_bytes"])):
            assert abs(n - n_ref) <= 0.02 * n_ref + 8, (codec, run["size"], run["bytes"], run["ref_bytes"])
        for p, p_ref in zip(as_list(run["psnr"]), as_list(run["ref_psnr"])):
            assert abs(p - p_ref) <= 0.1, (codec, run["size"], p, p_ref)
        if "symbols" in run:
            for n, n_ref in zip(run["symbols"], run["ref_symbols"]):
                assert abs(n - n_ref) <= 0.01 * n_ref + 4


def test_recon_head_lanes_change_nothing_but_the_graph_shape(dry):
    """DCVC_B200_HEAD_LANES (default 2): the four recon-head pairs of the HT-S decoder are parallel branches of the recon
    graph, each on its own scratch level.  Under emulation the result must be the one of the single-lane run bit for bit
    (the branches share nothing but their read-only input), the capture must really fork (and join: the shim ends an
    unjoined capture with cudaErrorStreamCaptureUnjoined, like the runtime), and the arena estimate must cover the extra
    scratch levels at 1080p and 4K."""
    def last_json(r):
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
        return json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    base = last_json(dry.result(("check", "hts")))                 # the default: two lanes
    one = last_json(dry.result(("check", "hts", "lanes1")))
    four = last_json(dry.result(("check", "hts", "lanes4")))
    assert one["capture_forks"] == 0 and base["capture_forks"] >= 1 and four["capture_forks"] >= 3
    for other in (one, four):
        assert other["runs"][0]["bytes"] == base["runs"][0]["bytes"]
        assert other["runs"][0]["psnr"] == base["runs"][0]["psnr"]
    plan = last_json(dry.result(("plan", "hts", "lanes4")))
    assert plan["capture_forks"] >= 1
    assert all(run["arena_overflow_blocks"] == 0 for run in plan["runs"])
    # Emulation runs the branches one after the other, so equal results say nothing about races.  The shim therefore
    # collects what every kernel of a multi-lane graph reads and writes and fails the launch when a range written on
    # one branch overlaps a range touched on another: the runs above passed that check, and a deliberately broken
    # wiring (all lanes on lane 0's scratch, a test-only switch) must trip it.
    racy = dry.result(("check", "hts", "lanes-racy"))
    assert racy.returncode != 0 and "lane race" in racy.stderr, racy.stderr[-1500:]


def _pytest_under_emulation(dry, name, args):
    dry.submit_pytest(name, args)
    r = dry.result(("pytest", name))
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
    _pytest_under_emulation(dry, name, args)


def test_smoke_entry_under_emulation(dry):
    """__graft_entry__.smoke() (the first thing the GPU box runs) end to end on the emulated runtime"""
    dry.submit_smoke()
    r = dry.result(("smoke",))
    assert r.returncode == 0 and "smoke-under-emulation ok" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_bench_control_flow_and_json_contract_under_emulation(dry):
    """bench.py is run unattended by the driver at round end: its whole control flow (decode / e2e / encode legs, the
    per-family profile, the HT-S and LD legs) runs here on the emulated runtime at a tiny size with made-up timings,
    and the JSON line must carry every key of the contract.  The line says it is invalid as a measurement."""
    dry.submit_bench()
    r = dry.result(("bench",))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, "bench.py must print exactly one JSON line"
    d = json.loads(lines[0])
    assert d["INVALID_test_size_override"] == "64x64"
    for key, typ in (("metric", str), ("value", float), ("unit", str), ("n_gpus", int), ("steps", int), ("warmup", int),
                     ("ms_per_step", float), ("higher_is_better", bool), ("scaling", str), ("dtype", str), ("data", str),
                     ("config", dict), ("e2e", dict), ("gpu_launches", int), ("clocks", dict), ("roofline", dict)):
        assert isinstance(d[key], typ), (key, d.get(key))
    assert "vs_baseline" in d and d["vs_baseline"] is None           # BASELINE.md publishes nothing for this exact metric
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 3 and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(d["e2e"])
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(d["roofline"])
    assert d["roofline"]["bound"] == "hbm" and d["roofline"]["unit"] == "GB/s"
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(d["clocks"])
    assert d["gpu_launches"] > 100
    assert d["hts_extra"]["published_b200_reference_fps"] is None          # an unpublished size
    for leg in ("hts", "ld", "hts_extra"):
        assert "error" not in d[leg] and d[leg]["decode_fps"] > 0 and d[leg]["encode_fps"] > 0, d[leg]
