"""TEST INFRASTRUCTURE (tests/test_host_dry_run.py): makes CPU torch answer to the handful of `cuda` idioms the
Python layer of the package and the `-m gpu` tests use, inside the subprocess whose CUDA runtime is the dry-run shim
(tests/cpp/cuda_dry_shim.cpp).  With it the unmodified GPU test files — model mirrors -> inference_extensions_cuda ->
proxies -> C ABI -> (emulated) kernels — run on host memory, so that the Python half and the tests themselves are
exercised before they ever meet a device.  Never imported by the package."""
import types

import torch


def _cpu(dev):
    if dev is None:
        return None
    d = torch.device(dev) if not isinstance(dev, torch.device) else dev
    return torch.device("cpu") if d.type == "cuda" else d


def apply():
    for name in ("empty", "zeros", "ones", "full", "tensor", "arange", "randn", "rand", "empty_like", "zeros_like",
                 "ones_like", "as_tensor"):
        orig = getattr(torch, name)

        def wrap(*a, __orig=orig, **k):
            if "device" in k:
                k["device"] = _cpu(k["device"])
            return __orig(*a, **k)
        setattr(torch, name, wrap)

    orig_to = torch.Tensor.to

    def to(self, *a, **k):
        a = list(a)
        if a and isinstance(a[0], (str, torch.device)):
            a[0] = _cpu(a[0])
        if "device" in k:
            k["device"] = _cpu(k["device"])
        return orig_to(self, *a, **k)
    torch.Tensor.to = to
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.Tensor.is_cuda = property(lambda self: True)

    class _Stream:
        cuda_stream = 0

        def __init__(self, *a, **k):
            pass

        def synchronize(self):
            pass

        def wait_stream(self, *_):
            pass

    class _Ctx:
        def __init__(self, *a, **k):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    class _Event:
        def __init__(self, *a, **k):
            pass

        def record(self, *a):
            pass

        def synchronize(self):
            pass

        def elapsed_time(self, other):
            return 0.0

    c = torch.cuda
    c.is_available = lambda: True
    c.device_count = lambda: 1
    c.current_device = lambda: 0
    c.set_device = lambda *_: None
    c.synchronize = lambda *a, **k: None
    c.current_stream = lambda *a, **k: _Stream()
    c.Stream = _Stream
    c.stream = _Ctx
    c.Event = _Event
    c.empty_cache = lambda: None
    return types.SimpleNamespace(stream=_Stream)
