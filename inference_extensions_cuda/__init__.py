"""Drop-in replacement for the reference's CUDA extension module `inference_extensions_cuda`
(src/layers/extensions/inference/bind.cpp:11-38), backed by libdcvc_b200.so.

Put the repository root on PYTHONPATH and the reference's own
`from inference_extensions_cuda import DMCIProxy` (src/models/image_model.py:197) /
`from inference_extensions_cuda import DMCHTSProxy` (src/models/video_model_ht.py:420) resolve here.
`from inference_extensions_cuda import DMCLDProxy` (src/models/video_model_ld.py:279) too.
DMCHTLProxy is experimental (never run on a device yet): it is only importable with DCVC_B200_EXPERIMENTAL_HTL=1;
otherwise importing it raises ImportError, which the reference turns into its NotImplementedError
(video_model_ht.py:424-428).
"""
import os as _os

from dcvc_b200.proxy import DMCHTSProxy, DMCIProxy, DMCLDProxy  # noqa: F401

__all__ = ["DMCIProxy", "DMCHTSProxy", "DMCLDProxy"]

if _os.environ.get("DCVC_B200_EXPERIMENTAL_HTL") == "1":
    from dcvc_b200.proxy import DMCHTLProxy  # noqa: F401
    __all__.append("DMCHTLProxy")
