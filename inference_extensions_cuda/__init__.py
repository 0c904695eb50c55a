"""Drop-in replacement for the reference's CUDA extension module `inference_extensions_cuda`
(src/layers/extensions/inference/bind.cpp:11-38), backed by libdcvc_b200.so.

Put the repository root on PYTHONPATH and the reference's own
`from inference_extensions_cuda import DMCIProxy` (src/models/image_model.py:197),
`from inference_extensions_cuda import DMCHTSProxy` / `DMCHTLProxy` (src/models/video_model_ht.py:420-423) and
`from inference_extensions_cuda import DMCLDProxy` (src/models/video_model_ld.py:279) resolve here
(tests/test_reference_surface_gpu.py drives the reference's unmodified models and test_video.py this way).
"""
from dcvc_b200.proxy import DMCHTLProxy, DMCHTSProxy, DMCIProxy, DMCLDProxy  # noqa: F401

__all__ = ["DMCIProxy", "DMCHTSProxy", "DMCHTLProxy", "DMCLDProxy"]
