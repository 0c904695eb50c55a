"""Drop-in replacement for the reference's CUDA extension module `inference_extensions_cuda`
(src/layers/extensions/inference/bind.cpp:11-38), backed by libdcvc_b200.so.

Put the repository root on PYTHONPATH and the reference's own
`from inference_extensions_cuda import DMCIProxy` (src/models/image_model.py:197) resolves here.
"""
from dcvc_b200.proxy import DMCIProxy  # noqa: F401

__all__ = ["DMCIProxy"]
