"""The drop-in boundary at the Python level: `inference_extensions_cuda` must expose the classes and methods the
reference's pybind module defines (src/layers/extensions/inference/bind.cpp:11-38), with the argument names / order of
the reference's call sites (image_model.py:194-217, video_model_ht.py:413-450, video_model_ld.py:273-306).
No device needed: the classes are only inspected, never constructed."""
import inspect
import os

import pytest

# bind.cpp:13-38 — class -> methods
REFERENCE_SURFACE = {
    "DMCIProxy": ["set_param", "compress", "decompress"],
    "DMCHTSProxy": ["set_param", "add_ref_feature_from_frame", "compress", "decompress"],
    "DMCHTLProxy": ["set_param", "add_ref_feature_from_frame", "compress", "decompress"],
    "DMCLDProxy": ["set_param", "add_ref_feature_from_frame", "compress", "decompress"],
}
# positional parameters after self, as the reference's Python callers pass them
CALL_SIGNATURES = {
    ("DMCIProxy", "set_param"): ["state_dict", "skip_threshold"],
    ("DMCIProxy", "compress"): ["x", "qp", "padding_b", "padding_r"],
    ("DMCIProxy", "decompress"): ["bit_stream", "qp", "height", "width", "ec_parallel"],
    ("DMCHTSProxy", "add_ref_feature_from_frame"): ["frame", "apply_adaptor"],
    ("DMCHTSProxy", "compress"): ["x", "qp", "reset_feature_memory", "padding_b", "padding_r"],
    ("DMCHTSProxy", "decompress"): ["bit_stream", "qp", "height", "width", "ec_parallel", "reset_feature_memory"],
    ("DMCHTLProxy", "add_ref_feature_from_frame"): ["frame", "apply_adaptor"],
    ("DMCHTLProxy", "compress"): ["x", "qp", "reset_feature_memory", "padding_b", "padding_r"],
    ("DMCHTLProxy", "decompress"): ["bit_stream", "qp", "height", "width", "ec_parallel", "reset_feature_memory"],
    ("DMCLDProxy", "add_ref_feature_from_frame"): ["frame", "apply_adaptor"],
    ("DMCLDProxy", "compress"): ["x", "qp", "reset_feature_memory", "padding_b", "padding_r"],
    ("DMCLDProxy", "decompress"): ["bit_stream", "qp", "height", "width", "ec_parallel", "reset_feature_memory"],
}


def test_plugin_exports_reference_classes_and_methods():
    import inference_extensions_cuda as ext
    for cls_name, methods in REFERENCE_SURFACE.items():
        cls = getattr(ext, cls_name)
        for m in methods:
            assert callable(getattr(cls, m)), f"{cls_name}.{m} missing"


@pytest.mark.parametrize("key", sorted(CALL_SIGNATURES))
def test_plugin_method_signatures(key):
    import inference_extensions_cuda as ext
    cls_name, method = key
    params = [p for p in inspect.signature(getattr(getattr(ext, cls_name), method)).parameters if p != "self"]
    assert params == CALL_SIGNATURES[key], (key, params)


def test_model_mirrors_keep_the_reference_api():
    """host-side mirrors of src/models: same method names and argument order as the reference models"""
    from dcvc_b200.model import DMC, DMCI, DMCLD
    sig = lambda f: [p for p in inspect.signature(f).parameters if p != "self"]  # noqa: E731
    assert sig(DMCI.compress) == ["x", "qp", "padding_b", "padding_r"]                      # image_model.py:194
    assert sig(DMCI.decompress) == ["bit_stream", "sps", "qp", "ec_part"]                   # image_model.py:214
    for cls in (DMC, DMCLD):
        assert sig(cls.add_ref_feature_from_frame) == ["frame", "apply_feature_adaptor"]    # video_model_ht.py:413
        assert sig(cls.compress) == ["x", "qp", "reset_feature_memory", "padding_b", "padding_r"]
        assert sig(cls.decompress) == ["bit_stream", "sps", "qp", "ec_part", "reset_feature_memory"]
        assert callable(cls.clear_dpb)
    assert DMCI.get_padding_size(1080, 1920, 16) == (0, 8)                                  # common_model.py:102-108
