// oracle/ref_shim.cpp — TEST INFRASTRUCTURE ONLY.
// A pybind11 binding around the *reference's own* RansEncoder/RansDecoder classes, compiled
// together with the reference sources where they lie (/root/reference/src/cpp/py_rans/{py_rans,
// rans}.cpp) by oracle/build_ref.py.  The reference's bind.cpp does not expose the decoded
// symbols to Python (get_decoded_tensor_cpp is C++-only, py_rans.h:64), so the oracle needs this
// extra getter to run the reference decoder.  No reference code is copied here.
#include "py_rans.h"

namespace py = pybind11;

PYBIND11_MODULE(dcvc_ref_rans, m)
{
    py::class_<RansEncoder>(m, "RansEncoder", py::module_local())   // module-local: MLCodec_extensions_cpp registers the same C++ types
        .def(py::init<>())
        .def("encode_y", py::overload_cast<const py::array_t<int16_t>&>(&RansEncoder::encode_y))
        .def("encode_z", py::overload_cast<const py::array_t<int8_t>&, const int, const int>(
                             &RansEncoder::encode_z))
        .def("flush", &RansEncoder::flush)
        .def("get_encoded_stream", &RansEncoder::get_encoded_stream)
        .def("reset", &RansEncoder::reset)
        .def("set_cdf",
             py::overload_cast<const py::array_t<int32_t>&, const py::array_t<int32_t>&, const int>(
                 &RansEncoder::set_cdf))
        .def("set_entropy_coder_parallel", &RansEncoder::set_entropy_coder_parallel);

    py::class_<RansDecoder>(m, "RansDecoder", py::module_local())
        .def(py::init<>())
        .def("set_stream", py::overload_cast<const py::array_t<uint8_t>&>(&RansDecoder::set_stream))
        .def("decode_y", py::overload_cast<const py::array_t<uint8_t>&>(&RansDecoder::decode_y))
        .def("decode_z", &RansDecoder::decode_z)
        .def("set_cdf",
             py::overload_cast<const py::array_t<int32_t>&, const py::array_t<int32_t>&, const int>(
                 &RansDecoder::set_cdf))
        .def("set_entropy_coder_parallel", &RansDecoder::set_entropy_coder_parallel)
        .def("get_decoded", [](RansDecoder& d, int n) {
            auto v = d.get_decoded_tensor_cpp();  // waits for the worker threads
            py::array_t<int8_t> out(n);
            std::copy(v->data(), v->data() + n, out.mutable_data());
            return out;
        });

    m.def("pmf_to_quantized_cdf", &pmf_to_quantized_cdf);
}
