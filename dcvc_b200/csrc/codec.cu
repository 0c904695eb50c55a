// temporary stubs (codec runtime lands next)
#include "../../include/dcvc_b200.h"
extern "C" {
int dcvc_create(int32_t, int32_t, dcvc_codec**) { return 1; }
int dcvc_destroy(dcvc_codec*) { return 1; }
const char* dcvc_codec_error(dcvc_codec*) { return "not implemented"; }
int dcvc_set_param(dcvc_codec*, const char*, const void*, int32_t, int32_t, const int64_t*, int32_t) { return 1; }
int dcvc_finalize_params(dcvc_codec*, float) { return 1; }
int dcvc_compress(dcvc_codec*, const void*, int32_t, int32_t, int64_t, int64_t, int64_t, int32_t, int32_t, int32_t, void*, const uint8_t**, int32_t*, int32_t*, void*) { return 1; }
int dcvc_decompress(dcvc_codec*, const uint8_t*, int32_t, int32_t, int32_t, int32_t, int32_t, void*, void*) { return 1; }
int64_t dcvc_kernel_launches(dcvc_codec*) { return 0; }
int dcvc_last_gpu_ms(dcvc_codec*, float*) { return 1; }
int dcvc_debug_fetch(dcvc_codec*, const char*, void*, int64_t, int64_t*) { return 1; }
}
