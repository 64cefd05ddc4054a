// A8/A9 ConvMixer -- kernels land next; entry points exist so the ABI is complete.
#include "common.cuh"
using namespace mvn;
extern "C" size_t mvn_conv_param_count(const mvn_conv_cfg*) { return 0; }
extern "C" size_t mvn_conv_workspace_bytes(const mvn_conv_cfg*) { return 0; }
extern "C" int mvn_conv_num_bn(const mvn_conv_cfg* c) { return c ? 1 + 2 * c->depth : 0; }
extern "C" int mvn_convmixer_fwd_stage(const mvn_conv_cfg*, int, const float*, const float*, float*, double*, float*, void*, size_t, void*) {
    set_error("convmixer: not built yet"); return MVN_E_UNSUPPORTED;
}
extern "C" int mvn_convmixer_bwd_stage(const mvn_conv_cfg*, int, const float*, const float*, const double*, double*, const float*, float*, void*, size_t, void*) {
    set_error("convmixer: not built yet"); return MVN_E_UNSUPPORTED;
}
