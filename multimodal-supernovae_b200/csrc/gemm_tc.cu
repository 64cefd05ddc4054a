// tcgen05 / TMEM / TMA tensor-core GEMM (prec==1).  Placeholder until the kernel lands: reports "unsupported" so
// launch_gemm falls through to the FFMA kernel.
#include "common.cuh"
namespace mvn {
int launch_gemm_tc(const float*, const float*, float*, const int32_t*, int, int, int, bool, const GemmEpilogue&, cudaStream_t) {
    return MVN_E_UNSUPPORTED;
}
}  // namespace mvn
