// affinity_tc.cu — stage 1 with the feature Gram matrix on the tcgen05 tensor cores.
// Placeholder until the TMA/tcgen05 kernel lands: reports EUNSUPPORTED so callers fail loudly.
#include "common.cuh"

namespace ancuts {

size_t affinity_tc_scratch_bytes(int n, int tdim, int ddim) {
    (void)n; (void)tdim; (void)ddim;
    return 0;
}

int launch_affinity_tc(int n, const double* pts, const float* tarl, int tdim, const float* dino, int ddim,
                       const uint8_t* tarl_zero, double alpha, double theta, double gamma, double prox,
                       float* W, long long ld, void* scratch, size_t scratch_bytes, cudaStream_t st) {
    (void)n; (void)pts; (void)tarl; (void)tdim; (void)dino; (void)ddim; (void)tarl_zero; (void)alpha; (void)theta;
    (void)gamma; (void)prox; (void)W; (void)ld; (void)scratch; (void)scratch_bytes; (void)st;
    set_error("affinity_impl=1 (tcgen05 Gram GEMM) is not built into this library yet");
    return ANCUTS_EUNSUPPORTED;
}

}  // namespace ancuts
