// Generic-nfft spectral path (placeholder until the mixed-radix kernels land).
#include "common.cuh"

struct osz_spec_plan;

int osz_generic_create(void **state, int nfft) {
    (void)nfft;
    *state = nullptr;
    return osz::fail(OSZ_ERR_UNSUPPORTED,
                     "spectra: nfft must be a power of two in [256, 8192] in this build");
}
void osz_generic_destroy(void *state) { (void)state; }
int osz_generic_exec(void *state, const osz_spec_plan *p, int mode, const double *x, int64_t ldx,
                     int64_t rows, int64_t nseg, double *out, int64_t ldp, cudaStream_t st) {
    (void)state; (void)p; (void)mode; (void)x; (void)ldx; (void)rows; (void)nseg; (void)out;
    (void)ldp; (void)st;
    return osz::fail(OSZ_ERR_UNSUPPORTED, "spectra: generic nfft path not built");
}
