// fp64_peak.cu -- measured FP64 roofline denominator: independent DFMA chains per thread,
// every SM filled, timed with CUDA events on the caller's stream.
#include "kernels.cuh"

namespace brdae {

template <int CHAINS>
__global__ void __launch_bounds__(256) dfma_probe(double *out, int iters, double a, double b) {
    double x[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) x[i] = (double)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += x[i];
    if (s == 123.456) out[0] = s;  // keeps the chains alive; never true in practice
}

int fp64_peak_probe(double ms_target, double *flops, cudaStream_t s, int sms) {
    constexpr int CHAINS = 8, BLOCK = 256;
    if (sms <= 0) return BRDAE_ERR_CUDA;
    const int grid = sms * 8;
    double *out = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = BRDAE_ERR_CUDA;
    double best = 0;
    int iters = 20000;
    if (cudaMalloc(&out, sizeof(double)) != cudaSuccess) return BRDAE_ERR_CUDA;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) goto done;
    dfma_probe<CHAINS><<<grid, BLOCK, 0, s>>>(out, 1000, 1.0000001, 1e-9);  // warm-up
    for (int rep = 0; rep < 4; ++rep) {
        if (cudaEventRecord(e0, s) != cudaSuccess) goto done;
        dfma_probe<CHAINS><<<grid, BLOCK, 0, s>>>(out, iters, 1.0000001, 1e-9);
        if (cudaEventRecord(e1, s) != cudaSuccess || cudaEventSynchronize(e1) != cudaSuccess) goto done;
        float ms = 0;
        if (cudaEventElapsedTime(&ms, e0, e1) != cudaSuccess || ms <= 0) goto done;
        const double f = 2.0 * CHAINS * (double)iters * BLOCK * (double)grid / (ms * 1e-3);
        if (f > best) best = f;
        if (ms < ms_target) iters = (int)(iters * (ms_target / ms) * 1.1) + 1;  // grow towards the target duration
    }
    *flops = best;
    rc = BRDAE_OK;
done:
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(out);
    return rc;
}

}  // namespace brdae
