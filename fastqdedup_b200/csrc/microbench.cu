// microbench.cu -- the integer-issue peak of this GPU, measured the way SURVEY.md section 8(d)
// asks for it: dependent-free streams of LOP3 and POPC (the two instruction classes of the
// XOR+POPC Hamming compare, reference distances.h:8-31), timed with CUDA events.  The bench line
// divides (candidate pairs x integer operations per pair) by this number to report how far the
// compare phase is from the integer roofline.  Not on the product path.
#include "common.h"

namespace fqd {

namespace {

constexpr int CHAINS = 8;      // independent accumulators per thread: no instruction waits for another
constexpr int UNROLL = 16;

// MODE 0: LOP3 only; 1: POPC (+ the add that consumes it, not counted); 2: the compare mix of 4 LOP3 per POPC
template <int MODE>
__global__ void __launch_bounds__(256) int_peak_kernel(uint32_t *out, uint32_t seed, int iters)
{
    uint32_t acc[CHAINS];
    const uint32_t a = seed * (threadIdx.x + 1u), b = ~seed + blockIdx.x;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) acc[c] = seed ^ (uint32_t)(c * 0x9E3779B9u) ^ threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
#pragma unroll
            for (int c = 0; c < CHAINS; c++) {
                if (MODE == 0) {
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(acc[c]) : "r"(a), "r"(b));
                } else if (MODE == 1) {
                    uint32_t p;
                    asm volatile("popc.b32 %0, %1;" : "=r"(p) : "r"(acc[c]));
                    acc[c] += p;
                } else {
                    uint32_t p;
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(acc[c]) : "r"(a), "r"(b));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(acc[c]) : "r"(b), "r"(a));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x1E;" : "+r"(acc[c]) : "r"(a), "r"(b));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x78;" : "+r"(acc[c]) : "r"(b), "r"(a));
                    asm volatile("popc.b32 %0, %1;" : "=r"(p) : "r"(acc[c]));
                    acc[c] ^= p;
                }
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) r ^= acc[c];
    if (r == 0x12345678u) out[0] = r;   // keeps the chains alive
}

template <int MODE>
int time_mode(fqd_context *ctx, uint32_t *d_out, double ops_per_inner, double *ops_per_s)
{
    cudaStream_t s = ctx->stream;
    const int blocks = ctx->sm_count * 8, iters = 2048;
    int_peak_kernel<MODE><<<blocks, 256, 0, s>>>(d_out, 0x2545F491u, 64);   // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        FQD_CUDA(cudaEventRecord(ctx->ev[0], s));
        int_peak_kernel<MODE><<<blocks, 256, 0, s>>>(d_out, 0x2545F491u, iters);
        FQD_CUDA(cudaEventRecord(ctx->ev[1], s));
        FQD_CUDA(cudaEventSynchronize(ctx->ev[1]));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
        if (ms < best) best = ms;
    }
    FQD_CUDA(cudaGetLastError());
    const double ops = (double)blocks * 256.0 * iters * UNROLL * CHAINS * ops_per_inner;
    *ops_per_s = ops / (best * 1e-3);
    return FQD_OK;
}

}  // namespace

}  // namespace fqd

using namespace fqd;

extern "C" int fqd_int_peak(fqd_context *ctx, double *lop3_ops_per_s, double *popc_ops_per_s, double *mixed_ops_per_s)
{
    if (!ctx) { set_error("null context"); return FQD_ERR_ARG; }
    FQD_CUDA(cudaSetDevice(ctx->device));
    uint32_t *d_out = nullptr;
    FQD_CUDA(cudaMalloc(&d_out, 256));
    double v[3] = {0, 0, 0};
    int rc = time_mode<0>(ctx, d_out, 1.0, &v[0]);
    if (rc == FQD_OK) rc = time_mode<1>(ctx, d_out, 1.0, &v[1]);
    if (rc == FQD_OK) rc = time_mode<2>(ctx, d_out, 5.0, &v[2]);
    cudaFree(d_out);
    if (lop3_ops_per_s) *lop3_ops_per_s = v[0];
    if (popc_ops_per_s) *popc_ops_per_s = v[1];
    if (mixed_ops_per_s) *mixed_ops_per_s = v[2];
    return rc;
}
