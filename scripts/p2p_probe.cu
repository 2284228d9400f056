// p2p_probe.cu -- how fast can a kernel on GPU 0 move bytes from / to the HBM of GPU 1 over NVLink, by access
// method?  (Measurement aid for the tile-sharded multi-GPU plan; not part of the library.)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/p2p_probe scripts/p2p_probe.cu && gpurun_out/p2p_probe
//
// One process, two GPUs with peer access.  Every method moves the same number of bytes; "both" = GPU 1 runs the
// mirror-image kernel at the same time (what a sharded job does: every rank reads and is read).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr size_t BYTES = 1ull << 30;

// coalesced 16-byte loads, UNROLL independent loads in flight per thread
template <int UNROLL>
__global__ void __launch_bounds__(256) read_ldg128(const uint4 *__restrict__ src, size_t n16, uint32_t *sink)
{
    uint32_t acc = 0;
    const size_t stride = (size_t)gridDim.x * 256 * UNROLL;
    for (size_t i = (size_t)blockIdx.x * 256 * UNROLL + threadIdx.x; i + 256 * (UNROLL - 1) < n16; i += stride) {
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) asm volatile("ld.global.cs.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(src + i + 256 * u));
#pragma unroll
        for (int u = 0; u < UNROLL; u++) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x12345u) *sink = acc;
}

// 8-byte loads, one in flight per thread, a dependent dummy chain after each (like apply_edges_kernel)
__global__ void __launch_bounds__(256) read_ldg64_serial(const uint2 *__restrict__ src, size_t n8, uint32_t *sink)
{
    uint32_t acc = 0;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n8; i += (size_t)gridDim.x * 256) {
        uint2 v;
        asm volatile("ld.global.cs.v2.b32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(src + i));
        acc = acc * 2654435761u + (v.x ^ v.y);
    }
    if (acc == 0x12345u) *sink = acc;
}

// bulk asynchronous copies (TMA engine) of CH bytes into shared memory, `tiles` chunks at stride `stride` bytes
__global__ void __launch_bounds__(128) read_bulk(const char *src, size_t tiles, size_t stride, uint32_t ch, uint32_t *sink)
{
    extern __shared__ __align__(128) char smem[];
    __shared__ __align__(8) uint64_t mbar;
    const uint32_t mbar_a = (uint32_t)__cvta_generic_to_shared(&mbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar_a), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t phase = 0, acc = 0;
    for (size_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_a), "r"(ch) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(src + t * stride), "r"(ch), "r"(mbar_a) : "memory");
        }
        uint32_t done;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(mbar_a), "r"(phase) : "memory");
        } while (!done);
        phase ^= 1;
        acc ^= reinterpret_cast<uint32_t *>(smem)[threadIdx.x];
        __syncthreads();
    }
    if (acc == 0x12345u) *sink = acc;
}

// coalesced 16-byte stores
__global__ void __launch_bounds__(256) write_st128(uint4 *dst, size_t n16)
{
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n16; i += (size_t)gridDim.x * 256)
        asm volatile("st.global.cs.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(dst + i), "r"((uint32_t)i), "r"(1u), "r"(2u), "r"(3u) : "memory");
}

// scattered 32-byte records (what a partition kernel that pushes records to the owner of their tile does)
__global__ void __launch_bounds__(256) write_scatter32(uint32_t *dst, size_t n32, size_t count)
{
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < count; i += (size_t)gridDim.x * 256) {
        const size_t slot = (i * 0x9E3779B97F4A7C15ull >> 20) % n32;
        asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst + slot * 8), "r"((uint32_t)i), "r"(1u), "r"(2u), "r"(3u),
                     "r"(4u), "r"(5u), "r"(6u), "r"(7u) : "memory");
    }
}

// bulk asynchronous stores of CH bytes from shared memory
__global__ void __launch_bounds__(128) write_bulk(char *dst, size_t tiles, size_t stride, uint32_t ch)
{
    extern __shared__ __align__(128) char smem[];
    for (uint32_t i = threadIdx.x; i < ch / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    for (size_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        if (threadIdx.x == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + t * stride),
                         "r"((uint32_t)__cvta_generic_to_shared(smem)), "r"(ch) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

struct Side {
    int dev;
    char *mine, *peer;
    uint32_t *sink;
    cudaStream_t s;
    cudaEvent_t e0, e1;
};

template <typename F>
void run(const char *name, Side *sd, bool both, double bytes, F launch)
{
    for (int rep = 0; rep < 2; rep++) {
        for (int k = 0; k < (both ? 2 : 1); k++) {
            CK(cudaSetDevice(sd[k].dev));
            CK(cudaEventRecord(sd[k].e0, sd[k].s));
            launch(sd[k]);
            CK(cudaEventRecord(sd[k].e1, sd[k].s));
        }
        for (int k = 0; k < (both ? 2 : 1); k++) { CK(cudaSetDevice(sd[k].dev)); CK(cudaStreamSynchronize(sd[k].s)); CK(cudaGetLastError()); }
    }
    float ms = 0;
    CK(cudaSetDevice(sd[0].dev));
    CK(cudaEventElapsedTime(&ms, sd[0].e0, sd[0].e1));
    printf("%-44s %-5s %8.3f ms  %8.1f GB/s per GPU\n", name, both ? "both" : "one", ms, bytes / ms / 1e6);
    fflush(stdout);
}

int main()
{
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    if (n < 2) { printf("needs 2 GPUs\n"); return 0; }
    Side sd[2];
    for (int k = 0; k < 2; k++) {
        sd[k].dev = k;
        CK(cudaSetDevice(k));
        CK(cudaDeviceEnablePeerAccess(1 - k, 0));
        CK(cudaMalloc(&sd[k].mine, BYTES));
        CK(cudaMemset(sd[k].mine, 1, BYTES));
        CK(cudaMalloc(&sd[k].sink, 256));
        CK(cudaStreamCreate(&sd[k].s));
        CK(cudaEventCreate(&sd[k].e0));
        CK(cudaEventCreate(&sd[k].e1));
    }
    sd[0].peer = sd[1].mine;
    sd[1].peer = sd[0].mine;
    for (int k = 0; k < 2; k++) { CK(cudaSetDevice(k)); CK(cudaDeviceSynchronize()); }
    const int sms = 148;
    for (int local = 1; local >= 0; local--) {
        printf("---- %s memory ----\n", local ? "LOCAL (reference)" : "PEER (NVLink)");
        for (int both = 0; both <= (local ? 0 : 1); both++) {
            auto mem = [&](Side &x) { return local ? x.mine : x.peer; };
            run("read  LDG.128 coalesced, 1 in flight/thread", sd, both, (double)BYTES, [&](Side &x) { read_ldg128<1><<<sms * 8, 256, 0, x.s>>>((const uint4 *)mem(x), BYTES / 16, x.sink); });
            run("read  LDG.128 coalesced, 4 in flight/thread", sd, both, (double)BYTES, [&](Side &x) { read_ldg128<4><<<sms * 8, 256, 0, x.s>>>((const uint4 *)mem(x), BYTES / 16, x.sink); });
            run("read  LDG.128 coalesced, 8 in flight/thread", sd, both, (double)BYTES, [&](Side &x) { read_ldg128<8><<<sms * 8, 256, 0, x.s>>>((const uint4 *)mem(x), BYTES / 16, x.sink); });
            run("read  LDG.64 serial (apply_edges pattern)", sd, both, (double)BYTES / 8, [&](Side &x) { read_ldg64_serial<<<sms * 8, 256, 0, x.s>>>((const uint2 *)mem(x), BYTES / 64, x.sink); });
            const uint32_t chs[] = {1024, 4096, 16384};
            for (uint32_t ch : chs) {
                char nm[96];
                snprintf(nm, sizeof nm, "read  bulk (TMA) %5u B chunks, dense", ch);
                const int occ = ch <= 4096 ? 12 : 8;
                run(nm, sd, both, (double)BYTES, [&](Side &x) { read_bulk<<<sms * occ, 128, ch, x.s>>>(mem(x), BYTES / ch, ch, ch, x.sink); });
            }
            run("read  bulk (TMA)  1024 B of every 16 KB", sd, both, (double)BYTES / 16, [&](Side &x) { read_bulk<<<sms * 12, 128, 1024, x.s>>>(mem(x), BYTES / 16384, 16384, 1024, x.sink); });
            run("read  bulk (TMA)  4096 B of every 16 KB", sd, both, (double)BYTES / 4, [&](Side &x) { read_bulk<<<sms * 12, 128, 4096, x.s>>>(mem(x), BYTES / 16384, 16384, 4096, x.sink); });
            run("write ST.128 coalesced", sd, both, (double)BYTES, [&](Side &x) { write_st128<<<sms * 8, 256, 0, x.s>>>((uint4 *)mem(x), BYTES / 16); });
            run("write 32 B records scattered over 1 GB", sd, both, (double)BYTES / 4, [&](Side &x) { write_scatter32<<<sms * 8, 256, 0, x.s>>>((uint32_t *)mem(x), BYTES / 32, BYTES / 128); });
            for (uint32_t ch : chs) {
                char nm[96];
                snprintf(nm, sizeof nm, "write bulk (TMA) %5u B chunks, dense", ch);
                const int occ = ch <= 4096 ? 12 : 8;
                run(nm, sd, both, (double)BYTES, [&](Side &x) { write_bulk<<<sms * occ, 128, ch, x.s>>>(mem(x), BYTES / ch, ch, ch); });
            }
        }
    }
    return 0;
}
