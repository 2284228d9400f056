// common.h -- host-side plumbing shared by the translation units of libfqd_b200.so:
// error reporting across the C ABI, the per-GPU context, stream-ordered device buffers.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <string>
#include <vector>

#include "../../include/fqd_b200.h"
#include "pipeline.cuh"

namespace fqd {

void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define FQD_CUDA(...)                                                                    \
    do {                                                                                 \
        cudaError_t _e = (__VA_ARGS__);                                                  \
        if (_e != cudaSuccess) return ::fqd::cuda_fail(_e, #__VA_ARGS__, __FILE__, __LINE__); \
    } while (0)

#define FQD_TRY(...)                   \
    do {                               \
        int _rc = (__VA_ARGS__);       \
        if (_rc != FQD_OK) return _rc; \
    } while (0)

}  // namespace fqd

// Result of the last job, kept on the device until the next one.
struct fqd_result {
    uint32_t U = 0;
    uint64_t n_records = 0;
    uint64_t n_selected = 0;
    uint32_t *ufirst = nullptr;
    uint32_t *ucount = nullptr;
    uint32_t *parent_full = nullptr;
    uint8_t *selected = nullptr;
    // tile-sharded job: the rank's OWN uniques; unique u sits at slot id_add + u * id_mul of the (job-wide) forests,
    // and the per-unique view labels clusters by their root slot
    uint32_t id_mul = 1, id_add = 0;
    bool roots_only = false;
};

// Job-lifetime device memory: one slab, bump allocation, reset at the start of every job.
// Steady state makes no CUDA allocation calls at all (a 100 M-read job needs ~9 GB of
// scratch; growing a pool inside the timed region cost tens of ms).
struct fqd_arena {
    char *base = nullptr;
    size_t cap = 0;
    size_t off = 0;          // bump pointer (may exceed cap: then the job ran on overflow chunks)
    size_t high = 0;         // high-water mark of off over the last job
    std::vector<void *> overflow;
    // sharded jobs read each other's slabs (peer memory): once a slab has been shown to other ranks it is only
    // ever replaced through arena_reserve, between jobs, after every rank has dropped its view of it
    bool shared = false;
    uint64_t generation = 0; // bumped whenever the slab is replaced
    size_t last_high = 0;    // high-water mark of the previous job (sizes the next reservation)
};

struct fqd_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> chunk_events;
    fqd_arena arena;
    fqd::DevCounters *d_ctr = nullptr;
    fqd::DevCounters *h_ctr = nullptr;   // pinned
    cudaEvent_t ev[12] = {};
    // pinned staging of the host packer (two chunks of packed rows in flight) and the events that free a slot
    void *pack_stage[2] = {nullptr, nullptr};
    size_t pack_stage_bytes = 0;
    cudaEvent_t pack_ev[2] = {nullptr, nullptr};
    fqd_result res;
    uint64_t h2d_bytes = 0;      // input bytes the current job has copied host -> device
    int sm_count = 148;
};

namespace fqd {

// arena allocation (see fqd_arena); dev_free is a no-op kept for symmetry
int dev_alloc(fqd_context *ctx, size_t bytes, void **p);
void dev_free(fqd_context *ctx, void *p);
int arena_reset(fqd_context *ctx);                    // start of a job: everything is released
inline size_t arena_mark(fqd_context *ctx) { return ctx->arena.off; }
void arena_release(fqd_context *ctx, size_t mark);    // drop everything allocated after mark
int arena_reserve(fqd_context *ctx, size_t bytes);    // (empty arena only) make the slab at least this large

// RAII holder for job-lifetime device buffers
struct DevBuf {
    fqd_context *ctx = nullptr;
    void *p = nullptr;
    DevBuf() = default;
    explicit DevBuf(fqd_context *c) : ctx(c) {}
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { reset(); }
    int alloc(fqd_context *c, size_t bytes)
    {
        reset();
        ctx = c;
        return dev_alloc(c, bytes ? bytes : 16, &p);
    }
    void reset()
    {
        if (p) dev_free(ctx, p);
        p = nullptr;
    }
    void *release()
    {
        void *r = p;
        p = nullptr;
        return r;
    }
    template <typename T>
    T *as() const { return reinterpret_cast<T *>(p); }
};

// The job with every pointer resolved to device memory.
struct DeviceJob {
    uint64_t n = 0;
    const uint8_t *keys = nullptr;
    const uint64_t *key_off = nullptr;
    const uint32_t *key_lens = nullptr;
    uint32_t key_stride = 0, key_len = 0;
    const uint8_t *quals = nullptr;
    const uint64_t *qual_off = nullptr;
    const uint32_t *qual_lens = nullptr;
    uint32_t qual_stride = 0, qual_len = 0;
    int d = 1, edit = 0, method = 2;
    bool filter_on = false;
    double max_err = 1.0;
    uint32_t phred_offset = 33;
    uint32_t max_len = 0;
    bool varlen = false;
    const uint32_t *weights = nullptr;
    // HOST jobs with fixed-stride rows: the rows are still in host memory; stage_dedupe copies
    // them chunk by chunk on the copy stream and ingests each chunk as soon as it has landed
    // (keys / quals above are the device destinations)
    const uint8_t *host_keys = nullptr;
    bool host_pack = false;       // ... and the keys may cross PCIe packed (host_pack.cpp): fixed-length rows of <= 64 symbols
    const uint8_t *host_quals = nullptr;
    uint32_t *bitmap = nullptr;   // device, (n+31)/32 words, zeroed by the pipeline; may be null
};

constexpr int RC_RETRY_ALPHABET = -100;   // internal: unknown bytes were seen, grow the alphabet
constexpr int RC_FALLBACK_REPLICATED = -102;   // internal: the tile-sharded plan gave up (skew), run the replicated-set plan
constexpr int RC_PACK_INVALID = -103;          // internal: the host packer met a byte outside ACGTN, take the ASCII path
constexpr int RC_NOT_IN_GROUP = -101;     // internal: the (K, PW) instantiation lives in another instance group

// What every rank of a sharded job knows about the others before the plan starts (agreed in
// cluster_sharded_common, api.cu).
constexpr int MAX_WORLD = 64;
struct ShardWorld {
    int world = 1;
    uint64_t base[MAX_WORLD + 1] = {};   // rank g holds the records [base[g], base[g+1])
    uint64_t n_total = 0, n_max = 0;     // all records; the largest shard
    size_t shared_off = 0;               // arena offset (the same on every rank) where the plan's buffers start
    char *peer_base[MAX_WORLD] = {};     // arena slab of rank g as this process sees it (own slab, peer memory
                                         // mapped through CUDA IPC, or another context of this process)
    bool peers_mapped = false;           // false: no peer views (the tile-sharded plan cannot run)
};

// pipeline.cu: runs the stages for one (K, PW) instantiation
int run_pipeline(fqd_context *ctx, const DeviceJob &job, const Codec &codec,
                 fqd_cluster_stats *stats, uint32_t unknown_out[8]);
// sharded plan over `world` ranks; this process drives n_local of them (all of them as
// virtual ranks when ex == nullptr, exactly one over NCCL otherwise)
struct Exchange;
int run_sharded(fqd_context **ctxs, const DeviceJob *jobs, const uint32_t *index_base,
                fqd_cluster_stats **stats, int n_local, Exchange *ex, const ShardWorld &W, const Codec &codec,
                uint32_t max_len, uint32_t unknown_out[8], bool replicated_plan);
// whether a job can take the tile-sharded plan, and how much arena it needs past W.shared_off
bool tile_plan_eligible(const DeviceJob &job, const Codec &codec, uint32_t max_len, int world);
size_t tile_plan_bytes(uint64_t n_total, uint64_t n_max, int world, int d, int method);

inline uint32_t env_u32(const char *name, uint32_t fallback)
{
    const char *e = getenv(name);
    return e && *e ? (uint32_t)strtoul(e, nullptr, 10) : fallback;
}

// partitions of the streaming plan: regions of TILE_R records filled to ~60 % on average
inline uint32_t tile_partitions(uint64_t n)
{
    const uint32_t fill_pct = env_u32("FQD_TILE_FILL_PCT", 60);
    const uint32_t pct = fill_pct < 20 ? 20 : fill_pct > 90 ? 90 : fill_pct;
    const uint64_t parts = (n * 100 + (uint64_t)TILE_R * pct - 1) / ((uint64_t)TILE_R * pct);
    return (uint32_t)(parts < 1 ? 1 : parts);
}

// host_pack.cpp
// A set of CUDA events that is destroyed on every exit path (the launch plans return early through FQD_TRY).
struct EventSet {
    std::vector<cudaEvent_t> ev;
    EventSet() = default;
    EventSet(const EventSet &) = delete;
    EventSet &operator=(const EventSet &) = delete;
    ~EventSet()
    {
        for (cudaEvent_t e : ev)
            if (e) cudaEventDestroy(e);
    }
    cudaError_t create(size_t n)
    {
        ev.assign(n, nullptr);
        for (auto &e : ev) {
            const cudaError_t rc = cudaEventCreate(&e);
            if (rc != cudaSuccess) return rc;
        }
        return cudaSuccess;
    }
    cudaEvent_t &operator[](size_t i) { return ev[i]; }
};

uint32_t packed_row_words(uint32_t key_length);
uint64_t pack_keys_parallel(const uint8_t *src, uint64_t n, uint32_t L, uint32_t stride, uint32_t *dst);
int pack_threads();
uint64_t plane_stream_words(uint64_t n, uint32_t L);
uint64_t pack_planes_parallel(const uint8_t *src, uint64_t n, uint32_t L, uint64_t *dst);

// largest key (in symbols) this build can pack for a given number of code bits
uint32_t max_supported_length(int bits);
int supported_bits(int needed_bits);

}  // namespace fqd
