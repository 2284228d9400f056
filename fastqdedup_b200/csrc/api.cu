// api.cu -- C ABI entry points of libfqd_b200.so (include/fqd_b200.h): context and
// memory plumbing, the batched clustering job and the function-level entry points that
// mirror _fastq.average_error_rate and _distance.within_distance.
#include <string.h>

#include <stdlib.h>

#include <algorithm>
#include <mutex>

#include "common.h"
#include "exchange.h"

namespace fqd {

static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof g_error, fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    const char *base = strrchr(file, '/');
    set_error("CUDA error %s (%s) at %s:%d in %s", cudaGetErrorName(e), cudaGetErrorString(e),
              base ? base + 1 : file, line, what);
    cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? FQD_ERR_NOMEM : FQD_ERR_CUDA;
}

int dev_alloc(fqd_context *ctx, size_t bytes, void **p)
{
    fqd_arena &a = ctx->arena;
    const size_t sz = (bytes + 255) & ~(size_t)255;
    *p = nullptr;
    if (a.base && a.off + sz <= a.cap) {
        *p = a.base + a.off;
    } else {
        // does not fit the slab (first job, or a bigger one): a plain allocation now, a
        // bigger slab at the next reset
        void *q = nullptr;
        FQD_CUDA(cudaMalloc(&q, sz));
        a.overflow.push_back(q);
        *p = q;
    }
    a.off += sz;
    if (a.off > a.high) a.high = a.off;
    return FQD_OK;
}

void dev_free(fqd_context *, void *) {}

int arena_reset(fqd_context *ctx)
{
    fqd_arena &a = ctx->arena;
    a.last_high = a.high > a.last_high || a.high > a.cap ? a.high : a.last_high;
    if (!a.overflow.empty() || a.high > a.cap) {
        FQD_CUDA(cudaStreamSynchronize(ctx->stream));
        for (void *q : a.overflow) cudaFree(q);
        a.overflow.clear();
        if (a.high > a.cap && !a.shared) {   // (a shared slab is only replaced by arena_reserve)
            if (a.base) cudaFree(a.base);
            a.base = nullptr;
            a.cap = 0;
            const size_t want = a.high + a.high / 8 + (64u << 20);
            void *q = nullptr;
            if (cudaMalloc(&q, want) == cudaSuccess) { a.base = (char *)q; a.cap = want; }
            else cudaGetLastError();   // stay on per-buffer allocations
        }
    }
    a.off = 0;
    a.high = 0;
    return FQD_OK;
}

int arena_reserve(fqd_context *ctx, size_t bytes)
{
    fqd_arena &a = ctx->arena;
    if (a.cap >= bytes) return FQD_OK;
    FQD_CUDA(cudaStreamSynchronize(ctx->stream));
    for (void *q : a.overflow) cudaFree(q);
    a.overflow.clear();
    if (a.base) cudaFree(a.base);
    a.base = nullptr;
    a.cap = 0;
    a.off = a.high = 0;
    a.generation++;
    void *q = nullptr;
    FQD_CUDA(cudaMalloc(&q, bytes));
    a.base = (char *)q;
    a.cap = bytes;
    return FQD_OK;
}

void arena_release(fqd_context *ctx, size_t mark)
{
    fqd_arena &a = ctx->arena;
    if (mark < a.off) a.off = mark;   // overflow chunks (if any) stay alive until the next reset
}

// Build the symbol coding for an alphabet (bytes in first-seen order, like the
// reference's Alphabet, _triemodule.c:32-67).  Codes are dense; rank is ASCII order + 1
// with PAD = 0 so that a proper prefix sorts first.
static int make_codec(const std::vector<uint8_t> &alphabet, bool varlen, Codec *c)
{
    memset(c->lut, 0xFF, sizeof c->lut);
    memset(c->rank, 0, sizeof c->rank);
    const int n = (int)alphabet.size();
    const int codes = n + (varlen ? 1 : 0);
    int need = 1;
    while ((1 << need) < codes) need++;
    const int bits = supported_bits(need);
    if (!bits) {
        set_error("alphabet of %d symbols is too large for this build", n);
        return FQD_ERR_UNSUPPORTED;
    }
    std::vector<uint8_t> sorted(alphabet);
    std::sort(sorted.begin(), sorted.end());
    c->swar = 0;
    // DNA: when every symbol is one of ACGTN use the table-free code of pack_key_acgtn
    bool dna = bits == 3;
    for (uint8_t ch : alphabet) dna = dna && (ch == 'A' || ch == 'C' || ch == 'G' || ch == 'T' || ch == 'N');
    if (dna) {
        const uint8_t letters[5] = {'A', 'C', 'G', 'N', 'T'};      // ASCII order
        for (int r = 0; r < 5; r++) {
            const uint8_t code = (letters[r] >> 1) & 7u;
            c->lut[letters[r]] = code;
            c->rank[code] = (uint8_t)(r + 1);
        }
        c->pad_code = SWAR_PAD_CODE;
        c->bits = 3;
        c->n_symbols = 5;
        c->varlen = varlen ? 1 : 0;
        // measured on B200 (100 M x 36 nt): the table-free packer is ~9 % slower than the shared-memory
        // table loop inside the latency-bound ingest kernel, so it is opt-in (FQD_SWAR=1)
        c->swar = getenv("FQD_SWAR") ? 1 : 0;
        return FQD_OK;
    }
    for (int i = 0; i < n; i++) {
        c->lut[alphabet[i]] = (uint8_t)i;
        const int r = (int)(std::lower_bound(sorted.begin(), sorted.end(), alphabet[i]) - sorted.begin());
        c->rank[i] = (uint8_t)std::min(255, r + 1);
    }
    c->pad_code = (uint8_t)std::min(255, n);   // only meaningful when varlen
    c->bits = (uint8_t)bits;
    c->n_symbols = (uint8_t)std::min(255, n);
    c->varlen = varlen ? 1 : 0;
    if (varlen && n >= (1 << bits)) {
        set_error("alphabet of %d symbols plus PAD does not fit %d bits", n, bits);
        return FQD_ERR_UNSUPPORTED;
    }
    return FQD_OK;
}

static int check_device(fqd_context *ctx)
{
    if (!ctx) { set_error("null context"); return FQD_ERR_ARG; }
    FQD_CUDA(cudaSetDevice(ctx->device));
    return FQD_OK;
}

}  // namespace fqd

using namespace fqd;

extern "C" {

const char *fqd_last_error(void) { return g_error; }

int fqd_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int fqd_context_create(int device_ordinal, fqd_context **out)
{
    if (!out) { set_error("null output pointer"); return FQD_ERR_ARG; }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error("no usable CUDA device (%s): fastqdedup_b200 has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return FQD_ERR_CUDA;
    }
    if (device_ordinal < 0 || device_ordinal >= n) {
        set_error("device ordinal %d out of range (0..%d)", device_ordinal, n - 1);
        return FQD_ERR_ARG;
    }
    fqd_context *ctx = new fqd_context();
    ctx->device = device_ordinal;
    FQD_CUDA(cudaSetDevice(device_ordinal));
    FQD_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    FQD_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    // The hot kernels touch random 32-byte sectors (hash slots, union-find parents): ask L2 not
    // to fetch whole 128-byte lines for them.  FQD_L2_FETCH overrides (32 / 64 / 128).
    {
        size_t gran = 32;
        if (const char *e = getenv("FQD_L2_FETCH")) gran = (size_t)atoi(e);
        if (gran == 32 || gran == 64 || gran == 128) {
            if (cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran) != cudaSuccess) cudaGetLastError();
        }
    }
    FQD_CUDA(cudaMalloc(&ctx->d_ctr, sizeof(DevCounters)));
    FQD_CUDA(cudaHostAlloc(&ctx->h_ctr, sizeof(DevCounters), cudaHostAllocDefault));
    for (auto &ev : ctx->ev) FQD_CUDA(cudaEventCreate(&ev));
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device_ordinal);
    *out = ctx;
    return FQD_OK;
}

void fqd_context_destroy(fqd_context *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (void *q : ctx->arena.overflow) cudaFree(q);
    if (ctx->arena.base) cudaFree(ctx->arena.base);
    for (auto &ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    if (ctx->d_ctr) cudaFree(ctx->d_ctr);
    if (ctx->h_ctr) cudaFreeHost(ctx->h_ctr);
    for (auto &ev : ctx->chunk_events) cudaEventDestroy(ev);
    for (int k = 0; k < 2; k++) {
        if (ctx->pack_stage[k]) cudaFreeHost(ctx->pack_stage[k]);
        if (ctx->pack_ev[k]) cudaEventDestroy(ctx->pack_ev[k]);
    }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int fqd_default_context(fqd_context **out)
{
    static std::mutex mu;
    static fqd_context *g_ctx = nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!g_ctx) {
        int dev = 0;
        const int n = fqd_device_count();
        const char *e = getenv("FQD_DEVICE");
        if (e && *e) dev = atoi(e);
        else if ((e = getenv("LOCAL_RANK")) && *e && n > 0) dev = atoi(e) % n;
        FQD_TRY(fqd_context_create(dev, &g_ctx));
    }
    *out = g_ctx;
    return FQD_OK;
}

int fqd_device_alloc(fqd_context *ctx, size_t bytes, void **dptr)
{
    FQD_TRY(check_device(ctx));
    FQD_CUDA(cudaMalloc(dptr, bytes ? bytes : 16));
    return FQD_OK;
}

int fqd_device_free(fqd_context *ctx, void *dptr)
{
    FQD_TRY(check_device(ctx));
    FQD_CUDA(cudaFree(dptr));
    return FQD_OK;
}

int fqd_device_upload(fqd_context *ctx, void *dptr, const void *host, size_t bytes)
{
    FQD_TRY(check_device(ctx));
    FQD_CUDA(cudaMemcpyAsync(dptr, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    FQD_CUDA(cudaStreamSynchronize(ctx->stream));
    return FQD_OK;
}

int fqd_device_download(fqd_context *ctx, void *host, const void *dptr, size_t bytes)
{
    FQD_TRY(check_device(ctx));
    FQD_CUDA(cudaMemcpyAsync(host, dptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    FQD_CUDA(cudaStreamSynchronize(ctx->stream));
    return FQD_OK;
}

int fqd_device_memset(fqd_context *ctx, void *dptr, int value, size_t bytes)
{
    FQD_TRY(check_device(ctx));
    FQD_CUDA(cudaMemsetAsync(dptr, value, bytes, ctx->stream));
    return FQD_OK;
}

int fqd_context_synchronize(fqd_context *ctx)
{
    FQD_TRY(check_device(ctx));
    FQD_CUDA(cudaStreamSynchronize(ctx->stream));
    return FQD_OK;
}

int fqd_host_alloc(size_t bytes, void **hptr)
{
    FQD_CUDA(cudaHostAlloc(hptr, bytes ? bytes : 16, cudaHostAllocDefault));
    return FQD_OK;
}

int fqd_host_free(void *hptr)
{
    FQD_CUDA(cudaFreeHost(hptr));
    return FQD_OK;
}

// ------------------------------------------------------------------------------------------
// fqd_cluster / fqd_cluster_sharded*
// ------------------------------------------------------------------------------------------

}  // extern "C"

namespace {

struct Resolved {
    DeviceJob dj;
    uint32_t len_min = 0, len_max = 0;
    uint32_t *host_bitmap = nullptr;   // where to copy the bitmap back to (HOST jobs)
    size_t bitmap_words = 0;
    cudaEvent_t h0 = nullptr, h1 = nullptr;   // owned: destroyed on every exit path of the callers
    Resolved() = default;
    Resolved(const Resolved &) = delete;
    Resolved &operator=(const Resolved &) = delete;
    ~Resolved() { drop_events(); }
    void drop_events()
    {
        if (h0) cudaEventDestroy(h0);
        if (h1) cudaEventDestroy(h1);
        h0 = h1 = nullptr;
    }
    void clear()
    {
        drop_events();
        dj = DeviceJob{};
        len_min = len_max = 0;
        host_bitmap = nullptr;
        bitmap_words = 0;
    }
};

int validate_job(const fqd_cluster_job *job)
{
    if (job->max_distance < 0) {
        set_error("max_distance should be non-negative");   // _triemodule.c:789-793
        return FQD_ERR_ARG;
    }
    if (job->method < 0 || job->method > 2) { set_error("unknown cluster dissection method %d", job->method); return FQD_ERR_ARG; }
    if (job->memory_space != FQD_MEM_HOST && job->memory_space != FQD_MEM_DEVICE) {
        set_error("unknown memory space %d", job->memory_space);
        return FQD_ERR_ARG;
    }
    if (job->n_records && !job->keys) { set_error("keys is NULL"); return FQD_ERR_ARG; }
    if (job->n_records >= 0xFFFFFFF0ull) { set_error("too many records for one job (%llu)", (unsigned long long)job->n_records); return FQD_ERR_UNSUPPORTED; }
    if (!job->key_offsets && job->key_length > job->key_stride) { set_error("key_length > key_stride"); return FQD_ERR_ARG; }
    return FQD_OK;
}

// Inputs -> device memory (arena), key length range of this shard.
int resolve_job(fqd_context *ctx, const fqd_cluster_job *job, uint32_t *keep_bitmap, Resolved &r)
{
    cudaStream_t s = ctx->stream;
    const uint64_t n = job->n_records;
    DeviceJob &dj = r.dj;
    ctx->h2d_bytes = 0;
    dj.n = n;
    dj.d = job->max_distance; dj.edit = job->use_edit_distance ? 1 : 0; dj.method = job->method;
    dj.max_err = job->max_average_error_rate;
    dj.filter_on = job->max_average_error_rate < 1.0 && job->quals != nullptr;   // __init__.py:235
    dj.phred_offset = job->phred_offset;
    dj.key_stride = job->key_stride; dj.key_len = job->key_length;
    dj.qual_stride = job->qual_stride; dj.qual_len = job->qual_length;
    if (dj.filter_on && !job->qual_offsets && job->qual_length > job->qual_stride) { set_error("qual_length > qual_stride"); return FQD_ERR_ARG; }
    r.bitmap_words = (size_t)((n + 31) / 32);
    FQD_CUDA(cudaEventCreate(&r.h0)); FQD_CUDA(cudaEventCreate(&r.h1));
    FQD_CUDA(cudaEventRecord(r.h0, s));
    if (job->memory_space == FQD_MEM_HOST) {
        auto up = [&](const void *src, size_t bytes, const void **dst) -> int {
            void *p = nullptr;
            FQD_TRY(dev_alloc(ctx, bytes ? bytes : 16, &p));
            if (bytes) FQD_CUDA(cudaMemcpyAsync(p, src, bytes, cudaMemcpyHostToDevice, s));
            ctx->h2d_bytes += bytes;
            *dst = p;
            return FQD_OK;
        };
        if (n) {
            size_t key_bytes;
            if (job->key_offsets) {
                key_bytes = (size_t)job->key_offsets[n];
                FQD_TRY(up(job->key_offsets, (n + 1) * 8, (const void **)&dj.key_off));
            } else {
                key_bytes = (size_t)n * job->key_stride;
                if (job->key_lengths) FQD_TRY(up(job->key_lengths, n * 4, (const void **)&dj.key_lens));
            }
            const bool stream_rows = !job->key_offsets && n >= (1u << 20) && !getenv("FQD_NO_OVERLAP");
            if (stream_rows) {
                void *p = nullptr;
                FQD_TRY(dev_alloc(ctx, key_bytes, &p));
                dj.keys = static_cast<const uint8_t *>(p);
                dj.host_keys = job->keys;
                dj.host_pack = !job->key_lengths && job->key_length == job->key_stride && (job->key_length & 3u) == 0 &&
                               job->key_length <= 64 && !getenv("FQD_NO_HOST_PACK");
            } else {
                FQD_TRY(up(job->keys, key_bytes, (const void **)&dj.keys));
            }
            if (dj.filter_on) {
                size_t qual_bytes;
                if (job->qual_offsets) {
                    qual_bytes = (size_t)job->qual_offsets[n];
                    FQD_TRY(up(job->qual_offsets, (n + 1) * 8, (const void **)&dj.qual_off));
                } else {
                    qual_bytes = (size_t)n * job->qual_stride;
                    if (job->qual_lengths) FQD_TRY(up(job->qual_lengths, n * 4, (const void **)&dj.qual_lens));
                }
                if (dj.host_keys && !job->qual_offsets) {
                    void *p = nullptr;
                    FQD_TRY(dev_alloc(ctx, qual_bytes, &p));
                    dj.quals = static_cast<const uint8_t *>(p);
                    dj.host_quals = job->quals;
                } else {
                    FQD_TRY(up(job->quals, qual_bytes, (const void **)&dj.quals));
                }
            }
            if (job->record_counts) FQD_TRY(up(job->record_counts, n * 4, (const void **)&dj.weights));
        }
        if (keep_bitmap) {
            void *p = nullptr;
            FQD_TRY(dev_alloc(ctx, std::max<size_t>(r.bitmap_words, 4) * 4, &p));
            dj.bitmap = static_cast<uint32_t *>(p);
            r.host_bitmap = keep_bitmap;
        }
    } else {
        dj.keys = job->keys; dj.key_off = job->key_offsets; dj.key_lens = job->key_offsets ? nullptr : job->key_lengths;
        if (dj.filter_on) { dj.quals = job->quals; dj.qual_off = job->qual_offsets; dj.qual_lens = job->qual_offsets ? nullptr : job->qual_lengths; }
        dj.bitmap = keep_bitmap;
        dj.weights = job->record_counts;
    }
    FQD_CUDA(cudaEventRecord(r.h1, s));

    // key length range: decides PW and whether PAD is needed
    if ((dj.key_off || dj.key_lens) && n) {
        DevCounters init{};
        init.len_min = 0xFFFFFFFFu;
        *ctx->h_ctr = init;
        FQD_CUDA(cudaMemcpyAsync(ctx->d_ctr, ctx->h_ctr, sizeof(DevCounters), cudaMemcpyHostToDevice, s));
        length_range_kernel<<<(unsigned)std::min<uint64_t>(ctx->sm_count * 8, (n + 255) / 256), 256, 0, s>>>(
            n, dj.key_off, dj.key_lens, ctx->d_ctr);
        FQD_CUDA(cudaGetLastError());
        FQD_CUDA(cudaMemcpyAsync(ctx->h_ctr, ctx->d_ctr, sizeof(DevCounters), cudaMemcpyDeviceToHost, s));
        FQD_CUDA(cudaStreamSynchronize(s));
        r.len_min = ctx->h_ctr->len_min;
        r.len_max = ctx->h_ctr->len_max;
        if (!dj.key_off && r.len_max > job->key_stride) { set_error("a key length exceeds key_stride"); return FQD_ERR_ARG; }
    } else {
        r.len_min = r.len_max = n ? job->key_length : 0;
    }
    return FQD_OK;
}

int parse_alphabet(const fqd_cluster_job *job, std::vector<uint8_t> &alphabet)
{
    const char *a = job->alphabet ? job->alphabet : "ACGTN";
    bool seen[256] = {};
    for (const char *p = a; *p; p++) {
        const uint8_t c = (uint8_t)*p;
        if (seen[c]) { set_error("Alphabet should consist of unique characters.Character %c was repeated. ", c); return FQD_ERR_ARG; }
        seen[c] = true;
        alphabet.push_back(c);
    }
    if (alphabet.empty()) alphabet.push_back('A');
    return FQD_OK;
}

int finish_job(fqd_context *ctx, Resolved &r, fqd_cluster_stats *stats)
{
    cudaStream_t s = ctx->stream;
    if (r.host_bitmap && r.bitmap_words)
        FQD_CUDA(cudaMemcpyAsync(r.host_bitmap, r.dj.bitmap, r.bitmap_words * 4, cudaMemcpyDeviceToHost, s));
    FQD_CUDA(cudaStreamSynchronize(s));
    stats->h2d_bytes = ctx->h2d_bytes;
    if (r.h0) {
        float t = 0.f;
        cudaEventElapsedTime(&t, r.h0, r.h1);
        stats->ms_h2d += t;
        r.drop_events();
    }
    return FQD_OK;
}

}  // namespace

struct fqd_comm {
    Exchange *ex = nullptr;
};

extern "C" {

int fqd_cluster(fqd_context *ctx, const fqd_cluster_job *job, fqd_cluster_stats *stats,
                uint32_t *keep_bitmap)
{
    FQD_TRY(check_device(ctx));
    if (!job || !stats) { set_error("null job/stats"); return FQD_ERR_ARG; }
    memset(stats, 0, sizeof *stats);
    FQD_TRY(validate_job(job));
    stats->total_records = job->n_records;
    ctx->res = fqd_result{};          // the previous job's result lives in the arena: gone now
    FQD_TRY(arena_reset(ctx));
    if (job->n_records == 0) return FQD_OK;
    Resolved r;
    FQD_TRY(resolve_job(ctx, job, keep_bitmap, r));
    r.dj.max_len = r.len_max;
    r.dj.varlen = r.len_min != r.len_max;
    std::vector<uint8_t> alphabet;
    FQD_TRY(parse_alphabet(job, alphabet));
    int rc = FQD_OK;
    const size_t inputs_mark = arena_mark(ctx);
    for (int attempt = 0; attempt < 3; attempt++) {
        arena_release(ctx, inputs_mark);
        Codec codec;
        FQD_TRY(make_codec(alphabet, r.dj.varlen, &codec));
        uint32_t unknown[8] = {};
        rc = run_pipeline(ctx, r.dj, codec, stats, unknown);
        if (rc != RC_RETRY_ALPHABET) break;
        for (int c = 0; c < 256; c++)
            if (unknown[c >> 5] & (1u << (c & 31))) alphabet.push_back((uint8_t)c);
    }
    if (rc == RC_RETRY_ALPHABET) { set_error("internal: alphabet did not converge"); rc = FQD_ERR_CUDA; }
    if (rc != FQD_OK) return rc;
    return finish_job(ctx, r, stats);
}

// every rank learns the `n` 64-bit values of every rank (NCCL ranks: a host-synchronous all-gather; ranks of
// this process: a copy)
static int gather_u64(fqd_context **ctxs, int n_local, Exchange *ex, int world, int n,
                      const std::vector<std::vector<uint64_t>> &mine, std::vector<uint64_t> &all)
{
    all.assign((size_t)world * n, 0);
    if (!ex) {
        for (int g = 0; g < n_local; g++)
            for (int i = 0; i < n; i++) all[(size_t)g * n + i] = mine[g][i];
        return FQD_OK;
    }
    fqd_context *ctx = ctxs[0];
    uint64_t *d_buf = nullptr;   // not from the arena: the slab may be replaced between two of these calls
    FQD_CUDA(cudaMalloc(&d_buf, (size_t)(world + 1) * n * 8));
    int rc = FQD_OK;
    if (cudaMemcpyAsync(d_buf, mine[0].data(), (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) rc = FQD_ERR_CUDA;
    if (rc == FQD_OK) rc = ex->allgather(d_buf, d_buf + n, (size_t)n * 8, ctx->stream);
    if (rc == FQD_OK && cudaMemcpyAsync(all.data(), d_buf + n, (size_t)world * n * 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
        rc = FQD_ERR_CUDA;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && rc == FQD_OK) rc = FQD_ERR_CUDA;
    cudaFree(d_buf);
    if (rc == FQD_ERR_CUDA) { cudaGetLastError(); set_error("CUDA failure in the rank agreement exchange"); }
    return rc;
}

// Views of the peers' arena slabs.  Ranks of this process: the slabs themselves (peer access enabled between
// different GPUs).  NCCL ranks: CUDA IPC handles, exchanged and (re)opened only when a slab changed.
static int map_peer_arenas(fqd_context **ctxs, int n_local, Exchange *ex, ShardWorld &W)
{
    W.peers_mapped = false;
    if (!ex) {
        for (int g = 0; g < n_local; g++) {
            W.peer_base[g] = ctxs[g]->arena.base;
            for (int r = 0; r < n_local; r++) {
                if (ctxs[r]->device == ctxs[g]->device) continue;
                int can = 0;
                FQD_CUDA(cudaDeviceCanAccessPeer(&can, ctxs[r]->device, ctxs[g]->device));
                if (!can) return FQD_OK;   // no peer path: the replicated-set plan runs instead
                FQD_CUDA(cudaSetDevice(ctxs[r]->device));
                const cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[g]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
                cudaGetLastError();
            }
        }
        W.peers_mapped = true;
        return FQD_OK;
    }
    fqd_context *ctx = ctxs[0];
    std::vector<std::vector<uint64_t>> mine(1, std::vector<uint64_t>(9, 0));
    cudaIpcMemHandle_t h;
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    bool ok = ctx->arena.base != nullptr && cudaIpcGetMemHandle(&h, ctx->arena.base) == cudaSuccess;
    if (!ok) cudaGetLastError();
    if (ok) memcpy(mine[0].data(), &h, 64);
    mine[0][8] = ok ? 1 : 0;
    std::vector<uint64_t> all;
    FQD_TRY(gather_u64(ctxs, n_local, ex, W.world, 9, mine, all));
    for (int g = 0; g < W.world; g++) ok = ok && all[(size_t)g * 9 + 8] == 1;
    if (!ok) return FQD_OK;            // some rank cannot export its slab: replicated-set plan
    ctx->arena.shared = true;
    for (int g = 0; g < W.world; g++) {
        if (g == ex->rank) { W.peer_base[g] = ctx->arena.base; continue; }
        const unsigned char *hb = reinterpret_cast<const unsigned char *>(&all[(size_t)g * 9]);
        if (ex->peer_open[g] && memcmp(ex->peer_handle[g], hb, 64) != 0) {
            cudaIpcCloseMemHandle(ex->peer_map[g]);
            ex->peer_open[g] = false;
        }
        if (!ex->peer_open[g]) {
            cudaIpcMemHandle_t ph;
            memcpy(&ph, hb, 64);
            void *p = nullptr;
            const cudaError_t e = cudaIpcOpenMemHandle(&p, ph, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                cudaGetLastError();
                if (getenv("FQD_TRACE")) fprintf(stderr, "[fqd trace] cudaIpcOpenMemHandle(rank %d): %s -> replicated-set plan\n", g, cudaGetErrorString(e));
                ok = false;
                continue;
            }
            ex->peer_map[g] = p;
            memcpy(ex->peer_handle[g], hb, 64);
            ex->peer_open[g] = true;
        }
        W.peer_base[g] = static_cast<char *>(ex->peer_map[g]);
    }
    // all ranks must agree on whether the views exist
    mine[0].assign(9, 0);
    mine[0][0] = ok ? 1 : 0;
    FQD_TRY(gather_u64(ctxs, n_local, ex, W.world, 9, mine, all));
    for (int g = 0; g < W.world; g++) ok = ok && all[(size_t)g * 9] == 1;
    W.peers_mapped = ok;
    return FQD_OK;
}

// Shared by the two sharded entry points: `n_local` shards of a `world`-rank job.
static int cluster_sharded_common(fqd_context **ctxs, int n_local, Exchange *ex, int world,
                                  const fqd_cluster_job *jobs, const uint64_t *index_bases,
                                  fqd_cluster_stats *stats, uint32_t **keep_bitmaps)
{
    std::vector<Resolved> R(n_local);
    std::vector<uint8_t> alphabet;
    FQD_TRY(parse_alphabet(&jobs[0], alphabet));
    ShardWorld W;
    W.world = world;
    uint32_t lmin = 0xFFFFFFFFu, lmax = 0;
    const bool want_tiles = world <= MAX_RANKS && !jobs[0].use_edit_distance && !getenv("FQD_SHARD_REPLICATED");
    for (int i = 0; i < n_local; i++) memset(&stats[i], 0, sizeof stats[i]);
    for (int attempt = 0;; attempt++) {
        lmin = 0xFFFFFFFFu; lmax = 0;
        int rc_local = FQD_OK;
        for (int i = 0; i < n_local && rc_local == FQD_OK; i++) {
            rc_local = check_device(ctxs[i]);
            if (rc_local == FQD_OK) rc_local = validate_job(&jobs[i]);
            if (rc_local == FQD_OK && index_bases[i] + jobs[i].n_records >= 0xFFFFFFF0ull) {
                set_error("global record index exceeds 32 bits");
                rc_local = FQD_ERR_UNSUPPORTED;
            }
            if (rc_local != FQD_OK) break;
            ctxs[i]->res = fqd_result{};
            R[i].clear();
            rc_local = arena_reset(ctxs[i]);
            if (rc_local == FQD_OK) rc_local = resolve_job(ctxs[i], &jobs[i], keep_bitmaps ? keep_bitmaps[i] : nullptr, R[i]);
            if (rc_local == FQD_OK && jobs[i].n_records) { lmin = std::min(lmin, R[i].len_min); lmax = std::max(lmax, R[i].len_max); }
        }
        // agreement: [rc, lmin, lmax, n, index base, arena capacity, arena offset after the inputs, previous high-water mark]
        std::vector<std::vector<uint64_t>> mine(n_local, std::vector<uint64_t>(8, 0));
        for (int i = 0; i < n_local; i++) {
            mine[i][0] = (uint64_t)rc_local;
            mine[i][1] = ex ? lmin : (jobs[i].n_records && rc_local == FQD_OK ? R[i].len_min : 0xFFFFFFFFu);
            mine[i][2] = ex ? lmax : (jobs[i].n_records && rc_local == FQD_OK ? R[i].len_max : 0);
            mine[i][3] = jobs[i].n_records;
            mine[i][4] = index_bases[i];
            mine[i][5] = ctxs[i]->arena.cap;
            mine[i][6] = ctxs[i]->arena.off;
            mine[i][7] = ctxs[i]->arena.last_high;
        }
        std::vector<uint64_t> all;
        FQD_TRY(gather_u64(ctxs, n_local, ex, world, 8, mine, all));
        int rc_any = FQD_OK;
        for (int g = 0; g < world; g++) if (all[(size_t)g * 8] != FQD_OK && rc_any == FQD_OK) rc_any = (int)all[(size_t)g * 8];
        if (rc_any != FQD_OK) {
            if (rc_local == FQD_OK) set_error("another rank of the sharded job failed before the plan started (status %d)", rc_any);
            return rc_local != FQD_OK ? rc_local : rc_any;
        }
        lmin = 0xFFFFFFFFu; lmax = 0;
        W.n_total = W.n_max = 0;
        size_t off = 0, prev_high = 0;
        for (int g = 0; g < world; g++) {
            const uint64_t *v = &all[(size_t)g * 8];
            if (v[3]) { lmin = std::min<uint32_t>(lmin, (uint32_t)v[1]); lmax = std::max<uint32_t>(lmax, (uint32_t)v[2]); }
            W.base[g] = v[4];
            W.base[g + 1] = v[4] + v[3];
            W.n_total += v[3];
            W.n_max = std::max(W.n_max, v[3]);
            off = std::max<size_t>(off, (size_t)v[6]);
            prev_high = std::max<size_t>(prev_high, (size_t)v[7]);
        }
        for (int g = 0; g + 1 < world; g++)
            if (all[(size_t)(g + 1) * 8 + 4] != W.base[g + 1]) { set_error("the shards of the ranks are not contiguous in rank order"); return FQD_ERR_ARG; }
        if (lmin == 0xFFFFFFFFu) lmin = lmax = 0;   // no records anywhere
        if (!want_tiles) break;
        // the tile-sharded plan lets ranks read each other's arena slabs: every slab must hold the whole job
        // (nothing in overflow chunks) and must not move while others look at it
        W.shared_off = (off + 4095) & ~(size_t)4095;
        size_t need = W.shared_off + tile_plan_bytes(W.n_total, W.n_max, world, jobs[0].max_distance, jobs[0].method);
        need = std::max(need, prev_high + prev_high / 16);
        bool any_grow = false;
        std::vector<bool> grows(world);
        for (int g = 0; g < world; g++) { grows[g] = (size_t)all[(size_t)g * 8 + 5] < need; any_grow = any_grow || grows[g]; }
        if (!any_grow) break;
        if (attempt >= 2) { set_error("internal: arena reservation did not converge"); return FQD_ERR_NOMEM; }
        // (1) everybody drops its view of the slabs about to be replaced, (2) barrier, (3) the owners replace them
        if (ex) {
            for (int g = 0; g < world; g++)
                if (grows[g] && ex->peer_open[g]) { cudaIpcCloseMemHandle(ex->peer_map[g]); ex->peer_open[g] = false; }
            std::vector<std::vector<uint64_t>> dummy(1, std::vector<uint64_t>(1, 0));
            FQD_TRY(gather_u64(ctxs, n_local, ex, world, 1, dummy, all));
        }
        for (int i = 0; i < n_local; i++) {
            const int g = ex ? ex->rank : i;
            if (!grows[g]) continue;
            FQD_CUDA(cudaSetDevice(ctxs[i]->device));
            // a little more than asked for, so that jobs of similar size do not trigger the protocol again
            int rc = arena_reserve(ctxs[i], need + need / 8);
            if (rc != FQD_OK) return rc;   // (the next agreement round would report it; out of memory is fatal anyway)
        }
    }
    if (want_tiles) FQD_TRY(map_peer_arenas(ctxs, n_local, ex, W));
    std::vector<DeviceJob> dj(n_local);
    std::vector<uint32_t> bases(n_local);
    std::vector<fqd_cluster_stats *> sp(n_local);
    std::vector<size_t> marks(n_local);
    for (int i = 0; i < n_local; i++) {
        R[i].dj.max_len = lmax;
        R[i].dj.varlen = lmin != lmax;
        bases[i] = (uint32_t)index_bases[i];
        sp[i] = &stats[i];
        marks[i] = arena_mark(ctxs[i]);
    }
    int rc = FQD_OK;
    bool replicated = !want_tiles || !W.peers_mapped;
    for (int attempt = 0; attempt < 4; attempt++) {
        for (int i = 0; i < n_local; i++) { arena_release(ctxs[i], marks[i]); dj[i] = R[i].dj; }
        Codec codec;
        FQD_TRY(make_codec(alphabet, lmin != lmax, &codec));
        uint32_t unknown[8] = {};
        rc = run_sharded(ctxs, dj.data(), bases.data(), sp.data(), n_local, ex, W, codec, lmax, unknown, replicated);
        if (rc == RC_FALLBACK_REPLICATED && !replicated) { replicated = true; continue; }
        if (rc != RC_RETRY_ALPHABET) break;
        for (int c = 0; c < 256; c++)
            if (unknown[c >> 5] & (1u << (c & 31))) alphabet.push_back((uint8_t)c);
    }
    if (rc == RC_RETRY_ALPHABET) { set_error("internal: alphabet did not converge"); rc = FQD_ERR_CUDA; }
    if (rc == RC_FALLBACK_REPLICATED) { set_error("internal: plan fallback did not converge"); rc = FQD_ERR_CUDA; }
    if (rc != FQD_OK) return rc;
    for (int i = 0; i < n_local; i++) {
        FQD_CUDA(cudaSetDevice(ctxs[i]->device));
        FQD_TRY(finish_job(ctxs[i], R[i], &stats[i]));
    }
    return FQD_OK;
}

int fqd_nccl_unique_id(uint8_t id[128]) { return nccl_unique_id(id); }

int fqd_comm_create(fqd_context *ctx, int rank, int world, const uint8_t id[128], fqd_comm **out)
{
    FQD_TRY(check_device(ctx));
    if (!out || world < 1 || rank < 0 || rank >= world) { set_error("bad rank/world"); return FQD_ERR_ARG; }
    if (world > 64) { set_error("at most 64 ranks"); return FQD_ERR_UNSUPPORTED; }
    Exchange *ex = nullptr;
    FQD_TRY(nccl_exchange_create(rank, world, id, &ex));
    *out = new fqd_comm{ex};
    return FQD_OK;
}

void fqd_comm_destroy(fqd_comm *comm)
{
    if (!comm) return;
    if (comm->ex)
        for (int g = 0; g < 64; g++)
            if (comm->ex->peer_open[g]) cudaIpcCloseMemHandle(comm->ex->peer_map[g]);
    delete comm->ex;
    delete comm;
}

int fqd_cluster_sharded(fqd_context *ctx, fqd_comm *comm, const fqd_cluster_job *job,
                        uint64_t index_base, fqd_cluster_stats *stats, uint32_t *keep_bitmap)
{
    if (!ctx || !comm || !job || !stats) { set_error("null argument"); return FQD_ERR_ARG; }
    uint32_t *bm[1] = {keep_bitmap};
    return cluster_sharded_common(&ctx, 1, comm->ex, comm->ex->world, job, &index_base, stats, bm);
}

int fqd_cluster_sharded_local(fqd_context **ctxs, int world, const fqd_cluster_job *jobs,
                              const uint64_t *index_bases, fqd_cluster_stats *stats,
                              uint32_t **keep_bitmaps)
{
    if (!ctxs || !jobs || !stats || !index_bases || world < 1) { set_error("null argument"); return FQD_ERR_ARG; }
    if (world > 64) { set_error("at most 64 ranks"); return FQD_ERR_UNSUPPORTED; }
    return cluster_sharded_common(ctxs, world, nullptr, world, jobs, index_bases, stats, keep_bitmaps);
}

int fqd_cluster_fetch(fqd_context *ctx, uint64_t *first, uint32_t *count, uint64_t *label,
                      uint8_t *selected)
{
    FQD_TRY(check_device(ctx));
    const uint32_t U = ctx->res.U;
    if (!U) return FQD_OK;
    cudaStream_t s = ctx->stream;
    std::vector<uint32_t> tmp(U);
    if (first || label) {
        FQD_CUDA(cudaMemcpyAsync(tmp.data(), ctx->res.ufirst, (size_t)U * 4, cudaMemcpyDeviceToHost, s));
        FQD_CUDA(cudaStreamSynchronize(s));
        if (first) for (uint32_t i = 0; i < U; i++) first[i] = tmp[i];
    }
    if (count) {
        FQD_CUDA(cudaMemcpyAsync(count, ctx->res.ucount, (size_t)U * 4, cudaMemcpyDeviceToHost, s));
    }
    if (selected) {
        FQD_CUDA(cudaMemcpyAsync(selected, ctx->res.selected, (size_t)U, cudaMemcpyDeviceToHost, s));
    }
    struct Scope { fqd_context *c; size_t m; ~Scope() { cudaStreamSynchronize(c->stream); arena_release(c, m); } } scope{ctx, arena_mark(ctx)};
    if (label && ctx->res.roots_only) {
        // tile-sharded job: this rank holds its own uniques only; label = root of the cluster in the job-wide id space
        DevBuf root;
        FQD_TRY(root.alloc(ctx, (size_t)U * 4));
        root_gid_kernel<<<(U + 255) / 256, 256, 0, s>>>(U, ctx->res.id_mul, ctx->res.id_add, ctx->res.parent_full, root.as<uint32_t>());
        FQD_CUDA(cudaGetLastError());
        std::vector<uint32_t> hroot(U);
        FQD_CUDA(cudaMemcpyAsync(hroot.data(), root.p, (size_t)U * 4, cudaMemcpyDeviceToHost, s));
        FQD_CUDA(cudaStreamSynchronize(s));
        for (uint32_t i = 0; i < U; i++) label[i] = hroot[i];
    } else if (label) {
        DevBuf minfirst, root;
        FQD_TRY(minfirst.alloc(ctx, (size_t)U * 4));
        FQD_TRY(root.alloc(ctx, (size_t)U * 4));
        FQD_CUDA(cudaMemsetAsync(minfirst.p, 0xFF, (size_t)U * 4, s));
        SelectParams sp{};
        sp.U = U; sp.ufirst = ctx->res.ufirst; sp.parent_full = ctx->res.parent_full;
        sp.root = root.as<uint32_t>(); sp.minfirst = minfirst.as<uint32_t>();
        label_min_kernel<<<(U + 255) / 256, 256, 0, s>>>(sp);
        FQD_CUDA(cudaGetLastError());
        std::vector<uint32_t> hroot(U), hmin(U);
        FQD_CUDA(cudaMemcpyAsync(hroot.data(), root.p, (size_t)U * 4, cudaMemcpyDeviceToHost, s));
        FQD_CUDA(cudaMemcpyAsync(hmin.data(), minfirst.p, (size_t)U * 4, cudaMemcpyDeviceToHost, s));
        FQD_CUDA(cudaStreamSynchronize(s));
        for (uint32_t i = 0; i < U; i++) label[i] = hmin[hroot[i]];
    }
    FQD_CUDA(cudaStreamSynchronize(s));
    return FQD_OK;
}

int fqd_cluster_fetch_selected(fqd_context *ctx, uint64_t *indices)
{
    FQD_TRY(check_device(ctx));
    const uint32_t U = ctx->res.U;
    if (!U) return FQD_OK;
    std::vector<uint32_t> f(U);
    std::vector<uint8_t> sel(U);
    FQD_CUDA(cudaMemcpyAsync(f.data(), ctx->res.ufirst, (size_t)U * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FQD_CUDA(cudaMemcpyAsync(sel.data(), ctx->res.selected, (size_t)U, cudaMemcpyDeviceToHost, ctx->stream));
    FQD_CUDA(cudaStreamSynchronize(ctx->stream));
    size_t k = 0;
    for (uint32_t i = 0; i < U; i++) if (sel[i]) indices[k++] = f[i];
    std::sort(indices, indices + k);
    return FQD_OK;
}

// ------------------------------------------------------------------------------------------
// function-level entry points
// ------------------------------------------------------------------------------------------

int fqd_average_error_rate(fqd_context *ctx, const uint8_t *phred, const uint64_t *offsets,
                           uint64_t n_strings, uint8_t phred_offset, double *out,
                           uint64_t *bad_index, uint32_t *bad_char)
{
    FQD_TRY(check_device(ctx));
    if (!n_strings) return FQD_OK;
    if (!offsets || !out) { set_error("null argument"); return FQD_ERR_ARG; }
    cudaStream_t s = ctx->stream;
    const size_t bytes = (size_t)offsets[n_strings];
    struct Scope { fqd_context *c; size_t m; ~Scope() { cudaStreamSynchronize(c->stream); arena_release(c, m); } } scope{ctx, arena_mark(ctx)};
    DevBuf d_ph, d_off, d_out;
    FQD_TRY(d_ph.alloc(ctx, bytes)); FQD_TRY(d_off.alloc(ctx, (n_strings + 1) * 8)); FQD_TRY(d_out.alloc(ctx, n_strings * 8));
    if (bytes) FQD_CUDA(cudaMemcpyAsync(d_ph.p, phred, bytes, cudaMemcpyHostToDevice, s));
    FQD_CUDA(cudaMemcpyAsync(d_off.p, offsets, (n_strings + 1) * 8, cudaMemcpyHostToDevice, s));
    DevCounters init{};
    init.phred_err = ~0ull;
    *ctx->h_ctr = init;
    FQD_CUDA(cudaMemcpyAsync(ctx->d_ctr, ctx->h_ctr, sizeof(DevCounters), cudaMemcpyHostToDevice, s));
    error_rate_kernel<<<(uint32_t)((n_strings + 127) / 128), 128, 0, s>>>(
        d_ph.as<uint8_t>(), d_off.as<uint64_t>(), n_strings, phred_offset, d_out.as<double>(), ctx->d_ctr);
    FQD_CUDA(cudaGetLastError());
    FQD_CUDA(cudaMemcpyAsync(out, d_out.p, n_strings * 8, cudaMemcpyDeviceToHost, s));
    FQD_CUDA(cudaMemcpyAsync(ctx->h_ctr, ctx->d_ctr, sizeof(DevCounters), cudaMemcpyDeviceToHost, s));
    FQD_CUDA(cudaStreamSynchronize(s));
    if (ctx->h_ctr->phred_err != ~0ull) {
        const uint32_t c = (uint32_t)(ctx->h_ctr->phred_err & 0xFF);
        if (bad_index) *bad_index = ctx->h_ctr->phred_err >> 8;
        if (bad_char) *bad_char = c;
        set_error("Character %c outside of valid phred range ('%c' to '%c')", (int)c, (int)phred_offset, 126);
        return FQD_ERR_PHRED;
    }
    return FQD_OK;
}

int fqd_within_distance(fqd_context *ctx, const uint8_t *a, const uint64_t *a_offsets,
                        const uint8_t *b, const uint64_t *b_offsets, uint64_t n_pairs,
                        int32_t max_distance, int32_t use_edit_distance, uint8_t *out)
{
    FQD_TRY(check_device(ctx));
    if (!n_pairs) return FQD_OK;
    if (!a_offsets || !b_offsets || !out) { set_error("null argument"); return FQD_ERR_ARG; }
    if (use_edit_distance && max_distance > MAX_BAND_D) {
        // a band wider than every string is the full matrix: clamp when that is exact
        uint64_t longest = 0;
        for (uint64_t i = 0; i < n_pairs; i++) {
            longest = std::max(longest, a_offsets[i + 1] - a_offsets[i]);
            longest = std::max(longest, b_offsets[i + 1] - b_offsets[i]);
        }
        if (longest > (uint64_t)MAX_BAND_D) {
            set_error("edit distances above %d are not supported for strings longer than that", MAX_BAND_D);
            return FQD_ERR_UNSUPPORTED;
        }
        max_distance = MAX_BAND_D;   // >= both lengths => always within distance, like the true predicate
    }
    cudaStream_t s = ctx->stream;
    const size_t ab = (size_t)a_offsets[n_pairs], bb = (size_t)b_offsets[n_pairs];
    struct Scope { fqd_context *c; size_t m; ~Scope() { cudaStreamSynchronize(c->stream); arena_release(c, m); } } scope{ctx, arena_mark(ctx)};
    DevBuf d_a, d_ao, d_b, d_bo, d_out;
    FQD_TRY(d_a.alloc(ctx, ab)); FQD_TRY(d_b.alloc(ctx, bb));
    FQD_TRY(d_ao.alloc(ctx, (n_pairs + 1) * 8)); FQD_TRY(d_bo.alloc(ctx, (n_pairs + 1) * 8));
    FQD_TRY(d_out.alloc(ctx, n_pairs));
    if (ab) FQD_CUDA(cudaMemcpyAsync(d_a.p, a, ab, cudaMemcpyHostToDevice, s));
    if (bb) FQD_CUDA(cudaMemcpyAsync(d_b.p, b, bb, cudaMemcpyHostToDevice, s));
    FQD_CUDA(cudaMemcpyAsync(d_ao.p, a_offsets, (n_pairs + 1) * 8, cudaMemcpyHostToDevice, s));
    FQD_CUDA(cudaMemcpyAsync(d_bo.p, b_offsets, (n_pairs + 1) * 8, cudaMemcpyHostToDevice, s));
    within_distance_kernel<<<(uint32_t)((n_pairs + 127) / 128), 128, 0, s>>>(
        d_a.as<uint8_t>(), d_ao.as<uint64_t>(), d_b.as<uint8_t>(), d_bo.as<uint64_t>(), n_pairs,
        max_distance, use_edit_distance ? 1 : 0, d_out.as<uint8_t>());
    FQD_CUDA(cudaGetLastError());
    FQD_CUDA(cudaMemcpyAsync(out, d_out.p, n_pairs, cudaMemcpyDeviceToHost, s));
    FQD_CUDA(cudaStreamSynchronize(s));
    return FQD_OK;
}

}  // extern "C"
