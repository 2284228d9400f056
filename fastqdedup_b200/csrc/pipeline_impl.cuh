// pipeline_impl.cuh -- launch plans of the clustering job as host templates over the packed-key type
// <K bit planes, PW words per plane>: one B200 (run_typed) and sharded over several ranks
// (run_sharded_typed).  Included by pipeline_inst.cu, which is compiled once per instance group
// (instances.h) so the groups build in parallel.  Kernels are in pipeline.cuh; DESIGN.md has the data
// layout and the per-kernel rooflines.
#pragma once
#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <type_traits>
#include <vector>

#include "common.h"
#include "exchange.h"

namespace fqd {

namespace {

inline uint32_t cdiv(uint64_t a, uint32_t b) { return (uint32_t)((a + b - 1) / b); }

int fetch_counters(fqd_context *ctx)
{
    FQD_CUDA(cudaMemcpyAsync(ctx->h_ctr, ctx->d_ctr, sizeof(DevCounters), cudaMemcpyDeviceToHost,
                             ctx->stream));
    FQD_CUDA(cudaStreamSynchronize(ctx->stream));
    // fold the spread statistics of the tile kernels (the host copy then holds plain totals, which is
    // also what goes back to the device when a plan rewrites the counters)
    DevCounters *h = ctx->h_ctr;
    bool any = false;
    for (uint32_t k = 0; k < STAT_SPREAD; k++) {
        any |= h->cand_spread[k] != 0 || h->merge_spread[k] != 0;
        h->n_candidates += h->cand_spread[k];
        h->n_merges += h->merge_spread[k];
        h->cand_spread[k] = 0;
        h->merge_spread[k] = 0;
    }
    if (any) {   // keep the device copy consistent with the folded host copy
        FQD_CUDA(cudaMemcpyAsync(&ctx->d_ctr->n_candidates, &h->n_candidates, sizeof h->n_candidates, cudaMemcpyHostToDevice, ctx->stream));
        FQD_CUDA(cudaMemcpyAsync(&ctx->d_ctr->n_merges, &h->n_merges, sizeof h->n_merges, cudaMemcpyHostToDevice, ctx->stream));
        FQD_CUDA(cudaMemsetAsync(ctx->d_ctr->cand_spread, 0, sizeof h->cand_spread, ctx->stream));
        FQD_CUDA(cudaMemsetAsync(ctx->d_ctr->merge_spread, 0, sizeof h->merge_spread, ctx->stream));
    }
    return FQD_OK;
}

int reset_counters(fqd_context *ctx)
{
    DevCounters zero{};
    zero.phred_err = ~0ull;
    zero.len_min = 0xFFFFFFFFu;
    *ctx->h_ctr = zero;
    FQD_CUDA(cudaMemcpyAsync(ctx->d_ctr, ctx->h_ctr, sizeof(DevCounters), cudaMemcpyHostToDevice, ctx->stream));
    return FQD_OK;
}

int exclusive_scan_inplace(fqd_context *ctx, uint32_t *data, uint32_t n, uint32_t *block_sums,
                           uint32_t *grand_total)
{
    const uint32_t nblocks = cdiv(n, SCAN_TILE);
    if (nblocks > 1024u * SCAN_ITEMS) {
        set_error("internal: scan of %u items exceeds the two-level limit", n);
        return FQD_ERR_UNSUPPORTED;
    }
    scan_reduce_kernel<<<nblocks, SCAN_THREADS, 0, ctx->stream>>>(data, n, block_sums);
    scan_sums_kernel<<<1, 1024, 0, ctx->stream>>>(block_sums, nblocks, grand_total);
    scan_apply_kernel<<<nblocks, SCAN_THREADS, 0, ctx->stream>>>(data, n, block_sums, grand_total);
    FQD_CUDA(cudaGetLastError());
    return FQD_OK;
}

// ---- the dense unique set and the per-unique state of one rank -------------------------------

struct Uniques {
    uint32_t U = 0;
    uint32_t *ukey = nullptr, *ucount = nullptr, *ufirst = nullptr;
};

struct Forest {
    uint32_t *parent_full = nullptr, *parent_one = nullptr, *best = nullptr, *root = nullptr;
    uint8_t *dominated = nullptr, *dead = nullptr, *deadroot = nullptr, *selected = nullptr;
    uint2 *edges = nullptr;
    unsigned long long edge_cap = 0, n_edges = 0;
};

struct StageTimes {
    float table_clear = 0, ingest_kernel = 0, dedupe_kernel = 0, ingest = 0, compare = 0, h2d = 0;
    uint32_t launches = 0;
    bool streamed = false, partitioned = false, passes_partitioned = false, pass0_fused = false, pass1_emitted = false;
};

// Pass 0 of the Hamming search done inside the dedupe tiles (partitioned.cuh, FUSED).  The flag
// bytes and the edge list are allocated before the unique count is known (record-sized).
struct FusedPass0 {
    bool want = false;      // the job qualifies (single GPU, Hamming, d >= 1, not adjacency)
    bool done = false;      // the dedupe stage did it: pass 0 is complete, its edges wait in `edges`
    uint8_t *dominated = nullptr, *dead = nullptr;
    uint2 *edges = nullptr;
    uint32_t edge_cap = 0;
    uint32_t *aux = nullptr;   // [0] edge count, [1] overflow flag
    uint32_t spill_lo = 0, spill_hi = 0;   // unique ids that came out of the spill path (compared by brute force)
    // pass 1 pre-partitioned by the dedupe tiles (NextPass): tiles sized by a guess of the unique count
    uint32_t *next_buf = nullptr, *next_cursor = nullptr;
    uint32_t next_nparts = 0;
    bool next_valid = false;
};
constexpr uint32_t FUSED_SPILL_BRUTE = 4096;   // at most this many spilled uniques are compared by brute force

template <typename T>
int arena(fqd_context *ctx, size_t count, T **p)
{
    void *q = nullptr;
    FQD_TRY(dev_alloc(ctx, std::max<size_t>(count * sizeof(T), 16), &q));
    *p = static_cast<T *>(q);
    return FQD_OK;
}

// ---- stage 1: filter + pack + exact dedupe of this rank's records ----------------------------
// Produces the dense unique arrays of `uq`.  Large jobs take the streaming plan (records
// partitioned into shared-memory sized tiles, partitioned.cuh); small ones, and jobs whose
// duplication is so skewed that even the spill buffer overflows, take the single-table plan.

// Launches `kernel(params)` over the records: in one go when they are in HBM, chunk by chunk
// behind the H2D copies when they are still in host memory.
template <typename Launch>
int for_each_input_chunk(fqd_context *ctx, const DeviceJob &job, IngestParams ip, uint32_t index_base,
                         uint32_t *keepmask, StageTimes &tt, Launch launch)
{
    cudaStream_t s = ctx->stream;
    const uint64_t n = job.n;
    if (!n) return FQD_OK;
    if (!job.host_keys) {
        launch(ip);
        tt.launches++;
        return FQD_OK;
    }
    // H2D of chunk i+1 overlaps the work on chunk i (PCIe is the e2e bottleneck)
    const uint64_t chunk = 4u << 20;   // records; a multiple of every block tile and of 32
    const size_t nchunks = (size_t)((n + chunk - 1) / chunk);
    while (ctx->chunk_events.size() < nchunks) {
        cudaEvent_t e;
        FQD_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->chunk_events.push_back(e);
    }
    FQD_CUDA(cudaEventRecord(ctx->ev[9], s));
    FQD_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev[9], 0));   // buffers are allocated / idle
    FQD_CUDA(cudaEventRecord(ctx->ev[10], ctx->copy_stream));
    for (size_t i = 0; i < nchunks; i++) {
        const uint64_t c0 = i * chunk, cn = std::min<uint64_t>(chunk, n - c0);
        FQD_CUDA(cudaMemcpyAsync(const_cast<uint8_t *>(job.keys) + c0 * job.key_stride,
                                 job.host_keys + c0 * job.key_stride, cn * job.key_stride,
                                 cudaMemcpyHostToDevice, ctx->copy_stream));
        if (job.host_quals)
            FQD_CUDA(cudaMemcpyAsync(const_cast<uint8_t *>(job.quals) + c0 * job.qual_stride,
                                     job.host_quals + c0 * job.qual_stride, cn * job.qual_stride,
                                     cudaMemcpyHostToDevice, ctx->copy_stream));
        ctx->h2d_bytes += cn * job.key_stride + (job.host_quals ? cn * job.qual_stride : 0);
        FQD_CUDA(cudaEventRecord(ctx->chunk_events[i], ctx->copy_stream));
        FQD_CUDA(cudaStreamWaitEvent(s, ctx->chunk_events[i], 0));
        IngestParams cp = ip;
        cp.n = cn;
        cp.keys = job.keys + c0 * job.key_stride;
        cp.key_lens = job.key_lens ? job.key_lens + c0 : nullptr;
        cp.quals = job.quals ? job.quals + c0 * job.qual_stride : nullptr;
        cp.qual_lens = job.qual_lens ? job.qual_lens + c0 : nullptr;
        cp.keepmask = keepmask ? keepmask + c0 / 32 : nullptr;
        cp.weights = job.weights ? job.weights + c0 : nullptr;
        cp.index_base = index_base + (uint32_t)c0;
        launch(cp);
        tt.launches++;
    }
    FQD_CUDA(cudaEventRecord(ctx->ev[11], ctx->copy_stream));
    tt.streamed = true;
    return FQD_OK;
}

// the ASCII-bits code of pack_key_acgtn (api.cu make_codec): the table-free packer applies
bool codec_is_dna(const Codec &c)
{
    return c.bits == 3 && c.n_symbols == 5 && c.pad_code == SWAR_PAD_CODE && c.lut['A'] == 0 && c.lut['C'] == 1 &&
           c.lut['T'] == 2 && c.lut['G'] == 3 && c.lut['N'] == 7;
}

// Split launch of a tile kernel (one GPU): tiles of up to TILE_R_SMALL records go through the instance with the small
// shared-memory tile (12 resident blocks per SM instead of 10 -- the tile kernels wait on barriers and round trips,
// more tiles in flight is what makes them faster), the rest through the full-size instance, whose blocks loop over
// the list classify_tiles_kernel wrote.  FQD_TILE_SPLIT=0: one full-size launch for all tiles.
inline bool tile_split_enabled(uint32_t nparts)
{
    const char *e = getenv("FQD_TILE_SPLIT");
    if (e) return atoi(e) != 0;
    return nparts >= 4096;   // (small jobs: one launch, the list is not worth a kernel)
}
struct TileSplit {
    uint32_t *list = nullptr, *n_list = nullptr;
};
inline int tile_split_prepare(fqd_context *ctx, const PartParams &q, TileSplit &ts, StageTimes &tt)
{
    FQD_TRY(arena(ctx, (size_t)q.nparts, &ts.list));
    FQD_TRY(arena(ctx, 4, &ts.n_list));
    FQD_CUDA(cudaMemsetAsync(ts.n_list, 0, 4, ctx->stream));
    classify_tiles_kernel<<<cdiv(q.nparts, 256), 256, 0, ctx->stream>>>(q.cursor, q.nparts, (uint32_t)TILE_R_SMALL, ts.list, ts.n_list);
    tt.launches++;
    return FQD_OK;
}
inline TileSource small_tiles(TileSource S) { S.cnt_lo = 0; S.cnt_hi = TILE_R_SMALL; return S; }
inline TileSource large_tiles(TileSource S, const TileSplit &ts)
{
    S.cnt_lo = TILE_R_SMALL; S.cnt_hi = TILE_R; S.list = ts.list; S.n_list = ts.n_list;
    return S;
}
constexpr int TILE_SMALL_BLOCKS = 12;   // resident blocks per SM the small-tile instances are compiled for

constexpr uint64_t PARTITION_MIN_RECORDS = 4u << 20;
constexpr uint64_t PARTITION_MIN_UNIQUES = 1u << 20;

// One GPU: a tile has one source, the partition buffer of this GPU.
inline TileSource single_source(const PartParams &q)
{
    TileSource S{};
    S.buf[0] = q.buf;
    S.cnt = q.cursor; S.cnt_stride = q.nparts;
    S.G = 1; S.self = 0; S.first_tile = 0; S.ntiles = q.nparts; S.region = q.region;
    return S;
}

inline EdgeSource single_edges(const uint2 *edges, const uint32_t *n_edges, uint32_t cap)
{
    EdgeSource E{};
    E.edges[0] = edges; E.cap[0] = cap; E.n_edges = n_edges; E.n_stride = 1; E.G = 1; E.self = 0;
    return E;
}

// HOST jobs with fixed-length ACGTN rows: the host packs chunk i+1 (threads, AVX-512: three plane streams per chunk,
// host_pack.cpp) while chunk i crosses PCIe at 3 bits per symbol and partition_planes_kernel unpacks the chunk before that.  RC_PACK_INVALID: a byte outside ACGTN
// (the caller clears the partition buffers and takes the ASCII path, which reports it).
//
// Packing is the slower of the two legs on the hosts measured (16 threads packed 100 M x 36 nt rows in 54 ms in the
// per-row format of round 2's first version, ~40 ms as plane streams; the packed chunk crosses PCIe in 25 ms), so the
// link would idle a good part of the time.  Hybrid: whenever the copies queued so far
// would drain before the next chunk is packed, that chunk is sent as it is (ASCII, straight from the caller's
// buffer -- no host work at all) and goes through the ASCII partition kernel; the packer threads never wait and the
// link carries raw rows in what would be its idle time.  FQD_HOST_PACK_HYBRID=0 packs every chunk.
template <int PW, int NW, typename LaunchAscii>
int launch_packed_chunks(fqd_context *ctx, const DeviceJob &job, const IngestParams &pp, uint32_t index_base, StageTimes &tt,
                         LaunchAscii launch_ascii)
{
    cudaStream_t s = ctx->stream;
    const uint64_t n = job.n;
    const uint32_t L = job.key_len;
    const uint64_t chunk = 4u << 20;   // records; a multiple of every block tile
    const size_t nchunks = (size_t)((n + chunk - 1) / chunk);
    // a chunk crosses PCIe as three plane streams (host_pack.cpp): 3 bits per symbol
    const size_t full_stage = (size_t)3 * plane_stream_words(chunk, L) * 8;
    const size_t stage_bytes = (size_t)3 * plane_stream_words(std::min<uint64_t>(chunk, n), L) * 8;
    if (ctx->pack_stage_bytes < stage_bytes) {
        for (int k = 0; k < 2; k++) {
            if (ctx->pack_stage[k]) FQD_CUDA(cudaFreeHost(ctx->pack_stage[k]));
            ctx->pack_stage[k] = nullptr;
        }
        ctx->pack_stage_bytes = 0;
        for (int k = 0; k < 2; k++) FQD_CUDA(cudaHostAlloc(&ctx->pack_stage[k], full_stage, cudaHostAllocDefault));
        ctx->pack_stage_bytes = full_stage;
    }
    for (int k = 0; k < 2; k++)
        if (!ctx->pack_ev[k]) FQD_CUDA(cudaEventCreateWithFlags(&ctx->pack_ev[k], cudaEventDisableTiming));
    while (ctx->chunk_events.size() < nchunks) {
        cudaEvent_t e;
        FQD_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->chunk_events.push_back(e);
    }
    // raw chunks need the caller's buffer to be page-locked (a pageable source makes the copy synchronous and slow)
    bool hybrid = !(getenv("FQD_HOST_PACK_HYBRID") && atoi(getenv("FQD_HOST_PACK_HYBRID")) == 0) && nchunks > 1;
    if (hybrid) {
        cudaPointerAttributes at{};
        hybrid = cudaPointerGetAttributes(&at, job.host_keys) == cudaSuccess && at.type == cudaMemoryTypeHost;
        cudaGetLastError();
    }
    FQD_CUDA(cudaEventRecord(ctx->ev[9], s));
    FQD_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev[9], 0));   // buffers are allocated / idle
    FQD_CUDA(cudaEventRecord(ctx->ev[10], ctx->copy_stream));
    uint8_t *dev = const_cast<uint8_t *>(job.keys);   // every chunk lands inside its own range of the ASCII buffer
    using clock = std::chrono::steady_clock;
    auto now = [&] { return std::chrono::duration<double>(clock::now().time_since_epoch()).count(); };
    double link_free_at = now();                      // estimate of when the copies queued so far are done
    // priors, refined below from what this job measures: ~50 GB/s over the link, ~6 ns per 36-nt row and packer thread
    // (a rank of a sharded job only has its share of the cores: few threads, and most chunks travel raw)
    double link_bps = 50e9, pack_s = (double)chunk * 6e-9 * ((double)L / 36.0) / (double)std::max(1, pack_threads());
    size_t packed_slot = 0, last_copy = (size_t)-1;
    double first_copy_t0 = 0.0;
    size_t first_copy_bytes = 0;
    bool first_copy_timed = false;
    for (size_t i = 0; i < nchunks; i++) {
        const uint64_t c0 = i * chunk, cn = std::min<uint64_t>(chunk, n - c0);
        double t = now();
        if (last_copy != (size_t)-1 && cudaEventQuery(ctx->chunk_events[last_copy]) == cudaSuccess) {
            if (!first_copy_timed && last_copy == 0 && t > first_copy_t0) {   // (an upper bound of its duration: good enough)
                link_bps = std::max(10e9, std::min(64e9, (double)first_copy_bytes / (t - first_copy_t0)));
                first_copy_timed = true;
            }
            link_free_at = std::min(link_free_at, t);   // the link has drained
        }
        // send raw when the link would run dry while this chunk is being packed
        const double est_pack = pack_s * (double)cn / (double)chunk;
        const uint64_t words = plane_stream_words(cn, L);   // per plane
        const bool tiny = (size_t)3 * words * 8 > (size_t)cn * L;   // (a handful of rows: the streams would not fit the chunk's range)
        const bool raw = tiny || (hybrid && (i == 0 || link_free_at - t < est_pack));
        IngestParams cp = pp;
        cp.n = cn;
        cp.index_base = index_base + (uint32_t)c0;
        size_t bytes;
        if (raw) {
            bytes = (size_t)cn * L;
            FQD_CUDA(cudaMemcpyAsync(dev + c0 * L, job.host_keys + c0 * job.key_stride, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
            cp.keys = dev + c0 * L;
        } else {
            const int slot = (int)(packed_slot++ & 1);
            if (packed_slot > 2) FQD_CUDA(cudaEventSynchronize(ctx->pack_ev[slot]));   // the copy out of this slot is done
            uint64_t *stage = static_cast<uint64_t *>(ctx->pack_stage[slot]);
            const double p0 = now();
            if (pack_planes_parallel(job.host_keys + c0 * job.key_stride, cn, L, stage) != cn) {   // (key_stride == L here)
                FQD_CUDA(cudaStreamSynchronize(ctx->copy_stream));
                return RC_PACK_INVALID;
            }
            t = now();
            const double took = (t - p0) * (double)chunk / (double)cn;
            pack_s = 0.5 * (pack_s + took);
            bytes = (size_t)3 * words * 8;
            FQD_CUDA(cudaMemcpyAsync(dev + c0 * L, stage, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
            FQD_CUDA(cudaEventRecord(ctx->pack_ev[slot], ctx->copy_stream));
            cp.keys = dev + c0 * L;
            cp.plane_words = words;
        }
        if (i == 0) { first_copy_t0 = t; first_copy_bytes = bytes; }
        link_free_at = std::max(link_free_at, t) + (double)bytes / link_bps;
        ctx->h2d_bytes += bytes;
        FQD_CUDA(cudaEventRecord(ctx->chunk_events[i], ctx->copy_stream));
        FQD_CUDA(cudaStreamWaitEvent(s, ctx->chunk_events[i], 0));
        last_copy = i;
        if (raw) launch_ascii(cp);
        else partition_planes_kernel<PW, NW><<<cdiv(cn, 256 * LEAN_ROWS), 256, 0, s>>>(cp);
        tt.launches++;
    }
    FQD_CUDA(cudaGetLastError());
    FQD_CUDA(cudaEventRecord(ctx->ev[11], ctx->copy_stream));
    tt.streamed = true;
    return FQD_OK;
}

// The partition pass of the streaming plan over the job's records: filter + pack + hash, every record appended to
// the tile of its hash (of pigeonhole block 0 of `part_blocks` when > 0, of the whole key otherwise).
template <int K, int PW>
int launch_partition(fqd_context *ctx, const DeviceJob &job, const Codec &codec, const IngestParams &ip, const PartParams &part,
                     uint32_t part_blocks, uint32_t index_base, StageTimes &tt)
{
    cudaStream_t s = ctx->stream;
    uint32_t stride = 0;
    if (!job.key_off) stride = job.key_stride;
    if (job.filter_on && !job.qual_off) stride = std::max(stride, job.qual_stride);
    const bool fixed_any = !job.key_off || (job.filter_on && !job.qual_off);
    constexpr uint32_t BR = 256u * INGEST_ROWS;   // records per block
    IngestParams pp = ip;
    pp.part = part;
    pp.phase = 0;
    pp.part_blocks = part_blocks;
    pp.codec.swar = codec.swar || (K == 3 && codec_is_dna(codec) && !getenv("FQD_NO_SWAR"));
    size_t smem = 1280;
    if (fixed_any && (size_t)stride * BR + 1280 <= 200 * 1024) {
        pp.stage_bytes = stride * BR;
        smem += pp.stage_bytes;
    }
    FQD_CUDA(cudaFuncSetAttribute(ingest_kernel<K, PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // the common large job has its own lean partition kernel (partitioned.cuh)
    int lean_nw = 0;
    if constexpr (K == 3 && slot_words(K * PW) == PART_RW) {
        if (codec_is_dna(codec) && !job.filter_on && !job.key_off && !job.key_lens && !job.weights && !job.varlen &&
            job.key_stride == job.key_len && job.key_len == job.max_len && (job.key_len & 3u) == 0 &&
            !getenv("FQD_NO_SWAR") && !getenv("FQD_NO_LEAN"))
            lean_nw = (int)(job.key_len >> 2);
    }
    // the lean ASCII partition kernel over one chunk of rows (cp.keys / cp.n / cp.index_base set by the caller)
    // (false: no instance for this key length -- the general ingest kernel takes the chunk)
    auto launch_lean = [&](const IngestParams &cp) -> bool {
        if constexpr (K == 3 && slot_words(K * PW) == PART_RW) {
            if (lean_nw == 3) { partition_dna_kernel<PW, 3><<<cdiv(cp.n, 256 * LEAN_ROWS), 256, 0, s>>>(cp); return true; }
            if (lean_nw == 6) { partition_dna_kernel<PW, 6><<<cdiv(cp.n, 256 * LEAN_ROWS), 256, 0, s>>>(cp); return true; }
            if constexpr (PW >= 2) {
                if (lean_nw == 9) { partition_dna_kernel<PW, 9><<<cdiv(cp.n, 256 * LEAN_ROWS), 256, 0, s>>>(cp); return true; }
                if (lean_nw == 12) { partition_dna_kernel<PW, 12><<<cdiv(cp.n, 256 * LEAN_ROWS), 256, 0, s>>>(cp); return true; }
            }
        }
        return false;
    };
    if constexpr (K == 3 && slot_words(K * PW) == PART_RW) {
        if (lean_nw && job.host_keys && job.host_pack && job.n) {
            int rc = RC_PACK_INVALID;
            if (lean_nw == 3) rc = launch_packed_chunks<PW, 3>(ctx, job, pp, index_base, tt, launch_lean);
            else if (lean_nw == 6) rc = launch_packed_chunks<PW, 6>(ctx, job, pp, index_base, tt, launch_lean);
            if constexpr (PW >= 2) {
                if (lean_nw == 9) rc = launch_packed_chunks<PW, 9>(ctx, job, pp, index_base, tt, launch_lean);
                else if (lean_nw == 12) rc = launch_packed_chunks<PW, 12>(ctx, job, pp, index_base, tt, launch_lean);
            }
            if (rc == FQD_OK) return FQD_OK;
            if (rc != RC_PACK_INVALID) return rc;
            // forget what the packed chunks appended, then the ASCII path (it names the unknown bytes)
            FQD_CUDA(cudaMemsetAsync(part.cursor, 0, (size_t)part.nparts * 4, s));
            if (part.spill_cnt) FQD_CUDA(cudaMemsetAsync(part.spill_cnt, 0, 8, s));
        }
    }
    FQD_TRY(for_each_input_chunk(ctx, job, pp, index_base, nullptr, tt, [&](const IngestParams &cp) {
        if (lean_nw && launch_lean(cp)) return;
        ingest_kernel<K, PW><<<cdiv(cp.n, BR), 256, smem, s>>>(cp);
    }));
    FQD_CUDA(cudaGetLastError());
    return FQD_OK;
}

template <int K, int PW>
int stage_dedupe(fqd_context *ctx, const DeviceJob &job, const Codec &codec, uint32_t index_base,
                 bool sharded, fqd_cluster_stats *st, uint32_t unknown_out[8], Uniques &uq,
                 StageTimes &tt, FusedPass0 *fp = nullptr)
{
    constexpr int KW = K * PW, RW = slot_words(KW);
    cudaStream_t s = ctx->stream;
    const uint64_t n = job.n;
    cudaEvent_t *ev = ctx->ev;

    IngestParams ip{};
    ip.n = n;
    ip.keys = job.keys; ip.key_off = job.key_off; ip.key_lens = job.key_lens;
    ip.key_stride = job.key_stride; ip.key_len = job.key_len;
    ip.quals = job.quals; ip.qual_off = job.qual_off; ip.qual_lens = job.qual_lens;
    ip.qual_stride = job.qual_stride; ip.qual_len = job.qual_len;
    ip.max_len = job.max_len;
    ip.filter_on = job.filter_on ? 1 : 0;
    ip.max_err = job.max_err;
    ip.phred_offset = job.phred_offset;
    ip.pad_code = codec.pad_code;
    ip.weights = job.weights;
    ip.index_base = index_base;
    ip.sharded = sharded ? 1 : 0;
    ip.ctr = ctx->d_ctr;
    ip.codec = codec;
    uint32_t stride = 0;
    if (!job.key_off) stride = job.key_stride;
    if (job.filter_on && !job.qual_off) stride = std::max(stride, job.qual_stride);
    const bool fixed_any = !job.key_off || (job.filter_on && !job.qual_off);

    auto check_input_errors = [&](const DevCounters &c, bool &retry) -> int {
        retry = false;
        bool any_unknown = false;
        for (int i = 0; i < 8; i++) { unknown_out[i] = c.unknown[i]; any_unknown |= c.unknown[i] != 0; }
        st->bad_record = ~0ull;
        if (c.phred_err != ~0ull) {
            st->bad_record = c.phred_err >> 8;
            st->bad_char = (uint32_t)(c.phred_err & 0xFF);
            if (!sharded) {
                set_error("Character %c outside of valid phred range ('%c' to '%c')",
                          (int)st->bad_char, (int)job.phred_offset, 126);
                return FQD_ERR_PHRED;
            }
        }
        if (any_unknown && !sharded) retry = true;
        return FQD_OK;
    };
    auto finish = [&](const DevCounters &c, uint32_t U) {
        st->total_records = n;
        st->discarded_records = c.n_discarded;
        st->number_of_sequences = job.weights ? c.sum_weights : n - c.n_discarded;
        uq.U = U;
    };

    // ================= streaming plan (partitioned.cuh) =================
    const size_t plan_mark = arena_mark(ctx);
    if constexpr (RW == PART_RW) {
        uint64_t part_min = PARTITION_MIN_RECORDS;
        if (const char *e = getenv("FQD_PARTITION_MIN")) part_min = strtoull(e, nullptr, 10);   // tests
        // attempt 0: records partitioned by pigeonhole block 0, pass 0 fused into the dedupe tiles;
        // attempt 1 (or the only one): partitioned by the whole key
        const int first_attempt = (fp && fp->want && !getenv("FQD_NO_FUSED_PASS0")) ? 0 : 1;
        for (int attempt = first_attempt; attempt < 2 && n >= part_min && n > 0 && !getenv("FQD_NO_PARTITION"); attempt++) {
            const bool fused = attempt == 0;
            FQD_TRY(reset_counters(ctx));
            FQD_CUDA(cudaEventRecord(ev[0], s));
            const uint32_t nparts = tile_partitions(n);
            const uint32_t spill_cap = (uint32_t)(n / 4 + 4096);
            uint32_t *buf, *cursor, *spill, *aux, *oversize;
            FQD_TRY(arena(ctx, (size_t)nparts * TILE_R * RW, &buf));
            FQD_TRY(arena(ctx, nparts, &cursor));
            FQD_TRY(arena(ctx, (size_t)spill_cap * RW, &spill));
            FQD_TRY(arena(ctx, 8, &aux));   // [0] spill count, [1] spill overflow, [2] dense uniques, [3] oversize partitions
            FQD_TRY(arena(ctx, nparts, &oversize));
            FQD_TRY(arena(ctx, n * KW, &uq.ukey));      // worst case: every record distinct
            FQD_TRY(arena(ctx, n, &uq.ucount));
            FQD_TRY(arena(ctx, n, &uq.ufirst));
            FQD_CUDA(cudaMemsetAsync(cursor, 0, (size_t)nparts * 4, s));
            FQD_CUDA(cudaMemsetAsync(aux, 0, 32, s));
            FQD_CUDA(cudaEventRecord(ev[1], s));
            const PartParams part{buf, cursor, nparts, spill, aux, spill_cap};
            FQD_TRY((launch_partition<K, PW>(ctx, job, codec, ip, part, fused ? (uint32_t)job.d + 1u : 0u, index_base, tt)));
            FQD_CUDA(cudaGetLastError());
            FQD_CUDA(cudaEventRecord(ev[2], s));
            DedupeOut out{uq.ukey, uq.ucount, uq.ufirst, aux + 2, oversize, aux + 3, sharded ? 1 : 0, (uint32_t)n, aux + 5};
            PassParams p0{};
            EdgeSink sink0{};
            if (fused) {
                p0.d = job.d; p0.edit = 0; p0.varlen = job.varlen ? 1 : 0; p0.method = job.method;
                p0.max_len = job.max_len; p0.pad_code = codec.pad_code; p0.V = 1; p0.world = 1;
                p0.pass_j = 0;
                p0.fix_st = 0;
                p0.fix_bl = block_start(job.max_len, 1u, (uint32_t)job.d + 1u);
                p0.dominated = fp->dominated; p0.dead = fp->dead; p0.ctr = ctx->d_ctr;
                sink0 = EdgeSink{fp->edges, fp->aux, fp->edge_cap, fp->aux + 1};
                NextPass nx{};
                if (fp->next_buf) {
                    FQD_CUDA(cudaMemsetAsync(fp->next_cursor, 0, (size_t)fp->next_nparts * 4, s));
                    nx.next = PartParams{fp->next_buf, fp->next_cursor, fp->next_nparts, nullptr, nullptr, 0};
                    nx.pass_j = 1;
                    nx.st = block_start(job.max_len, 1u, (uint32_t)job.d + 1u);
                    nx.bl = block_start(job.max_len, 2u, (uint32_t)job.d + 1u) - nx.st;
                }
                if (tile_split_enabled(nparts)) {
                    TileSplit ts;
                    FQD_TRY(tile_split_prepare(ctx, part, ts, tt));
                    dedupe_tile_kernel<K, PW, true, TILE_R_SMALL, TILE_SMALL_BLOCKS><<<nparts, TILE_THREADS, 0, s>>>(
                        small_tiles(single_source(part)), out, p0, sink0, nx);
                    dedupe_tile_kernel<K, PW, true><<<ctx->sm_count * 10, TILE_THREADS, 0, s>>>(
                        large_tiles(single_source(part), ts), out, p0, sink0, nx);
                    tt.launches++;
                } else {
                    dedupe_tile_kernel<K, PW, true><<<nparts, TILE_THREADS, 0, s>>>(single_source(part), out, p0, sink0, nx);
                }
            } else {
                if (tile_split_enabled(nparts)) {
                    TileSplit ts;
                    FQD_TRY(tile_split_prepare(ctx, part, ts, tt));
                    dedupe_tile_kernel<K, PW, false, TILE_R_SMALL, TILE_SMALL_BLOCKS><<<nparts, TILE_THREADS, 0, s>>>(
                        small_tiles(single_source(part)), out, p0, sink0, NextPass{});
                    dedupe_tile_kernel<K, PW, false><<<ctx->sm_count * 10, TILE_THREADS, 0, s>>>(
                        large_tiles(single_source(part), ts), out, p0, sink0, NextPass{});
                    tt.launches++;
                } else {
                    dedupe_tile_kernel<K, PW, false><<<nparts, TILE_THREADS, 0, s>>>(single_source(part), out, p0, sink0, NextPass{});
                }
            }
            tt.launches++;
            FQD_CUDA(cudaGetLastError());
            FQD_CUDA(cudaEventRecord(ev[3], s));
            uint32_t h_aux[4] = {};
            FQD_CUDA(cudaMemcpyAsync(h_aux, aux, sizeof h_aux, cudaMemcpyDeviceToHost, s));
            FQD_TRY(fetch_counters(ctx));
            const DevCounters c1 = *ctx->h_ctr;
            bool retry = false;
            FQD_TRY(check_input_errors(c1, retry));
            if (retry) return RC_RETRY_ALPHABET;
            if (!h_aux[1]) {
                uint32_t U = h_aux[2];
                if (h_aux[3]) {
                    // oversize partitions + their spilled records: single-table dedupe, appended to the dense arrays
                    const uint32_t n_over = h_aux[3], n_spill = h_aux[0];
                    const uint64_t n_rec = (uint64_t)n_over * TILE_R + n_spill;
                    const uint64_t capacity = n_rec + (n_rec >> 1) + 1024;
                    uint32_t *table, *uslot;
                    FQD_TRY(arena(ctx, capacity * RW, &table));
                    FQD_TRY(arena(ctx, n_rec, &uslot));
                    FQD_CUDA(cudaMemsetAsync(table, 0xFF, capacity * RW * 4, s));
                    FQD_CUDA(cudaMemsetAsync(aux + 4, 0, 4, s));
                    const TableRef tr{table, capacity, uslot, ctx->d_ctr};
                    spill_insert_kernel<K, PW><<<n_over * (TILE_R / 256), 256, 0, s>>>(buf, oversize, TILE_R, tr, aux + 4);
                    if (n_spill) spill_insert_kernel<K, PW><<<cdiv(n_spill, 256), 256, 0, s>>>(spill, nullptr, n_spill, tr, aux + 4);
                    uint32_t n_claimed = 0;
                    FQD_CUDA(cudaMemcpyAsync(&n_claimed, aux + 4, 4, cudaMemcpyDeviceToHost, s));
                    FQD_CUDA(cudaStreamSynchronize(s));
                    if (n_claimed)
                        gather_nonzero_kernel<K, PW><<<cdiv(n_claimed, 256), 256, 0, s>>>(n_claimed, table, uslot, uq.ukey, uq.ucount,
                                                                                           uq.ufirst, aux + 2, sharded ? 1 : 0);
                    FQD_CUDA(cudaGetLastError());
                    FQD_CUDA(cudaEventRecord(ev[3], s));
                    FQD_CUDA(cudaMemcpyAsync(&U, aux + 2, 4, cudaMemcpyDeviceToHost, s));
                    FQD_TRY(fetch_counters(ctx));
                    if (ctx->h_ctr->table_full) { set_error("internal: spill table overflow"); return FQD_ERR_NOMEM; }
                    tt.launches += 3;
                }
                finish(c1, U);
                if (fused) {
                    // complete only if no tile buffer overflowed and no partition took the spill path
                    uint32_t h_f[2] = {};
                    FQD_CUDA(cudaMemcpyAsync(h_f, fp->aux, sizeof h_f, cudaMemcpyDeviceToHost, s));
                    FQD_CUDA(cudaStreamSynchronize(s));
                    // the uniques of oversize partitions (ids [h_aux[2], U)) were not compared inside a tile:
                    // a few of them are finished by brute force, many mean pass 0 is redone the ordinary way
                    fp->spill_lo = h_aux[2];
                    fp->spill_hi = U;
                    fp->done = !h_f[1];
                    fp->next_valid = fp->done && fp->next_buf != nullptr;
                    if (getenv("FQD_TRACE"))
                        fprintf(stderr, "[fqd trace] fused pass 0: edges %u overflow %u oversize partitions %u spill %u (%u uniques) -> %s\n",
                                h_f[0], h_f[1], h_aux[3], h_aux[0], U - h_aux[2], fp->done ? "done" : "redo");
                }
                cudaEventElapsedTime(&tt.table_clear, ev[0], ev[1]);
                cudaEventElapsedTime(&tt.ingest_kernel, ev[1], ev[2]);
                cudaEventElapsedTime(&tt.dedupe_kernel, ev[2], ev[3]);
                cudaEventElapsedTime(&tt.ingest, ev[0], ev[3]);
                if (tt.streamed) {
                    FQD_CUDA(cudaEventSynchronize(ctx->ev[11]));
                    cudaEventElapsedTime(&tt.h2d, ctx->ev[10], ctx->ev[11]);
                }
                st->ms_h2d = tt.h2d;
                tt.partitioned = true;
                return FQD_OK;
            }
            // the spill buffer overflowed (a few keys dominate the input): single-table plan instead
            arena_release(ctx, plan_mark);
            tt.launches = 0;
            if (fused) {   // the tiles numbered the uniques differently: forget what they flagged
                FQD_CUDA(cudaMemsetAsync(fp->aux, 0, 16, s));
                if (fp->dominated) {
                    FQD_CUDA(cudaMemsetAsync(fp->dominated, 0, n, s));
                    FQD_CUDA(cudaMemsetAsync(fp->dead, 0, n, s));
                }
            }
        }
    }

    // ================= single-table plan =================
    FQD_TRY(reset_counters(ctx));
    FQD_CUDA(cudaEventRecord(ev[0], s));
    const uint64_t capacity = std::max<uint64_t>(1024, n + (n >> 1) + 64);
    if (capacity >= 0xFFFFFFF0ull) {
        set_error("too many records for one job on one GPU (%llu)", (unsigned long long)n);
        return FQD_ERR_UNSUPPORTED;
    }
    uint32_t *table, *uslot, *keepmask;
    FQD_TRY(arena(ctx, capacity * RW, &table));
    FQD_TRY(arena(ctx, std::max<uint64_t>(n, 1), &uslot));
    FQD_TRY(arena(ctx, cdiv(std::max<uint64_t>(n, 1), 32), &keepmask));
    FQD_CUDA(cudaMemsetAsync(table, 0xFF, capacity * RW * sizeof(uint32_t), s));
    FQD_CUDA(cudaEventRecord(ev[1], s));
    ip.tab.table = table; ip.tab.capacity = capacity; ip.tab.uslot = uslot; ip.tab.ctr = ctx->d_ctr;
    ip.keepmask = keepmask;
    constexpr uint32_t BR = 256u * INGEST_ROWS;   // records per block
    size_t smem = 1280;
    if (fixed_any && (size_t)stride * BR + 1280 <= 200 * 1024) {
        ip.stage_bytes = stride * BR;
        smem += ip.stage_bytes;
    }
    FQD_CUDA(cudaFuncSetAttribute(ingest_kernel<K, PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ip.phase = 0;
    FQD_TRY(for_each_input_chunk(ctx, job, ip, index_base, keepmask, tt, [&](const IngestParams &cp) {
        ingest_kernel<K, PW><<<cdiv(cp.n, BR), 256, smem, s>>>(cp);
    }));
    FQD_CUDA(cudaEventRecord(ev[2], s));
    FQD_CUDA(cudaGetLastError());
    FQD_TRY(fetch_counters(ctx));
    const DevCounters c1 = *ctx->h_ctr;
    bool retry = false;
    FQD_TRY(check_input_errors(c1, retry));
    if (retry) return RC_RETRY_ALPHABET;
    if (c1.table_full) { set_error("internal: dedupe table overflow"); return FQD_ERR_NOMEM; }
    if (job.filter_on && c1.n_discarded && !sharded) {
        ip.phase = 1;
        ingest_kernel<K, PW><<<cdiv(n, BR), 256, smem, s>>>(ip);
        tt.launches++;
        FQD_CUDA(cudaGetLastError());
    }
    const uint32_t U = c1.n_unique;
    finish(c1, U);
    FQD_TRY(arena(ctx, (size_t)U * KW, &uq.ukey));
    FQD_TRY(arena(ctx, U, &uq.ucount));
    FQD_TRY(arena(ctx, U, &uq.ufirst));
    if (U) {
        gather_kernel<K, PW><<<cdiv(U, 256), 256, 0, s>>>(U, table, uslot, uq.ukey, uq.ucount, uq.ufirst,
                                                          nullptr, nullptr, nullptr);
        tt.launches++;
    }
    FQD_CUDA(cudaGetLastError());
    FQD_CUDA(cudaEventRecord(ev[3], s));
    FQD_CUDA(cudaEventSynchronize(ev[3]));
    cudaEventElapsedTime(&tt.table_clear, ev[0], ev[1]);
    cudaEventElapsedTime(&tt.ingest_kernel, ev[1], ev[2]);
    cudaEventElapsedTime(&tt.ingest, ev[0], ev[3]);
    if (tt.streamed) {
        FQD_CUDA(cudaEventSynchronize(ctx->ev[11]));
        cudaEventElapsedTime(&tt.h2d, ctx->ev[10], ctx->ev[11]);
    }
    st->ms_h2d = tt.h2d;
    return FQD_OK;
}

// ---- stage 2: forest / flag arrays over U uniques ----------------------------------------------

int stage_forest_alloc(fqd_context *ctx, int method, uint32_t U, Forest &f, const FusedPass0 *fp = nullptr)
{
    cudaStream_t s = ctx->stream;
    FQD_TRY(arena(ctx, U, &f.parent_full));
    FQD_TRY(arena(ctx, U, &f.root));
    FQD_TRY(arena(ctx, U, &f.selected));
    if (method == METHOD_DIRECTIONAL) {
        FQD_TRY(arena(ctx, U, &f.parent_one));
        FQD_TRY(arena(ctx, U, &f.deadroot));
        if (fp && fp->dominated) {   // allocated and zeroed before the dedupe stage, maybe already written by its tiles
            f.dominated = fp->dominated;
            f.dead = fp->dead;
        } else {
            FQD_TRY(arena(ctx, U, &f.dominated));
            FQD_TRY(arena(ctx, U, &f.dead));
            FQD_CUDA(cudaMemsetAsync(f.dominated, 0, std::max<size_t>(U, 1), s));
            FQD_CUDA(cudaMemsetAsync(f.dead, 0, std::max<size_t>(U, 1), s));
        }
        FQD_CUDA(cudaMemsetAsync(f.deadroot, 0, std::max<size_t>(U, 1), s));
    }
    if (method != METHOD_ADJACENCY) FQD_TRY(arena(ctx, U, &f.best));
    if (method == METHOD_ADJACENCY) {
        f.edge_cap = 2ull * U + (1ull << 16);
        FQD_TRY(arena(ctx, f.edge_cap, &f.edges));
    }
    return FQD_OK;
}

// ---- stage 3: pigeonhole passes over the buckets this rank owns --------------------------------

template <int K, int PW>
int stage_passes(fqd_context *ctx, const DeviceJob &job, const Codec &codec, const Uniques &uq,
                 Forest &f, int rank, int world, fqd_cluster_stats *st, StageTimes &tt, int first_pass = 0,
                 int end_pass = -1, uint32_t u_lo = 0, const FusedPass0 *pre = nullptr)
{
    constexpr int KW = K * PW, FW = fat_words(KW);
    cudaStream_t s = ctx->stream;
    const uint32_t U = uq.U;
    // passes [first_pass, npass) over the uniques [u_lo, U) (a sub-range only for the streaming plan:
    // the uniques that left the dedupe stage through the spill path redo pass 0 among themselves)
    const int npass_all = (job.d > 0 && U > 1) ? job.d + 1 : 0;
    const int npass = end_pass < 0 ? npass_all : std::min(end_pass, npass_all);
    if (end_pass < 0) st->n_passes = npass_all;
    if (npass <= first_pass) return FQD_OK;
    const int V = job.edit ? (job.varlen ? 2 * job.d + 1 : 1) * (job.d + 1) : 1;
    const uint64_t E = (uint64_t)U * V;
    if (E >= 0xFFFFFFF0ull) { set_error("too many pigeonhole entries (%llu)", (unsigned long long)E); return FQD_ERR_UNSUPPORTED; }
    const bool fat = !job.edit;
    PassParams pp{};
    pp.U = U; pp.u_lo = u_lo; pp.ukey = uq.ukey; pp.ucount = uq.ucount;
    pp.d = job.d; pp.edit = job.edit; pp.varlen = job.varlen ? 1 : 0; pp.method = job.method;
    pp.max_len = job.max_len; pp.pad_code = codec.pad_code;
    pp.V = V; pp.my_rank = rank; pp.world = world;
    pp.parent_full = f.parent_full; pp.parent_one = f.parent_one;
    pp.dominated = f.dominated; pp.dead = f.dead;
    pp.edges = f.edges; pp.edge_cap = f.edge_cap; pp.ctr = ctx->d_ctr;
    for (int i = 0; i < 256; i++) pp.rank_of_code[i] = codec.rank[i];
    EventSet cev;   // per pass: before / after the compare
    FQD_CUDA(cev.create((size_t)2 * npass));

    // Hamming passes of large jobs: partition by the block hash, multimap in L2 (partitioned.cuh)
    bool use_part = false;
    if constexpr (FW == PART_RW) {
        uint64_t part_min = PARTITION_MIN_UNIQUES;
        if (const char *e = getenv("FQD_PARTITION_MIN")) part_min = strtoull(e, nullptr, 10);   // tests
        use_part = fat && (U >= part_min || u_lo) && !getenv("FQD_NO_PARTITION") && !getenv("FQD_NO_PARTITION_PASSES");
    }
    const uint32_t own_avg = (U - u_lo) / (uint32_t)std::max(world, 1);   // entries this rank's buckets receive
    const uint32_t nparts = tile_partitions(own_avg);
    uint32_t *pbuf = nullptr, *pcursor = nullptr, *paux = nullptr;   // paux per pass: [0] edge count, [1] overflow flag
    EdgeSink sink{};
    if (use_part) {
        FQD_TRY(arena(ctx, (size_t)nparts * TILE_R * FW, &pbuf));
        FQD_TRY(arena(ctx, (size_t)npass * nparts, &pcursor));
        FQD_TRY(arena(ctx, (size_t)npass * 4, &paux));
        sink.cap = (uint32_t)std::min<uint64_t>(0xFFFFFFF0ull, (uint64_t)U + (1u << 16));
        FQD_TRY(arena(ctx, (size_t)sink.cap, &sink.edges));
    }

    // legacy plan (small jobs, Levenshtein, skewed buckets): counting sort by block hash + compare
    uint32_t NB = 1024;
    while (NB < (1u << 24) && NB < E / 2) NB <<= 1;
    uint32_t *cnt = nullptr, *fill = nullptr, *rank_arr = nullptr, *entries = nullptr, *block_sums = nullptr, *grand = nullptr;
    auto legacy_alloc = [&]() -> int {
        if (cnt) return FQD_OK;
        FQD_TRY(arena(ctx, (size_t)NB + 1, &cnt));
        FQD_TRY(arena(ctx, E, &rank_arr));
        if (!fat) FQD_TRY(arena(ctx, (size_t)2 * NB, &fill));
        FQD_TRY(arena(ctx, fat ? E * FW : E * 2, &entries));
        FQD_TRY(arena(ctx, (size_t)cdiv(NB, SCAN_TILE) + 16, &block_sums));
        FQD_TRY(arena(ctx, 4, &grand));
        pp.nb_mask = NB - 1;
        pp.cnt = cnt; pp.fill = fill; pp.rank = rank_arr; pp.entries = reinterpret_cast<uint2 *>(entries); pp.fat = entries;
        return FQD_OK;
    };
    auto legacy_pass = [&](int j) -> int {
        FQD_TRY(legacy_alloc());
        pp.pass_j = j;
        pp.fix_st = block_start(job.max_len, (uint32_t)j, (uint32_t)job.d + 1u);
        pp.fix_bl = block_start(job.max_len, (uint32_t)j + 1u, (uint32_t)job.d + 1u) - pp.fix_st;
        FQD_CUDA(cudaMemsetAsync(cnt, 0, ((size_t)NB + 1) * 4, s));
        sig_count_kernel<K, PW><<<cdiv(U, 256), 256, 0, s>>>(pp);
        FQD_TRY(exclusive_scan_inplace(ctx, cnt, NB, block_sums, grand));
        if (fat) scatter_fat_kernel<K, PW><<<cdiv(U, 256), 256, 0, s>>>(pp);
        else {
            FQD_CUDA(cudaMemsetAsync(fill, 0, (size_t)2 * NB * 4, s));
            scatter_kernel<K, PW><<<cdiv(U, 256), 256, 0, s>>>(pp);
        }
        pp.n_entries = (uint32_t)E;   // upper bound; the kernel stops at cnt[NB]
        tt.launches += 6;             // sig_count, 3 scan kernels, scatter, compare
        FQD_CUDA(cudaEventRecord(cev[2 * j], s));
        if (fat) compare_fat_kernel<K, PW><<<cdiv(E, 256), 256, 0, s>>>(pp);
        else {
            // very big buckets (short blocks: hundreds of keys per block value) take the dense tiles, the rest the
            // per-entry walk.  Measured on config 4 (1.6 M keys of 24 nt), builds first in every bucket: d = 2 / 8-nt
            // blocks, ~73 entries per bucket: walk 6.3 ms, dense 6.6 ms of compare; d = 1 / 12-nt blocks: 0.27 vs 0.46 ms
            const double block_values = std::pow(4.0, std::min<double>(pp.fix_bl, 20.0));
            bool dense = (double)E > 256.0 * block_values;
            if (const char *e = getenv("FQD_COMPARE_DENSE")) dense = atoi(e) != 0;   // measurement switch
            if (dense) compare_dense_kernel<K, PW><<<cdiv(E, 32u * dense_warps<KW>()), 32 * dense_warps<KW>(), 0, s>>>(pp);
            else compare_kernel<K, PW><<<cdiv(E, 256), 256, 0, s>>>(pp);
        }
        FQD_CUDA(cudaEventRecord(cev[2 * j + 1], s));
        FQD_CUDA(cudaGetLastError());
        return FQD_OK;
    };

    // one Hamming pass in tiles: partition (unless the dedupe tiles already did it), tile kernel, hooks
    auto tile_pass = [&](int j, bool prepart) -> int {
        if constexpr (FW == PART_RW) {
            pp.pass_j = j;
            pp.fix_st = block_start(job.max_len, (uint32_t)j, (uint32_t)job.d + 1u);
            pp.fix_bl = block_start(job.max_len, (uint32_t)j + 1u, (uint32_t)job.d + 1u) - pp.fix_st;
            PartParams qp{pbuf, pcursor + (size_t)j * nparts, nparts, nullptr, nullptr, 0};
            sink.n_edges = paux + 4 * j;
            sink.overflow = paux + 4 * j + 1;
            if (prepart) {
                // (only the uniques that left the dedupe stage through the spill path are missing)
                qp = PartParams{pre->next_buf, pre->next_cursor, pre->next_nparts, nullptr, nullptr, 0};
                if (pre->spill_hi > pre->spill_lo) {
                    PassParams sp = pp;
                    sp.u_lo = pre->spill_lo;
                    bucket_partition_kernel<K, PW><<<cdiv(pre->spill_hi - pre->spill_lo, 256 * BP_ROWS), 256, 0, s>>>(sp, qp);
                    tt.launches++;
                }
            } else {
                bucket_partition_kernel<K, PW><<<cdiv(U - u_lo, 256 * BP_ROWS), 256, 0, s>>>(pp, qp);
                tt.launches++;
            }
            FQD_CUDA(cudaEventRecord(cev[2 * j], s));
            if (tile_split_enabled(qp.nparts)) {
                TileSplit ts;
                FQD_TRY(tile_split_prepare(ctx, qp, ts, tt));
                bucket_tile_kernel<K, PW, TILE_R_SMALL, TILE_SMALL_BLOCKS><<<qp.nparts, TILE_THREADS, 0, s>>>(
                    small_tiles(single_source(qp)), pp, sink);
                bucket_tile_kernel<K, PW><<<ctx->sm_count * 10, TILE_THREADS, 0, s>>>(large_tiles(single_source(qp), ts), pp, sink);
                tt.launches++;
            } else {
                bucket_tile_kernel<K, PW><<<qp.nparts, TILE_THREADS, 0, s>>>(single_source(qp), pp, sink);
            }
            apply_edges_kernel<<<ctx->sm_count * 8, 256, 0, s>>>(single_edges(sink.edges, sink.n_edges, sink.cap), f.parent_full,
                                                                f.parent_one, EdgeFlags{}, ctx->d_ctr);
            FQD_CUDA(cudaEventRecord(cev[2 * j + 1], s));
            FQD_CUDA(cudaGetLastError());
            tt.launches += 2;
        }
        return FQD_OK;
    };
    auto read_aux = [&](std::vector<uint32_t> &h_aux) -> int {
        h_aux.assign((size_t)npass * 4, 0);
        FQD_CUDA(cudaMemcpyAsync(h_aux.data(), paux, h_aux.size() * 4, cudaMemcpyDeviceToHost, s));
        FQD_CUDA(cudaStreamSynchronize(s));
        return FQD_OK;
    };

    for (int attempt = 0; attempt < 2; attempt++) {
        if (use_part) {
            FQD_CUDA(cudaMemsetAsync(pcursor, 0, (size_t)npass * nparts * 4, s));
            FQD_CUDA(cudaMemsetAsync(paux, 0, (size_t)npass * 16, s));
            const bool pre1 = pre && pre->next_valid && attempt == 0 && u_lo == 0;   // pass 1 arrives pre-partitioned
            for (int j = first_pass; j < npass; j++) FQD_TRY(tile_pass(j, pre1 && j == 1));
            std::vector<uint32_t> h_aux;
            FQD_TRY(read_aux(h_aux));
            if (pre1 && first_pass <= 1 && npass > 1 && h_aux[4 * 1 + 1]) {
                // more uniques than the dedupe stage guessed overflowed the pre-partitioned tiles: pass 1 again,
                // partitioned the ordinary way (the edges it already found are true edges: harmless)
                if (getenv("FQD_TRACE")) fprintf(stderr, "[fqd trace] pass 1: pre-partitioned tiles overflowed, partitioning again\n");
                FQD_CUDA(cudaMemsetAsync(paux + 4, 0, 16, s));
                FQD_TRY(tile_pass(1, false));
                FQD_TRY(read_aux(h_aux));
            }
            else if (pre1 && first_pass <= 1 && npass > 1) tt.pass1_emitted = true;
            // passes whose partitions outgrew a tile (few distinct block values): counting-sort plan
            tt.passes_partitioned = true;
            for (int j = first_pass; j < npass; j++) {
                if (getenv("FQD_TRACE"))
                    fprintf(stderr, "[fqd trace] pass %d over uniques [%u, %u): %u edges%s%s\n", j, u_lo, U, h_aux[4 * j],
                            pre1 && j == 1 ? " (tiles filled by the dedupe stage)" : "",
                            h_aux[4 * j + 1] ? ", a tile overflowed -> counting-sort plan" : "");
                if (h_aux[4 * j + 1]) { tt.passes_partitioned = false; FQD_TRY(legacy_pass(j)); }
            }
        } else {
            for (int j = first_pass; j < npass; j++) FQD_TRY(legacy_pass(j));
        }
        if (job.method != METHOD_ADJACENCY) break;
        FQD_TRY(fetch_counters(ctx));
        if (ctx->h_ctr->n_edges <= f.edge_cap) break;
        if (attempt == 1) { set_error("internal: adjacency edge list overflow"); return FQD_ERR_NOMEM; }
        // edge list overflowed: size it exactly, reset the forest and redo the passes
        f.edge_cap = ctx->h_ctr->n_edges + 16;
        FQD_TRY(arena(ctx, f.edge_cap, &f.edges));
        pp.edges = f.edges; pp.edge_cap = f.edge_cap;
        ctx->h_ctr->n_edges = 0; ctx->h_ctr->n_merges = 0; ctx->h_ctr->n_candidates = 0;
        FQD_CUDA(cudaMemcpyAsync(ctx->d_ctr, ctx->h_ctr, sizeof(DevCounters), cudaMemcpyHostToDevice, s));
        iota_kernel<<<cdiv(U, 256), 256, 0, s>>>(f.parent_full, U);
        tt.launches++;
    }
    FQD_CUDA(cudaStreamSynchronize(s));
    for (int j = first_pass; j < npass; j++) {
        float t = 0.f;
        cudaEventElapsedTime(&t, cev[2 * j], cev[2 * j + 1]);
        tt.compare += t;
    }
    return FQD_OK;
}

// ---- stage 4: dissection + output ------------------------------------------------------------------

template <int K, int PW>
int stage_select(fqd_context *ctx, const DeviceJob &job, const Codec &codec, const Uniques &uq,
                 Forest &f, uint32_t bitmap_base, uint32_t bitmap_n, StageTimes &tt, bool own_only = false)
{
    cudaStream_t s = ctx->stream;
    const uint32_t U = uq.U;
    SelectParams sp{};
    sp.U = U; sp.ukey = uq.ukey; sp.ucount = uq.ucount; sp.ufirst = uq.ufirst;
    sp.parent_full = f.parent_full; sp.parent_one = f.parent_one;
    sp.best = f.best; sp.root = f.root;
    sp.dominated = f.dominated; sp.dead = f.dead; sp.deadroot = f.deadroot;
    sp.selected = f.selected;
    sp.method = job.method; sp.ctr = ctx->d_ctr;
    sp.bitmap = job.bitmap; sp.bitmap_base = bitmap_base; sp.bitmap_n = bitmap_n;
    sp.own_only = own_only ? 1 : 0;
    for (int i = 0; i < 256; i++) sp.rank_of_code[i] = codec.rank[i];
    if (job.bitmap) FQD_CUDA(cudaMemsetAsync(job.bitmap, 0, (size_t)cdiv(std::max<uint32_t>(bitmap_n, 1), 32) * 4, s));
    if (!U) return FQD_OK;
    if (job.method == METHOD_DIRECTIONAL) {
        root_best_kernel<K, PW><<<cdiv(U, 256), 256, 0, s>>>(sp, 1);
        tt.launches++;
    } else if (job.method == METHOD_HIGHEST) {
        root_best_kernel<K, PW><<<cdiv(U, 256), 256, 0, s>>>(sp, 0);
        tt.launches++;
    } else {
        FQD_TRY(arena(ctx, U, &sp.state));
        FQD_TRY(arena(ctx, U, &sp.stamp));
        FQD_CUDA(cudaMemsetAsync(sp.state, 0, U, s));
        FQD_CUDA(cudaMemsetAsync(sp.stamp, 0, (size_t)U * 4, s));
        sp.edges = f.edges;
        sp.n_edges = f.n_edges;
        for (uint32_t round = 1;; round++) {
            sp.round = round;
            FQD_CUDA(cudaMemsetAsync(&ctx->d_ctr->undecided, 0, 4, s));
            if (sp.n_edges) { adj_edge_kernel<<<cdiv(sp.n_edges, 256), 256, 0, s>>>(sp); tt.launches++; }
            adj_node_kernel<<<cdiv(U, 256), 256, 0, s>>>(sp);
            tt.launches++;
            FQD_TRY(fetch_counters(ctx));
            if (ctx->h_ctr->undecided == 0) break;
            if (round > U + 2) { set_error("internal: adjacency rounds did not converge"); return FQD_ERR_CUDA; }
        }
    }
    select_kernel<<<cdiv(U, 256), 256, 0, s>>>(sp);
    tt.launches++;
    FQD_CUDA(cudaGetLastError());
    return FQD_OK;
}

void publish_result(fqd_context *ctx, const Uniques &uq, const Forest &f, uint64_t n_records,
                    uint64_t n_selected)
{
    ctx->res.U = uq.U;
    ctx->res.n_records = n_records;
    ctx->res.n_selected = n_selected;
    ctx->res.ufirst = uq.ufirst;
    ctx->res.ucount = uq.ucount;
    ctx->res.parent_full = f.parent_full;
    ctx->res.selected = f.selected;
}

// ---- single-GPU plan ---------------------------------------------------------------------------------

template <int K, int PW>
int run_typed(fqd_context *ctx, const DeviceJob &job, const Codec &codec, fqd_cluster_stats *st,
              uint32_t unknown_out[8])
{
    cudaStream_t s = ctx->stream;
    ctx->res = fqd_result{};
    st->key_bits = K;
    st->key_words = K * PW;
    cudaEvent_t *ev = ctx->ev;   // 4..8 belong to this plan, 0..3 to stage_dedupe
    FQD_CUDA(cudaEventRecord(ev[4], s));
    StageTimes tt;
    Uniques uq;
    // Hamming jobs: pass 0 can run inside the dedupe tiles (its flag bytes and edge list are
    // needed before the unique count is known, so they are sized by the record count)
    FusedPass0 fp;
    uint64_t part_min = PARTITION_MIN_RECORDS;
    if (const char *e = getenv("FQD_PARTITION_MIN")) part_min = strtoull(e, nullptr, 10);   // tests
    fp.want = !job.edit && job.d >= 1 && job.method != METHOD_ADJACENCY && job.n >= part_min && job.n > 1 &&
              slot_words(K * PW) == PART_RW;
    if (fp.want) {
        fp.edge_cap = (uint32_t)std::min<uint64_t>(0x7FFFFFF0ull, job.n / 2 + (1u << 16));
        FQD_TRY(arena(ctx, (size_t)fp.edge_cap, &fp.edges));
        FQD_TRY(arena(ctx, 4, &fp.aux));
        FQD_CUDA(cudaMemsetAsync(fp.aux, 0, 16, s));
        if (job.d >= 1 && !getenv("FQD_NO_NEXT_EMIT")) {
            // the dedupe tiles also hand every unique to its pass-1 tile; U is unknown, guess n/2
            // (more uniques than that overflow those tiles and pass 1 partitions the ordinary way)
            fp.next_nparts = tile_partitions(std::max<uint64_t>(job.n / 2, 1u << 16));
            FQD_TRY(arena(ctx, (size_t)fp.next_nparts * TILE_R * PART_RW, &fp.next_buf));
            FQD_TRY(arena(ctx, fp.next_nparts, &fp.next_cursor));
        }
        if (job.method == METHOD_DIRECTIONAL) {
            FQD_TRY(arena(ctx, job.n, &fp.dominated));
            FQD_TRY(arena(ctx, job.n, &fp.dead));
            FQD_CUDA(cudaMemsetAsync(fp.dominated, 0, job.n, s));
            FQD_CUDA(cudaMemsetAsync(fp.dead, 0, job.n, s));
        }
    }
    FQD_TRY(stage_dedupe<K, PW>(ctx, job, codec, 0, false, st, unknown_out, uq, tt, &fp));
    const uint32_t U = uq.U;
    st->number_of_uniques = U;
    if (U > EDGE_ID) { set_error("too many unique keys for one GPU (%u)", U); return FQD_ERR_UNSUPPORTED; }   // ids share a word with the edge state
    FQD_CUDA(cudaEventRecord(ev[5], s));
    Forest f;
    FQD_TRY(stage_forest_alloc(ctx, job.method, U, f, &fp));
    if (U) {
        init_forest_kernel<<<cdiv(U, 256), 256, 0, s>>>(U, f.parent_full, f.parent_one, f.best);
        tt.launches++;
    }
    FQD_CUDA(cudaGetLastError());
    FQD_CUDA(cudaEventRecord(ev[6], s));
    int first_pass = 0;
    if (fp.done && U > 1) {
        apply_edges_kernel<<<ctx->sm_count * 8, 256, 0, s>>>(single_edges(fp.edges, fp.aux, fp.edge_cap), f.parent_full, f.parent_one,
                                                            EdgeFlags{}, ctx->d_ctr);
        tt.launches++;
        tt.pass0_fused = true;
        first_pass = 1;
        if (fp.spill_hi > fp.spill_lo + 1 && fp.spill_hi - fp.spill_lo <= FUSED_SPILL_BRUTE) {
            PassParams bp{};
            bp.U = U; bp.ukey = uq.ukey; bp.ucount = uq.ucount;
            bp.d = job.d; bp.varlen = job.varlen ? 1 : 0; bp.method = job.method;
            bp.max_len = job.max_len; bp.pad_code = codec.pad_code;
            bp.parent_full = f.parent_full; bp.parent_one = f.parent_one;
            bp.dominated = f.dominated; bp.dead = f.dead; bp.ctr = ctx->d_ctr;
            for (int i = 0; i < 256; i++) bp.rank_of_code[i] = codec.rank[i];
            const uint32_t nb = cdiv(fp.spill_hi - fp.spill_lo, 256);
            range_pairs_kernel<K, PW><<<dim3(nb, nb), 256, 0, s>>>(bp, fp.spill_lo, fp.spill_hi);
            tt.launches++;
        } else if (fp.spill_hi > fp.spill_lo + 1) {
            // many spilled uniques: pass 0 among themselves the ordinary way (their bucket mates spilled too)
            FQD_TRY(stage_passes<K, PW>(ctx, job, codec, uq, f, 0, 1, st, tt, 0, 1, fp.spill_lo));
        }
    }
    FQD_TRY(stage_passes<K, PW>(ctx, job, codec, uq, f, 0, 1, st, tt, first_pass, -1, 0, &fp));
    if (job.method == METHOD_ADJACENCY) {
        FQD_TRY(fetch_counters(ctx));
        f.n_edges = ctx->h_ctr->n_edges;
    }
    FQD_CUDA(cudaEventRecord(ev[7], s));
    FQD_TRY(stage_select<K, PW>(ctx, job, codec, uq, f, 0, (uint32_t)job.n, tt));
    FQD_TRY(fetch_counters(ctx));
    FQD_CUDA(cudaEventRecord(ev[8], s));
    FQD_CUDA(cudaStreamSynchronize(s));
    const DevCounters c2 = *ctx->h_ctr;
    st->number_of_clusters = (uint64_t)U - c2.n_merges;
    st->number_selected = c2.n_selected;
    st->candidate_pairs = c2.n_candidates;
    cudaEventElapsedTime(&st->ms_total, ev[4], ev[8]);
    cudaEventElapsedTime(&st->ms_ingest, ev[4], ev[5]);
    cudaEventElapsedTime(&st->ms_gather, ev[5], ev[6]);
    cudaEventElapsedTime(&st->ms_neighbour, ev[6], ev[7]);
    cudaEventElapsedTime(&st->ms_select, ev[7], ev[8]);
    st->ms_compare = tt.compare;
    st->ms_table_clear = tt.table_clear;
    st->ms_ingest_kernel = tt.ingest_kernel;
    st->ms_bucket_build = st->ms_neighbour - tt.compare;
    st->launches = tt.launches;
    st->plan_flags = (tt.partitioned ? FQD_PLAN_DEDUPE_PARTITIONED : 0u) | (tt.passes_partitioned ? FQD_PLAN_PASSES_PARTITIONED : 0u) |
                     (tt.pass0_fused ? FQD_PLAN_PASS0_FUSED : 0u) | (tt.pass1_emitted ? FQD_PLAN_PASS1_TILES_EMITTED : 0u);
    st->ms_partition_kernel = tt.partitioned ? tt.ingest_kernel : 0.f;
    st->ms_dedupe_kernel = tt.dedupe_kernel;
    publish_result(ctx, uq, f, job.n, c2.n_selected);
    return FQD_OK;
}

// ---- sharded plan ------------------------------------------------------------------------------------
//
// `S` holds one entry per rank driven by this process: exactly one when `ex` is an NCCL
// exchange (one process per GPU), all of them when ex == nullptr (virtual ranks, exchanged
// with device copies).  Phases are written "for every local shard"; collectives sit between.

struct Shard {
    fqd_context *ctx = nullptr;
    DeviceJob job;
    uint32_t index_base = 0;
    fqd_cluster_stats *st = nullptr;
    Uniques local, owned, all;
    Forest f;
    StageTimes tt;
    std::vector<uint32_t> send_cnt, recv_cnt;    // records per peer
    uint32_t *send = nullptr, *recv = nullptr;
    uint32_t n_recv = 0;
    uint2 *pairs[2] = {nullptr, nullptr};
    uint32_t n_pairs[2] = {0, 0};
};

int sync_all(std::vector<Shard> &S)
{
    for (auto &sh : S) {
        FQD_CUDA(cudaSetDevice(sh.ctx->device));
        FQD_CUDA(cudaStreamSynchronize(sh.ctx->stream));
    }
    return FQD_OK;
}

// every rank learns the n values of every rank
int gather_host_u64(std::vector<Shard> &S, Exchange *ex, int world, int n,
                    const std::vector<std::vector<uint64_t>> &mine, std::vector<uint64_t> &all)
{
    all.assign((size_t)world * n, 0);
    if (!ex) {
        for (int g = 0; g < world; g++)
            for (int i = 0; i < n; i++) all[(size_t)g * n + i] = mine[g][i];
        return FQD_OK;
    }
    fqd_context *ctx = S[0].ctx;
    const size_t mark = arena_mark(ctx);
    uint64_t *d_in, *d_out;
    FQD_TRY(arena(ctx, n, &d_in));
    FQD_TRY(arena(ctx, (size_t)world * n, &d_out));
    FQD_CUDA(cudaMemcpyAsync(d_in, mine[0].data(), (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    FQD_TRY(ex->allgather(d_in, d_out, (size_t)n * 8, ctx->stream));
    FQD_CUDA(cudaMemcpyAsync(all.data(), d_out, (size_t)world * n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FQD_CUDA(cudaStreamSynchronize(ctx->stream));
    arena_release(ctx, mark);
    return FQD_OK;
}

// variable all-gather of one device array per rank into dst (same layout on every rank)
int gather_device(std::vector<Shard> &S, Exchange *ex, int world, const std::vector<const void *> &src,
                  const std::vector<void *> &dst, const std::vector<size_t> &bytes)
{
    std::vector<size_t> off(world + 1, 0);
    for (int g = 0; g < world; g++) off[g + 1] = off[g] + bytes[g];
    if (ex) return ex->allgatherv(src[0], dst[0], off.data(), bytes.data(), S[0].ctx->stream);
    FQD_TRY(sync_all(S));
    for (int r = 0; r < world; r++) {
        FQD_CUDA(cudaSetDevice(S[r].ctx->device));
        for (int g = 0; g < world; g++)
            if (bytes[g])
                FQD_CUDA(cudaMemcpyAsync(static_cast<char *>(dst[r]) + off[g], src[g], bytes[g],
                                         cudaMemcpyDefault, S[r].ctx->stream));
    }
    return sync_all(S);
}

template <int K, int PW>
int run_sharded_typed(std::vector<Shard> &S, Exchange *ex, int world, const Codec &codec,
                      uint32_t unknown_out[8])
{
    constexpr int KW = K * PW, RW = slot_words(KW);
    const int L = (int)S.size();               // shards driven by this process
    auto rank_of = [&](int i) { return ex ? ex->rank : i; };
    EventSet e0, e1;   // per local rank: start / end of the job on its stream
    e0.ev.assign(L, nullptr);
    e1.ev.assign(L, nullptr);
    // FQD_TRACE=1: host wall-clock per phase (every phase ends synchronised)
    const bool trace = getenv("FQD_TRACE") && rank_of(0) == 0;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_prev = now();
    auto lap = [&](const char *what) {
        if (!trace) return;
        for (auto &sh : S) { cudaSetDevice(sh.ctx->device); cudaStreamSynchronize(sh.ctx->stream); }
        const double t = now();
        fprintf(stderr, "[fqd trace] %-28s %8.3f ms\n", what, t - t_prev);
        t_prev = t;
    };

    // ---- phase 1: local dedupe ----
    int rc_local = FQD_OK;
    for (int i = 0; i < L; i++) {
        Shard &sh = S[i];
        FQD_CUDA(cudaSetDevice(sh.ctx->device));
        sh.ctx->res = fqd_result{};
        sh.st->key_bits = K; sh.st->key_words = KW;
        FQD_CUDA(cudaEventCreate(&e0[i])); FQD_CUDA(cudaEventCreate(&e1[i]));
        FQD_CUDA(cudaEventRecord(e0[i], sh.ctx->stream));
        uint32_t unk[8] = {};
        // a failure of the local stage (arena, spill overflow, unsupported key) must not return before the exchange
        // below: the other ranks would wait in it forever.  It travels in the agreement vector instead.
        const int rc = stage_dedupe<K, PW>(sh.ctx, sh.job, codec, sh.index_base, true, sh.st, unk, sh.local, sh.tt);
        if (rc != FQD_OK && rc_local == FQD_OK) rc_local = rc;
        for (int k = 0; k < 8; k++) unknown_out[k] |= unk[k];
    }
    // error / alphabet agreement across ranks: [bad_record, bad_char, unknown x8, status of the local stage]
    {
        constexpr int AW = 11;
        std::vector<std::vector<uint64_t>> mine(L, std::vector<uint64_t>(AW));
        for (int i = 0; i < L; i++) {
            mine[i][0] = S[i].st->bad_record;
            mine[i][1] = S[i].st->bad_char;
            for (int k = 0; k < 8; k++) mine[i][2 + k] = unknown_out[k];
            mine[i][10] = (uint64_t)(int64_t)rc_local;
        }
        std::vector<uint64_t> all;
        FQD_TRY(gather_host_u64(S, ex, world, AW, mine, all));
        uint64_t bad = ~0ull, bad_char = 0;
        bool any_unknown = false;
        for (int g = 0; g < world; g++) {
            const int rc_g = (int)(int64_t)all[(size_t)g * AW + 10];
            if (rc_g != FQD_OK) {   // every rank leaves here, with the same status
                if (rc_local == FQD_OK) set_error("rank %d failed in its local dedupe stage (status %d)", g, rc_g);
                return rc_local != FQD_OK ? rc_local : rc_g;
            }
        }
        for (int g = 0; g < world; g++) {
            if (all[(size_t)g * AW] < bad) { bad = all[(size_t)g * AW]; bad_char = all[(size_t)g * AW + 1]; }
            for (int k = 0; k < 8; k++) {
                unknown_out[k] |= (uint32_t)all[(size_t)g * AW + 2 + k];
                any_unknown |= all[(size_t)g * AW + 2 + k] != 0;
            }
        }
        if (bad != ~0ull) {
            for (auto &sh : S) { sh.st->bad_record = bad; sh.st->bad_char = (uint32_t)bad_char; }
            set_error("Character %c outside of valid phred range ('%c' to '%c')", (int)bad_char,
                      (int)S[0].job.phred_offset, 126);
            return FQD_ERR_PHRED;
        }
        if (any_unknown) return RC_RETRY_ALPHABET;
    }

    lap("agreement");
    // ---- phase 2: send every local unique to its owner ----
    for (auto &sh : S) {
        FQD_CUDA(cudaSetDevice(sh.ctx->device));
        cudaStream_t s = sh.ctx->stream;
        const uint32_t U = sh.local.U;
        uint32_t *owner_cnt;
        FQD_TRY(arena(sh.ctx, 2 * 64, &owner_cnt));
        FQD_CUDA(cudaMemsetAsync(owner_cnt, 0, 2 * 64 * 4, s));
        if (U) owner_count_kernel<K, PW><<<cdiv(U, 256), 256, 0, s>>>(U, sh.local.ukey, (uint32_t)world, owner_cnt);
        sh.send_cnt.assign(world, 0);
        FQD_CUDA(cudaMemcpyAsync(sh.send_cnt.data(), owner_cnt, world * 4, cudaMemcpyDeviceToHost, s));
        FQD_CUDA(cudaStreamSynchronize(s));
        std::vector<uint32_t> cursor(world, 0);
        for (int g = 1; g < world; g++) cursor[g] = cursor[g - 1] + sh.send_cnt[g - 1];
        FQD_CUDA(cudaMemcpyAsync(owner_cnt + 64, cursor.data(), world * 4, cudaMemcpyHostToDevice, s));
        FQD_TRY(arena(sh.ctx, (size_t)std::max<uint32_t>(U, 1) * RW, &sh.send));
        if (U) owner_scatter_kernel<K, PW><<<cdiv(U, 256), 256, 0, s>>>(U, sh.local.ukey, sh.local.ucount, sh.local.ufirst,
                                                                       (uint32_t)world, owner_cnt + 64, sh.send);
        FQD_CUDA(cudaGetLastError());
        FQD_CUDA(cudaStreamSynchronize(s));   // `cursor` is host memory
        sh.tt.launches += 2;
    }
    lap("owner partition");
    std::vector<uint64_t> cnt_matrix;   // [src][dst]
    {
        std::vector<std::vector<uint64_t>> mine(L, std::vector<uint64_t>(world));
        for (int i = 0; i < L; i++) for (int g = 0; g < world; g++) mine[i][g] = S[i].send_cnt[g];
        FQD_TRY(gather_host_u64(S, ex, world, world, mine, cnt_matrix));
    }
    for (int i = 0; i < L; i++) {
        Shard &sh = S[i];
        const int r = rank_of(i);
        sh.recv_cnt.assign(world, 0);
        sh.n_recv = 0;
        for (int g = 0; g < world; g++) { sh.recv_cnt[g] = (uint32_t)cnt_matrix[(size_t)g * world + r]; sh.n_recv += sh.recv_cnt[g]; }
        FQD_CUDA(cudaSetDevice(sh.ctx->device));
        FQD_TRY(arena(sh.ctx, (size_t)std::max<uint32_t>(sh.n_recv, 1) * RW, &sh.recv));
    }
    if (ex) {
        Shard &sh = S[0];
        std::vector<size_t> so(world), sb(world), ro(world), rb(world);
        size_t a = 0, b = 0;
        for (int g = 0; g < world; g++) {
            so[g] = a; sb[g] = (size_t)sh.send_cnt[g] * RW * 4; a += sb[g];
            ro[g] = b; rb[g] = (size_t)sh.recv_cnt[g] * RW * 4; b += rb[g];
        }
        FQD_TRY(ex->alltoallv(sh.send, so.data(), sb.data(), sh.recv, ro.data(), rb.data(), sh.ctx->stream));
    } else {
        FQD_TRY(sync_all(S));
        for (int r = 0; r < world; r++) {
            FQD_CUDA(cudaSetDevice(S[r].ctx->device));
            size_t roff = 0;
            for (int g = 0; g < world; g++) {
                size_t soff = 0;
                for (int k = 0; k < r; k++) soff += (size_t)S[g].send_cnt[k] * RW * 4;
                const size_t bytes = (size_t)S[g].send_cnt[r] * RW * 4;
                if (bytes)
                    FQD_CUDA(cudaMemcpyAsync(reinterpret_cast<char *>(S[r].recv) + roff,
                                             reinterpret_cast<char *>(S[g].send) + soff, bytes,
                                             cudaMemcpyDefault, S[r].ctx->stream));
                roff += bytes;
            }
        }
        FQD_TRY(sync_all(S));
    }

    lap("all-to-all");
    // ---- phase 3: owners merge (sum of counts, min of first) ----
    for (auto &sh : S) {
        FQD_CUDA(cudaSetDevice(sh.ctx->device));
        cudaStream_t s = sh.ctx->stream;
        const uint64_t n = sh.n_recv;
        const uint64_t capacity = std::max<uint64_t>(1024, n + (n >> 1) + 64);
        uint32_t *table, *uslot, *kept;
        FQD_TRY(arena(sh.ctx, capacity * RW, &table));
        FQD_TRY(arena(sh.ctx, std::max<uint64_t>(n, 1), &uslot));
        FQD_TRY(arena(sh.ctx, 4, &kept));
        FQD_CUDA(cudaMemsetAsync(table, 0xFF, capacity * RW * 4, s));
        FQD_CUDA(cudaMemsetAsync(kept, 0, 16, s));
        FQD_CUDA(cudaMemsetAsync(&sh.ctx->d_ctr->n_unique, 0, 4, s));
        TableRef tab{table, capacity, uslot, sh.ctx->d_ctr};
        if (n) merge_insert_kernel<K, PW><<<cdiv(n, 256), 256, 0, s>>>((uint32_t)n, sh.recv, tab);
        FQD_CUDA(cudaGetLastError());
        FQD_TRY(fetch_counters(sh.ctx));
        if (sh.ctx->h_ctr->table_full) { set_error("internal: merge table overflow"); return FQD_ERR_NOMEM; }
        const uint32_t Um = sh.ctx->h_ctr->n_unique;
        FQD_TRY(arena(sh.ctx, (size_t)std::max<uint32_t>(Um, 1) * KW, &sh.owned.ukey));
        FQD_TRY(arena(sh.ctx, std::max<uint32_t>(Um, 1), &sh.owned.ucount));
        FQD_TRY(arena(sh.ctx, std::max<uint32_t>(Um, 1), &sh.owned.ufirst));
        if (Um) gather_nonzero_kernel<K, PW><<<cdiv(Um, 256), 256, 0, s>>>(Um, table, uslot, sh.owned.ukey,
                                                                          sh.owned.ucount, sh.owned.ufirst, kept);
        FQD_CUDA(cudaGetLastError());
        uint32_t h_kept = 0;
        FQD_CUDA(cudaMemcpyAsync(&h_kept, kept, 4, cudaMemcpyDeviceToHost, s));
        FQD_CUDA(cudaStreamSynchronize(s));
        sh.owned.U = h_kept;
        sh.tt.launches += 2;
    }

    lap("owner merge");
    // ---- phase 4: replicate the merged unique set on every rank ----
    std::vector<uint64_t> totals;
    {
        std::vector<std::vector<uint64_t>> mine(L, std::vector<uint64_t>(4));
        for (int i = 0; i < L; i++) {
            mine[i][0] = S[i].owned.U;
            mine[i][1] = S[i].st->total_records;
            mine[i][2] = S[i].st->discarded_records;
            mine[i][3] = S[i].st->number_of_sequences;
        }
        FQD_TRY(gather_host_u64(S, ex, world, 4, mine, totals));
    }
    uint64_t U_total = 0, n_total = 0, n_disc = 0, n_seq = 0;
    std::vector<size_t> Uo(world);
    for (int g = 0; g < world; g++) {
        Uo[g] = (size_t)totals[(size_t)g * 4];
        U_total += Uo[g]; n_total += totals[(size_t)g * 4 + 1];
        n_disc += totals[(size_t)g * 4 + 2]; n_seq += totals[(size_t)g * 4 + 3];
    }
    if (U_total > ENT_UID) { set_error("too many unique keys (%llu)", (unsigned long long)U_total); return FQD_ERR_UNSUPPORTED; }
    for (auto &sh : S) {
        FQD_CUDA(cudaSetDevice(sh.ctx->device));
        sh.all.U = (uint32_t)U_total;
        FQD_TRY(arena(sh.ctx, (size_t)std::max<uint64_t>(U_total, 1) * KW, &sh.all.ukey));
        FQD_TRY(arena(sh.ctx, std::max<uint64_t>(U_total, 1), &sh.all.ucount));
        FQD_TRY(arena(sh.ctx, std::max<uint64_t>(U_total, 1), &sh.all.ufirst));
    }
    for (int arr = 0; arr < 3; arr++) {
        std::vector<const void *> src(L);
        std::vector<void *> dst(L);
        std::vector<size_t> bytes(world);
        const size_t unit = arr == 0 ? (size_t)KW * 4 : 4;
        for (int g = 0; g < world; g++) bytes[g] = Uo[g] * unit;
        for (int i = 0; i < L; i++) {
            src[i] = arr == 0 ? (void *)S[i].owned.ukey : arr == 1 ? (void *)S[i].owned.ucount : (void *)S[i].owned.ufirst;
            dst[i] = arr == 0 ? (void *)S[i].all.ukey : arr == 1 ? (void *)S[i].all.ucount : (void *)S[i].all.ufirst;
        }
        FQD_TRY(gather_device(S, ex, world, src, dst, bytes));
    }

    lap("replicate (allgather)");
    // ---- phase 5: pigeonhole passes over the owned buckets ----
    const uint32_t U = (uint32_t)U_total;
    for (int i = 0; i < L; i++) {
        Shard &sh = S[i];
        FQD_CUDA(cudaSetDevice(sh.ctx->device));
        cudaStream_t s = sh.ctx->stream;
        FQD_TRY(stage_forest_alloc(sh.ctx, sh.job.method, U, sh.f));
        if (U) {
            init_forest_kernel<<<cdiv(U, 256), 256, 0, s>>>(U, sh.f.parent_full, sh.f.parent_one, sh.f.best);
            sh.tt.launches++;
        }
        FQD_CUDA(cudaMemsetAsync(&sh.ctx->d_ctr->n_merges, 0, 4, s));
        FQD_CUDA(cudaMemsetAsync(&sh.ctx->d_ctr->n_edges, 0, 8, s));
        FQD_CUDA(cudaMemsetAsync(&sh.ctx->d_ctr->n_candidates, 0, 8, s));
        FQD_TRY(stage_passes<K, PW>(sh.ctx, sh.job, codec, sh.all, sh.f, rank_of(i), world, sh.st, sh.tt));
    }

    lap("passes");
    // ---- phase 6: merge forests, flags and edge lists across ranks ----
    const int method = S[0].job.method;
    for (int which = 0; which < 2; which++) {
        if (which == 1 && method != METHOD_DIRECTIONAL) break;
        std::vector<std::vector<uint64_t>> mine(L, std::vector<uint64_t>(1));
        for (int i = 0; i < L; i++) {
            Shard &sh = S[i];
            FQD_CUDA(cudaSetDevice(sh.ctx->device));
            cudaStream_t s = sh.ctx->stream;
            uint32_t *np;
            FQD_TRY(arena(sh.ctx, std::max<uint32_t>(U, 1), &sh.pairs[which]));
            FQD_TRY(arena(sh.ctx, 4, &np));
            FQD_CUDA(cudaMemsetAsync(np, 0, 16, s));
            const uint32_t *parent = which == 0 ? sh.f.parent_full : sh.f.parent_one;
            if (U) forest_links_kernel<<<cdiv(U, 256), 256, 0, s>>>(U, parent, sh.pairs[which], np);
            FQD_CUDA(cudaGetLastError());
            uint32_t h = 0;
            FQD_CUDA(cudaMemcpyAsync(&h, np, 4, cudaMemcpyDeviceToHost, s));
            FQD_CUDA(cudaStreamSynchronize(s));
            sh.n_pairs[which] = h;
            mine[i][0] = h;
            sh.tt.launches++;
        }
        std::vector<uint64_t> all;
        FQD_TRY(gather_host_u64(S, ex, world, 1, mine, all));
        size_t total = 0;
        std::vector<size_t> bytes(world);
        for (int g = 0; g < world; g++) { bytes[g] = (size_t)all[g] * 8; total += (size_t)all[g]; }
        std::vector<const void *> src(L);
        std::vector<void *> dst(L);
        for (int i = 0; i < L; i++) {
            FQD_CUDA(cudaSetDevice(S[i].ctx->device));
            uint2 *buf;
            FQD_TRY(arena(S[i].ctx, std::max<size_t>(total, 1), &buf));
            src[i] = S[i].pairs[which];
            dst[i] = buf;
        }
        FQD_TRY(gather_device(S, ex, world, src, dst, bytes));
        for (int i = 0; i < L; i++) {
            FQD_CUDA(cudaSetDevice(S[i].ctx->device));
            uint32_t *parent = which == 0 ? S[i].f.parent_full : S[i].f.parent_one;
            if (total) apply_pairs_kernel<<<cdiv(total, 256), 256, 0, S[i].ctx->stream>>>((uint32_t)total, (const uint2 *)dst[i], parent);
            FQD_CUDA(cudaGetLastError());
            S[i].tt.launches++;
        }
    }
    if (method == METHOD_DIRECTIONAL && U) {
        for (int which = 0; which < 2; which++) {
            if (ex) {
                Shard &sh = S[0];
                FQD_TRY(ex->allreduce_max_u8(which ? sh.f.dead : sh.f.dominated, U, sh.ctx->stream));
            } else if (world > 1) {
                FQD_TRY(sync_all(S));
                uint8_t *acc = which ? S[0].f.dead : S[0].f.dominated;   // arena blocks are 256-byte padded
                const size_t n4 = ((size_t)U + 3) / 4;
                FQD_CUDA(cudaSetDevice(S[0].ctx->device));
                uint32_t *tmp;
                FQD_TRY(arena(S[0].ctx, n4, &tmp));
                for (int g = 1; g < world; g++) {
                    FQD_CUDA(cudaMemcpyAsync(tmp, which ? S[g].f.dead : S[g].f.dominated, U, cudaMemcpyDefault, S[0].ctx->stream));
                    max_u8_kernel<<<cdiv(n4, 256), 256, 0, S[0].ctx->stream>>>(n4, (uint32_t *)acc, tmp);
                }
                FQD_TRY(sync_all(S));
                for (int g = 1; g < world; g++) {
                    FQD_CUDA(cudaSetDevice(S[g].ctx->device));
                    FQD_CUDA(cudaMemcpyAsync(which ? S[g].f.dead : S[g].f.dominated, acc, U, cudaMemcpyDefault, S[g].ctx->stream));
                }
                FQD_TRY(sync_all(S));
            }
        }
    }
    if (method == METHOD_ADJACENCY) {
        std::vector<std::vector<uint64_t>> mine(L, std::vector<uint64_t>(1));
        for (int i = 0; i < L; i++) {
            FQD_CUDA(cudaSetDevice(S[i].ctx->device));
            FQD_TRY(fetch_counters(S[i].ctx));
            mine[i][0] = S[i].ctx->h_ctr->n_edges;
        }
        std::vector<uint64_t> all;
        FQD_TRY(gather_host_u64(S, ex, world, 1, mine, all));
        size_t total = 0;
        std::vector<size_t> bytes(world);
        for (int g = 0; g < world; g++) { bytes[g] = (size_t)all[g] * 8; total += (size_t)all[g]; }
        std::vector<const void *> src(L);
        std::vector<void *> dst(L);
        for (int i = 0; i < L; i++) {
            FQD_CUDA(cudaSetDevice(S[i].ctx->device));
            uint2 *buf;
            FQD_TRY(arena(S[i].ctx, std::max<size_t>(total, 1), &buf));
            src[i] = S[i].f.edges;
            dst[i] = buf;
        }
        FQD_TRY(gather_device(S, ex, world, src, dst, bytes));
        for (int i = 0; i < L; i++) { S[i].f.edges = (uint2 *)dst[i]; S[i].f.n_edges = total; S[i].f.edge_cap = total; }
    }

    lap("forest/flag merge");
    // ---- phase 7: every rank finishes the dissection; each writes the bitmap of its own records ----
    uint64_t cand_total = 0;
    {
        std::vector<std::vector<uint64_t>> mine(L, std::vector<uint64_t>(1));
        for (int i = 0; i < L; i++) {
            FQD_CUDA(cudaSetDevice(S[i].ctx->device));
            FQD_TRY(fetch_counters(S[i].ctx));
            mine[i][0] = S[i].ctx->h_ctr->n_candidates;
        }
        std::vector<uint64_t> all;
        FQD_TRY(gather_host_u64(S, ex, world, 1, mine, all));
        for (int g = 0; g < world; g++) cand_total += all[g];
    }
    for (int i = 0; i < L; i++) {
        Shard &sh = S[i];
        FQD_CUDA(cudaSetDevice(sh.ctx->device));
        cudaStream_t s = sh.ctx->stream;
        FQD_CUDA(cudaMemsetAsync(&sh.ctx->d_ctr->n_selected, 0, 4, s));
        FQD_TRY(stage_select<K, PW>(sh.ctx, sh.job, codec, sh.all, sh.f, sh.index_base, (uint32_t)sh.job.n, sh.tt, true));
        uint32_t *roots;
        FQD_TRY(arena(sh.ctx, 4, &roots));
        FQD_CUDA(cudaMemsetAsync(roots, 0, 16, s));
        if (U) count_roots_kernel<<<cdiv(U, 256), 256, 0, s>>>(U, sh.f.parent_full, roots);
        FQD_CUDA(cudaGetLastError());
        uint32_t h_roots = 0;
        FQD_CUDA(cudaMemcpyAsync(&h_roots, roots, 4, cudaMemcpyDeviceToHost, s));
        FQD_TRY(fetch_counters(sh.ctx));
        FQD_CUDA(cudaEventRecord(e1[i], s));
        FQD_CUDA(cudaStreamSynchronize(s));
        fqd_cluster_stats *st = sh.st;
        st->total_records = n_total;
        st->discarded_records = n_disc;
        st->number_of_sequences = n_seq;
        st->number_of_uniques = U_total;
        st->number_of_clusters = h_roots;
        st->number_selected = sh.ctx->h_ctr->n_selected;   // own keys only; summed below
        st->candidate_pairs = cand_total;
        cudaEventElapsedTime(&st->ms_total, e0[i], e1[i]);
        st->ms_ingest = sh.tt.ingest;
        st->ms_ingest_kernel = sh.tt.ingest_kernel;
        st->ms_table_clear = sh.tt.table_clear;
        st->ms_compare = sh.tt.compare;
        st->launches = sh.tt.launches + 1;
        st->plan_flags = (sh.tt.partitioned ? FQD_PLAN_DEDUPE_PARTITIONED : 0u) | (sh.tt.passes_partitioned ? FQD_PLAN_PASSES_PARTITIONED : 0u);
        st->ms_partition_kernel = sh.tt.partitioned ? sh.tt.ingest_kernel : 0.f;
        st->ms_dedupe_kernel = sh.tt.dedupe_kernel;
    }
    {
        std::vector<std::vector<uint64_t>> mine(L, std::vector<uint64_t>(1));
        for (int i = 0; i < L; i++) mine[i][0] = S[i].st->number_selected;
        std::vector<uint64_t> all;
        FQD_TRY(gather_host_u64(S, ex, world, 1, mine, all));
        uint64_t total = 0;
        for (int g = 0; g < world; g++) total += all[g];
        for (auto &sh : S) {
            sh.st->number_selected = total;
            publish_result(sh.ctx, sh.all, sh.f, n_total, total);
        }
    }
    lap("select + stats");
    return FQD_OK;
}

#include "sharded_tiles.cuh"

}  // namespace

}  // namespace fqd
