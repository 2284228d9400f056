// pipeline.cu -- launch plan of the clustering job on one B200 (see pipeline.cuh for the
// kernels and DESIGN.md for the data layout / rooflines).
#include <algorithm>

#include "common.h"

namespace fqd {

namespace {

inline uint32_t cdiv(uint64_t a, uint32_t b) { return (uint32_t)((a + b - 1) / b); }

struct Timer {
    fqd_context *ctx;
    int next = 0;
    explicit Timer(fqd_context *c) : ctx(c) {}
    int mark()
    {
        cudaEventRecord(ctx->ev[next], ctx->stream);
        return next++;
    }
    float ms(int a, int b)
    {
        float t = 0.f;
        cudaEventElapsedTime(&t, ctx->ev[a], ctx->ev[b]);
        return t;
    }
};

int fetch_counters(fqd_context *ctx)
{
    FQD_CUDA(cudaMemcpyAsync(ctx->h_ctr, ctx->d_ctr, sizeof(DevCounters), cudaMemcpyDeviceToHost,
                             ctx->stream));
    FQD_CUDA(cudaStreamSynchronize(ctx->stream));
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

template <int K, int PW>
int run_typed(fqd_context *ctx, const DeviceJob &job, const Codec &codec, fqd_cluster_stats *st,
              uint32_t unknown_out[8])
{
    constexpr int KW = K * PW, RW = slot_words(KW), FW = fat_words(KW);
    cudaStream_t s = ctx->stream;
    Timer tm(ctx);
    const uint64_t n = job.n;

    // ---- counters ----
    DevCounters zero{};
    zero.phred_err = ~0ull;
    zero.len_min = 0xFFFFFFFFu;
    *ctx->h_ctr = zero;
    FQD_CUDA(cudaMemcpyAsync(ctx->d_ctr, ctx->h_ctr, sizeof(DevCounters), cudaMemcpyHostToDevice, s));

    ctx->res = fqd_result{};
    ctx->res.n_records = n;

    st->key_bits = K;
    st->key_words = KW;

    const int t_begin = tm.mark();

    // ---- ingest: filter + pack + exact dedupe into the HBM table ----
    const uint64_t capacity = std::max<uint64_t>(1024, n + (n >> 1) + 64);
    if (capacity >= 0xFFFFFFF0ull) {
        set_error("too many records for one job on one GPU (%llu)", (unsigned long long)n);
        return FQD_ERR_UNSUPPORTED;
    }
    DevBuf table, uslot, keepmask;
    FQD_TRY(table.alloc(ctx, capacity * RW * sizeof(uint32_t)));
    FQD_TRY(uslot.alloc(ctx, n * sizeof(uint32_t)));
    FQD_TRY(keepmask.alloc(ctx, (size_t)cdiv(n, 32) * sizeof(uint32_t)));
    FQD_CUDA(cudaMemsetAsync(table.p, 0xFF, capacity * RW * sizeof(uint32_t), s));
    const int t_cleared = tm.mark();
    uint32_t launches = 0;

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
    ip.table = table.as<uint32_t>();
    ip.capacity = capacity;
    ip.uslot = uslot.as<uint32_t>();
    ip.keepmask = keepmask.as<uint32_t>();
    ip.weights = job.weights;
    ip.ctr = ctx->d_ctr;
    ip.codec = codec;
    // shared-memory staging of fixed-stride rows
    uint32_t stride = 0;
    if (!job.key_off) stride = job.key_stride;
    if (job.filter_on && !job.qual_off) stride = std::max(stride, job.qual_stride);
    const bool fixed_any = !job.key_off || (job.filter_on && !job.qual_off);
    size_t smem = 1280;
    if (fixed_any && (size_t)stride * 256 + 1280 <= 200 * 1024) {
        ip.stage_bytes = stride * 256;
        smem += ip.stage_bytes;
    }
    FQD_CUDA(cudaFuncSetAttribute(ingest_kernel<K, PW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    ip.phase = 0;
    ingest_kernel<K, PW><<<cdiv(n, 256), 256, smem, s>>>(ip);
    launches++;
    const int t_ingest_k = tm.mark();
    FQD_CUDA(cudaGetLastError());
    FQD_TRY(fetch_counters(ctx));
    const DevCounters c1 = *ctx->h_ctr;
    if (c1.phred_err != ~0ull) {
        st->bad_record = c1.phred_err >> 8;
        st->bad_char = (uint32_t)(c1.phred_err & 0xFF);
        set_error("Character %c outside of valid phred range ('%c' to '%c')",
                  (int)st->bad_char, (int)job.phred_offset, 126);
        return FQD_ERR_PHRED;
    }
    bool any_unknown = false;
    for (int i = 0; i < 8; i++) { unknown_out[i] = c1.unknown[i]; any_unknown |= c1.unknown[i] != 0; }
    if (any_unknown) return RC_RETRY_ALPHABET;
    if (c1.table_full) { set_error("internal: dedupe table overflow"); return FQD_ERR_NOMEM; }
    if (job.filter_on && c1.n_discarded) {
        ip.phase = 1;
        ingest_kernel<K, PW><<<cdiv(n, 256), 256, smem, s>>>(ip);
        launches++;
        FQD_CUDA(cudaGetLastError());
    }
    const uint32_t U = c1.n_unique;
    st->total_records = n;
    st->discarded_records = c1.n_discarded;
    st->number_of_sequences = job.weights ? c1.sum_weights : n - c1.n_discarded;
    st->number_of_uniques = U;
    const int t_ingest = tm.mark();
    if (U > ENT_UID) { set_error("too many unique keys for one GPU (%u)", U); return FQD_ERR_UNSUPPORTED; }

    // ---- gather ----
    const bool directional = job.method == METHOD_DIRECTIONAL;
    DevBuf ukey, ucount, ufirst, parent_full, parent_one, best, selected;
    FQD_TRY(ukey.alloc(ctx, (size_t)U * KW * 4));
    FQD_TRY(ucount.alloc(ctx, (size_t)U * 4));
    FQD_TRY(ufirst.alloc(ctx, (size_t)U * 4));
    FQD_TRY(parent_full.alloc(ctx, (size_t)U * 4));
    FQD_TRY(selected.alloc(ctx, (size_t)U));
    if (directional) FQD_TRY(parent_one.alloc(ctx, (size_t)U * 4));
    if (job.method != METHOD_ADJACENCY) FQD_TRY(best.alloc(ctx, (size_t)U * 4));
    if (U) {
        gather_kernel<K, PW><<<cdiv(U, 256), 256, 0, s>>>(U, table.as<uint32_t>(), uslot.as<uint32_t>(),
                                                          ukey.as<uint32_t>(), ucount.as<uint32_t>(),
                                                          ufirst.as<uint32_t>(), parent_full.as<uint32_t>(),
                                                          parent_one.as<uint32_t>(), best.as<uint32_t>());
        launches++;
    }
    FQD_CUDA(cudaGetLastError());
    table.reset(); uslot.reset(); keepmask.reset();
    const int t_gather = tm.mark();

    // ---- pigeonhole passes ----
    DevBuf dominated, dead, deadroot, edges, rootbuf;
    FQD_TRY(rootbuf.alloc(ctx, (size_t)U * 4));
    if (directional) {
        FQD_TRY(dominated.alloc(ctx, U)); FQD_TRY(dead.alloc(ctx, U)); FQD_TRY(deadroot.alloc(ctx, U));
        FQD_CUDA(cudaMemsetAsync(dominated.p, 0, U ? U : 1, s));
        FQD_CUDA(cudaMemsetAsync(dead.p, 0, U ? U : 1, s));
        FQD_CUDA(cudaMemsetAsync(deadroot.p, 0, U ? U : 1, s));
    }
    unsigned long long edge_cap = 0;
    if (job.method == METHOD_ADJACENCY) {
        edge_cap = 2ull * U + (1ull << 16);
        FQD_TRY(edges.alloc(ctx, edge_cap * sizeof(uint2)));
    }
    const int npass = (job.d > 0 && U > 1) ? job.d + 1 : 0;
    st->n_passes = npass;
    float ms_compare = 0.f;
    if (npass) {
        const int V = job.edit ? (job.varlen ? 2 * job.d + 1 : 1) * (job.d + 1) : 1;
        const uint64_t E = (uint64_t)U * V;
        if (E >= 0xFFFFFFF0ull) { set_error("too many pigeonhole entries (%llu)", (unsigned long long)E); return FQD_ERR_UNSUPPORTED; }
        uint32_t NB = 1024;
        while (NB < (1u << 24) && NB < E / 2) NB <<= 1;
        DevBuf cnt, rank, entries, block_sums, grand;
        FQD_TRY(cnt.alloc(ctx, ((size_t)NB + 1) * 4));
        FQD_TRY(rank.alloc(ctx, E * 4));
        const bool fat = !job.edit;
        FQD_TRY(entries.alloc(ctx, fat ? E * FW * sizeof(uint32_t) : E * sizeof(uint2)));
        FQD_TRY(block_sums.alloc(ctx, (size_t)cdiv(NB, SCAN_TILE) * 4 + 64));
        FQD_TRY(grand.alloc(ctx, 16));
        PassParams pp{};
        pp.U = U; pp.ukey = ukey.as<uint32_t>(); pp.ucount = ucount.as<uint32_t>();
        pp.d = job.d; pp.edit = job.edit; pp.varlen = job.varlen ? 1 : 0; pp.method = job.method;
        pp.max_len = job.max_len; pp.pad_code = codec.pad_code;
        pp.V = V; pp.nb_mask = NB - 1;
        pp.cnt = cnt.as<uint32_t>(); pp.rank = rank.as<uint32_t>(); pp.entries = entries.as<uint2>(); pp.fat = entries.as<uint32_t>();
        pp.parent_full = parent_full.as<uint32_t>(); pp.parent_one = parent_one.as<uint32_t>();
        pp.dominated = dominated.as<uint8_t>(); pp.dead = dead.as<uint8_t>();
        pp.edges = edges.as<uint2>(); pp.edge_cap = edge_cap; pp.ctr = ctx->d_ctr;
        for (int i = 0; i < 256; i++) pp.rank_of_code[i] = codec.rank[i];
        std::vector<cudaEvent_t> cev(2 * npass);
        for (auto &e : cev) FQD_CUDA(cudaEventCreate(&e));
        for (int attempt = 0; attempt < 2; attempt++) {
            for (int j = 0; j < npass; j++) {
                pp.pass_j = j;
                FQD_CUDA(cudaMemsetAsync(cnt.p, 0, ((size_t)NB + 1) * 4, s));
                sig_count_kernel<K, PW><<<cdiv(U, 256), 256, 0, s>>>(pp);
                FQD_TRY(exclusive_scan_inplace(ctx, cnt.as<uint32_t>(), NB, block_sums.as<uint32_t>(),
                                               grand.as<uint32_t>()));
                if (fat) scatter_fat_kernel<K, PW><<<cdiv(U, 256), 256, 0, s>>>(pp);
                else scatter_kernel<K, PW><<<cdiv(U, 256), 256, 0, s>>>(pp);
                pp.n_entries = (uint32_t)E;   // upper bound; the kernel stops at cnt[NB]
                launches += 6;   // sig_count, 3 scan kernels, scatter, compare
                FQD_CUDA(cudaEventRecord(cev[2 * j], s));
                if (fat) compare_fat_kernel<K, PW><<<cdiv(E, 256), 256, 0, s>>>(pp);
                else compare_kernel<K, PW><<<cdiv(E, 256), 256, 0, s>>>(pp);
                FQD_CUDA(cudaEventRecord(cev[2 * j + 1], s));
                FQD_CUDA(cudaGetLastError());
            }
            if (job.method != METHOD_ADJACENCY) break;
            FQD_TRY(fetch_counters(ctx));
            if (ctx->h_ctr->n_edges <= edge_cap) break;
            if (attempt == 1) { set_error("internal: adjacency edge list overflow"); return FQD_ERR_NOMEM; }
            // edge list overflowed: size it exactly, reset the forest and redo the passes
            edge_cap = ctx->h_ctr->n_edges + 16;
            FQD_TRY(edges.alloc(ctx, edge_cap * sizeof(uint2)));
            pp.edges = edges.as<uint2>(); pp.edge_cap = edge_cap;
            ctx->h_ctr->n_edges = 0; ctx->h_ctr->n_merges = 0; ctx->h_ctr->n_candidates = 0;
            FQD_CUDA(cudaMemcpyAsync(ctx->d_ctr, ctx->h_ctr, sizeof(DevCounters), cudaMemcpyHostToDevice, s));
            iota_kernel<<<cdiv(U, 256), 256, 0, s>>>(parent_full.as<uint32_t>(), U);
            launches++;
        }
        FQD_CUDA(cudaStreamSynchronize(s));
        for (int j = 0; j < npass; j++) {
            float t = 0.f;
            cudaEventElapsedTime(&t, cev[2 * j], cev[2 * j + 1]);
            ms_compare += t;
        }
        for (auto &e : cev) cudaEventDestroy(e);
    }
    const int t_pass = tm.mark();

    // ---- components + dissection ----
    SelectParams sp{};
    sp.U = U; sp.ukey = ukey.as<uint32_t>(); sp.ucount = ucount.as<uint32_t>(); sp.ufirst = ufirst.as<uint32_t>();
    sp.parent_full = parent_full.as<uint32_t>(); sp.parent_one = parent_one.as<uint32_t>();
    sp.best = best.as<uint32_t>(); sp.root = rootbuf.as<uint32_t>();
    sp.dominated = dominated.as<uint8_t>(); sp.dead = dead.as<uint8_t>(); sp.deadroot = deadroot.as<uint8_t>();
    sp.selected = selected.as<uint8_t>();
    sp.method = job.method; sp.ctr = ctx->d_ctr; sp.bitmap = job.bitmap;
    for (int i = 0; i < 256; i++) sp.rank_of_code[i] = codec.rank[i];
    if (job.bitmap) FQD_CUDA(cudaMemsetAsync(job.bitmap, 0, (size_t)cdiv(n, 32) * 4, s));
    DevBuf state, stamp;
    if (U) {
        if (job.method == METHOD_DIRECTIONAL) {
            root_best_kernel<K, PW><<<cdiv(U, 256), 256, 0, s>>>(sp, 1);
            launches++;
        } else if (job.method == METHOD_HIGHEST) {
            root_best_kernel<K, PW><<<cdiv(U, 256), 256, 0, s>>>(sp, 0);
            launches++;
        } else {
            FQD_TRY(state.alloc(ctx, U)); FQD_TRY(stamp.alloc(ctx, (size_t)U * 4));
            FQD_CUDA(cudaMemsetAsync(state.p, 0, U, s));
            FQD_CUDA(cudaMemsetAsync(stamp.p, 0, (size_t)U * 4, s));
            sp.state = state.as<uint8_t>(); sp.stamp = stamp.as<uint32_t>();
            sp.edges = edges.as<uint2>();
            FQD_TRY(fetch_counters(ctx));
            sp.n_edges = std::min<unsigned long long>(ctx->h_ctr->n_edges, edge_cap);
            if (ctx->h_ctr->n_edges > edge_cap) { set_error("internal: adjacency edge list overflow"); return FQD_ERR_NOMEM; }
            for (uint32_t round = 1;; round++) {
                sp.round = round;
                FQD_CUDA(cudaMemsetAsync(&ctx->d_ctr->undecided, 0, 4, s));
                if (sp.n_edges) { adj_edge_kernel<<<cdiv(sp.n_edges, 256), 256, 0, s>>>(sp); launches++; }
                adj_node_kernel<<<cdiv(U, 256), 256, 0, s>>>(sp);
                launches++;
                FQD_TRY(fetch_counters(ctx));
                if (ctx->h_ctr->undecided == 0) break;
                if (round > U + 2) { set_error("internal: adjacency rounds did not converge"); return FQD_ERR_CUDA; }
            }
        }
        select_kernel<<<cdiv(U, 256), 256, 0, s>>>(sp);
        launches++;
        FQD_CUDA(cudaGetLastError());
    }
    FQD_TRY(fetch_counters(ctx));
    const int t_end = tm.mark();
    FQD_CUDA(cudaStreamSynchronize(s));
    const DevCounters c2 = *ctx->h_ctr;
    st->number_of_clusters = (uint64_t)U - c2.n_merges;
    st->number_selected = c2.n_selected;
    st->candidate_pairs = c2.n_candidates;
    st->ms_total = tm.ms(t_begin, t_end);
    st->ms_ingest = tm.ms(t_begin, t_ingest);
    st->ms_gather = tm.ms(t_ingest, t_gather);
    st->ms_neighbour = tm.ms(t_gather, t_pass);
    st->ms_select = tm.ms(t_pass, t_end);
    st->ms_compare = ms_compare;
    st->ms_table_clear = tm.ms(t_begin, t_cleared);
    st->ms_ingest_kernel = tm.ms(t_cleared, t_ingest_k);
    st->ms_bucket_build = st->ms_neighbour - ms_compare;
    st->launches = launches;

    ctx->res.U = U;
    ctx->res.n_selected = c2.n_selected;
    ctx->res.ufirst = (uint32_t *)ufirst.release();
    ctx->res.ucount = (uint32_t *)ucount.release();
    ctx->res.parent_full = (uint32_t *)parent_full.release();
    ctx->res.selected = (uint8_t *)selected.release();
    return FQD_OK;
}

}  // namespace

// (K, PW) instantiations of this build: key length <= 32*PW symbols, alphabet (+PAD) < 2^K.
#define FQD_INSTANCES(X) \
    X(3, 1) X(3, 2) X(3, 3) X(3, 4) X(3, 5) X(3, 8) X(3, 10) \
    X(4, 1) X(4, 2) X(4, 4) X(4, 8)                           \
    X(8, 1) X(8, 2) X(8, 4)

int supported_bits(int needed)
{
    if (needed <= 3) return 3;
    if (needed <= 4) return 4;
    if (needed <= 8) return 8;
    return 0;
}

uint32_t max_supported_length(int bits)
{
    uint32_t best = 0;
#define X(K_, PW_) if (bits == K_ && 32u * PW_ > best) best = 32u * PW_;
    FQD_INSTANCES(X)
#undef X
    return best;
}

int run_pipeline(fqd_context *ctx, const DeviceJob &job, const Codec &codec,
                 fqd_cluster_stats *stats, uint32_t unknown_out[8])
{
    const int bits = codec.bits;
    const uint32_t pw_needed = std::max(1u, (job.max_len + 31u) / 32u);
    // smallest instantiated PW that fits
    int best_pw = 0;
#define X(K_, PW_) if (bits == K_ && (uint32_t)PW_ >= pw_needed && (best_pw == 0 || PW_ < best_pw)) best_pw = PW_;
    FQD_INSTANCES(X)
#undef X
    if (!best_pw) {
        set_error("keys of %u symbols over a %d-bit alphabet exceed what this build packs "
                  "(max %u symbols)", job.max_len, bits, max_supported_length(bits));
        return FQD_ERR_UNSUPPORTED;
    }
#define X(K_, PW_) if (bits == K_ && best_pw == PW_) return run_typed<K_, PW_>(ctx, job, codec, stats, unknown_out);
    FQD_INSTANCES(X)
#undef X
    return FQD_ERR_UNSUPPORTED;
}

}  // namespace fqd
