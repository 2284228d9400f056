// sharded_tiles.cuh -- the tile-sharded plan of a multi-GPU job (included by pipeline_impl.cuh inside its
// anonymous namespace).  DESIGN.md section 7.
//
// The streaming plan of one GPU (partitioned.cuh) spreads the records over shared-memory sized tiles by a hash of
// their pigeonhole block 0; a tile is deduplicated, searched (pass 0) and re-emitted for pass 1 by one thread block.
// Sharding it over G ranks needs no new algorithm, only an owner per tile:
//
//   1. every rank packs ITS records and appends them to its own copy of the tile regions (each ~1/G full);
//   2. the fill counters are exchanged (one small all-to-all -- it is also the barrier between the phases);
//   3. the owner's dedupe_tile_kernel stages a tile by fetching its G fragments straight out of the peers' HBM
//      (cp.async.bulk on peer memory over NVLink): the all-to-all of the records is fused into the tile kernel;
//      unique ids are job-wide (local id * G + rank);
//   4. the uniques leave for the tiles of the next pass the same way (emit locally, exchange counters, the owner
//      of a pass tile pulls its fragments);
//   5. nothing per-key is replicated except the union-find forests: every rank's apply_edges_kernel streams the
//      edge lists of ALL ranks out of their HBM while it hooks (fused all-gather), and an edge carries its
//      consequence for the dissection in its state bits;
//   6. keys decide on their owner; where a component's answer needs the keys of several members, candidate records
//      are reduced by every rank from the peers' lists; the keep bit of a selected key is OR-ed into the bitmap of
//      the rank that holds the record (remote atomic).
//
// What the plan cannot do (a pigeonhole bucket or a key family that outgrows a tile beyond the spill path, more
// edges than the lists hold) is detected, agreed on by all ranks, and the job is handed to the replicated-set
// plan (run_sharded_typed), which has no size assumptions.
//
// The same code drives (a) one rank per process, collectives over NCCL, peer memory through CUDA IPC, and (b) all
// ranks inside one process as virtual ranks (tests on a one-GPU box): then a "peer" pointer is simply a pointer.

struct TileRank {
    // buffers other ranks read or write (identical arena offsets on every rank)
    uint32_t *tiles0 = nullptr, *spill = nullptr, *tilesN = nullptr, *tilesP[2] = {nullptr, nullptr};
    uint2 *edges = nullptr, *adj = nullptr;
    uint32_t *cand = nullptr, *cand_root = nullptr, *bitmap = nullptr;
    // private
    uint32_t *cursor0 = nullptr, *cnt0 = nullptr, *cursorN = nullptr, *cntN = nullptr, *cursorP = nullptr, *cntP = nullptr;
    uint32_t *ctrs = nullptr;        // device counters of the plan, see CTR_*
    uint32_t *oversize = nullptr;
    uint32_t *gath_in = nullptr, *gath = nullptr;   // small device-side gathers: [G][4]
    uint32_t *stat_in = nullptr, *stat_all = nullptr;   // status blocks (STATUS_WORDS uint64 per rank)
    uint32_t *a2a_send = nullptr;                   // counter exchange: G blocks of (nper counters, edge count)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;   // the first edge application runs beside the next pass
    uint32_t U_tiles = 0, U = 0;     // uniques out of the tiles / upper bound including the spill path
    uint32_t *root_of = nullptr, *loc_of = nullptr;
    uint8_t *linked = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = FQD_OK;
};
enum { CTR_SPILL = 0, CTR_SPILL_OVER = 1, CTR_UNIQUE = 2, CTR_OVERSIZE = 3, CTR_CLAIMED = 4, CTR_EDGES = 5, CTR_EDGE_OVER = 6,
       CTR_UNIQUE_OVER = 7, CTR_CAND = 8, CTR_WORDS = 16 };

template <int K, int PW>
int run_sharded_tiles(std::vector<Shard> &S, Exchange *ex, const ShardWorld &W, const Codec &codec, uint32_t unknown_out[8])
{
    constexpr int KW = K * PW, RW = slot_words(KW);
    if constexpr (RW != PART_RW) {
        return RC_FALLBACK_REPLICATED;
    } else {
    const int L = (int)S.size(), G = W.world;
    auto rank_of = [&](int i) { return ex ? ex->rank : i; };
    const DeviceJob &j0 = S[0].job;
    const int d = j0.d, method = j0.method;
    const uint64_t N = W.n_total;
    // FQD_TRACE=1: device time per phase of this process' first rank (CUDA events on its stream, no extra
    // synchronisation; printed by rank 0, by every rank with FQD_TRACE=2)
    const char *trace_env = getenv("FQD_TRACE");
    const bool trace = trace_env && (rank_of(0) == 0 || atoi(trace_env) >= 2);
    std::vector<std::pair<const char *, cudaEvent_t>> marks_ev;
    auto lap = [&](const char *what) {
        if (!trace) return;
        cudaEvent_t e;
        cudaSetDevice(S[0].ctx->device);
        if (cudaEventCreate(&e) != cudaSuccess) return;
        cudaEventRecord(e, S[0].ctx->stream);
        marks_ev.emplace_back(what, e);
    };
    auto lap_report = [&]() {
        if (marks_ev.empty()) return;
        cudaSetDevice(S[0].ctx->device);
        cudaStreamSynchronize(S[0].ctx->stream);
        std::string line = "[fqd trace] rank " + std::to_string(rank_of(0)) + " tiles (device ms):";
        float total = 0.f;
        for (size_t k = 1; k < marks_ev.size(); k++) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, marks_ev[k - 1].second, marks_ev[k].second);
            total += ms;
            char buf[96];
            snprintf(buf, sizeof buf, " %s %.3f |", marks_ev[k].first, ms);
            line += buf;
        }
        char buf[48];
        snprintf(buf, sizeof buf, " total %.3f", total);
        fprintf(stderr, "%s%s\n", line.c_str(), buf);
        for (auto &m : marks_ev) cudaEventDestroy(m.second);
        marks_ev.clear();
    };
    lap("start");
    std::vector<TileRank> T(L);
    struct Cleanup {
        std::vector<TileRank> &T;
        ~Cleanup()
        {
            for (auto &t : T)
                for (cudaEvent_t e : {t.e0, t.e1, t.ev_fork, t.ev_join})
                    if (e) cudaEventDestroy(e);
        }
    } cleanup{T};

    // ---- sizes every rank derives from the agreed numbers (so that the shared buffers sit at the same offsets) ----
    const bool fused = d >= 1 && method != METHOD_ADJACENCY && !getenv("FQD_NO_FUSED_PASS0");
    const uint32_t nper0 = cdiv(tile_partitions(std::max<uint64_t>(N, 1)), (uint32_t)G), nparts0 = nper0 * (uint32_t)G;
    const uint64_t guessU = std::max<uint64_t>(N / 2, 1u << 16);
    const bool emit_next = fused && !getenv("FQD_NO_NEXT_EMIT");
    const uint32_t nperN = cdiv(tile_partitions(guessU), (uint32_t)G), npartsN = nperN * (uint32_t)G;
    // a rank's region of a tile receives ~1/G of it: sized for that plus Poisson slack (surplus -> spill path)
    const uint32_t fill_pct = std::min(90u, std::max(20u, env_u32("FQD_TILE_FILL_PCT", 60)));
    const uint32_t avg_frag = (uint32_t)TILE_R * fill_pct / (100u * (uint32_t)G);
    const uint32_t region = getenv("FQD_SHARD_FULL_REGIONS") ? (uint32_t)TILE_R
                                                             : std::min<uint32_t>(TILE_R, ((avg_frag * 3 / 2 + 48 + 31) / 32) * 32);
    const int peer_ldg = getenv("FQD_PEER_LDG") ? 1 : 0;
    const uint32_t spill_cap = (uint32_t)(W.n_max / 4 + 4096);
    const uint32_t cap_u = (uint32_t)std::min<uint64_t>((EDGE_ID + 1ull) / (uint64_t)G - 1, N / G + N / (4ull * G) + (1u << 16));
    const uint32_t cap_e = (uint32_t)std::min<uint64_t>(EDGE_ID, W.n_max + (1u << 16));
    const uint32_t cap_c = std::min<uint32_t>(cap_u, (1u << 28) - 1);
    const uint64_t cap_adj = method == METHOD_ADJACENCY ? 2 * W.n_max + (1u << 16) : 0;
    const uint32_t bm_words = (uint32_t)((N + 31) / 32 + 2);   // every rank marks its selected keys over the records of the whole job

    // peer view of a pointer into this rank's slab
    auto in_slab = [&](const Shard &sh, const void *p, size_t bytes) {
        const char *c = static_cast<const char *>(p);
        return c >= sh.ctx->arena.base && c + bytes <= sh.ctx->arena.base + sh.ctx->arena.cap;
    };
    auto peer_ptr = [&](const Shard &sh, int g, const void *mine) {
        return W.peer_base[g] + (static_cast<const char *>(mine) - sh.ctx->arena.base);
    };

    // ---- exchanges ----
    // Counter exchange (all-to-all): rank g sends rank r the fill counters of r's tiles in g's buffer plus g's current
    // edge count; received as cnt[g * (nper + 1) + j].  It is also the barrier between a fill phase and the tile kernels.
    auto exchange_counters = [&](auto cursor_of, auto cnt_of, uint32_t nper) -> int {
        const size_t words = (size_t)nper + 1;
        for (int i = 0; i < L; i++) {
            if (T[i].rc != FQD_OK || !T[i].a2a_send) continue;
            FQD_CUDA(cudaSetDevice(S[i].ctx->device));
            pack_counters_kernel<<<cdiv(words * G, 256), 256, 0, S[i].ctx->stream>>>(cursor_of(i), nper, (uint32_t)G, T[i].ctrs + CTR_EDGES,
                                                                                      T[i].a2a_send);
            FQD_CUDA(cudaGetLastError());
        }
        if (ex) {
            std::vector<size_t> off(G), bytes(G, words * 4);
            for (int g = 0; g < G; g++) off[g] = (size_t)g * words * 4;
            return ex->alltoallv(T[0].a2a_send, off.data(), bytes.data(), cnt_of(0), off.data(), bytes.data(), S[0].ctx->stream);
        }
        FQD_TRY(sync_all(S));
        for (int r = 0; r < G; r++) {
            FQD_CUDA(cudaSetDevice(S[r].ctx->device));
            for (int g = 0; g < G; g++)
                FQD_CUDA(cudaMemcpyAsync(cnt_of(r) + (size_t)g * words, T[g].a2a_send + (size_t)r * words, words * 4, cudaMemcpyDefault,
                                         S[r].ctx->stream));
        }
        return sync_all(S);
    };
    // all-gather of `words` uint32 per rank (device to device, stream ordered; also the barrier between phases)
    auto allgather_u32 = [&](auto src_of, auto dst_of, size_t words) -> int {
        if (ex) return ex->allgather(src_of(0), dst_of(0), words * 4, S[0].ctx->stream);
        FQD_TRY(sync_all(S));
        for (int r = 0; r < G; r++) {
            FQD_CUDA(cudaSetDevice(S[r].ctx->device));
            for (int g = 0; g < G; g++)
                FQD_CUDA(cudaMemcpyAsync(dst_of(r) + (size_t)g * words, src_of(g), words * 4, cudaMemcpyDefault, S[r].ctx->stream));
        }
        return sync_all(S);
    };
    // Agreement point: every rank's status block (packed on the device by pack_status_kernel: no host round trip
    // before the exchange) on every rank -- ONE host synchronisation.  `tail` is enqueued behind the exchange, before
    // the synchronisation.  Returns the first failing status, the same on all ranks.
    std::vector<uint64_t> all((size_t)G * STATUS_WORDS, 0);
    auto agree = [&](auto tail) -> int {
        for (int i = 0; i < L; i++) {
            FQD_CUDA(cudaSetDevice(S[i].ctx->device));
            StatusParams sp{(long long)T[i].rc, T[i].ctrs, S[i].ctx->d_ctr, cap_u, spill_cap, cap_e, cap_c,
                            reinterpret_cast<unsigned long long *>(T[i].stat_in)};
            if (!T[i].ctrs || !T[i].stat_in) { set_error("internal: status buffers were not allocated"); return FQD_ERR_NOMEM; }
            pack_status_kernel<<<1, 32, 0, S[i].ctx->stream>>>(sp);
            FQD_CUDA(cudaGetLastError());
        }
        FQD_TRY(allgather_u32([&](int i) { return T[i].stat_in; }, [&](int i) { return T[i].stat_all; }, (size_t)STATUS_WORDS * 2));
        for (int i = 0; i < L; i++) {
            FQD_CUDA(cudaSetDevice(S[i].ctx->device));
            FQD_TRY(tail(i));
            if (i == 0)
                FQD_CUDA(cudaMemcpyAsync(all.data(), T[0].stat_all, (size_t)G * STATUS_WORDS * 8, cudaMemcpyDeviceToHost, S[0].ctx->stream));
        }
        FQD_TRY(sync_all(S));
        for (int g = 0; g < G; g++) {
            const int rc = (int)(long long)all[(size_t)g * STATUS_WORDS];
            if (rc != FQD_OK) {
                bool mine_failed = false;
                for (int i = 0; i < L; i++) mine_failed |= T[i].rc == rc;
                if (!mine_failed && rc > 0) set_error("rank %d of the sharded job failed with status %d", g, rc);
                return rc;
            }
        }
        return FQD_OK;
    };
    auto no_tail = [](int) { return FQD_OK; };

    // =====================================================================================================
    // phase 1: buffers
    // =====================================================================================================
    for (int i = 0; i < L; i++) {
        Shard &sh = S[i];
        TileRank &t = T[i];
        fqd_context *ctx = sh.ctx;
        cudaStream_t s = ctx->stream;
        t.rc = [&]() -> int {
            FQD_CUDA(cudaSetDevice(ctx->device));
            ctx->res = fqd_result{};
            sh.st->key_bits = K; sh.st->key_words = KW;
            FQD_CUDA(cudaEventCreate(&t.e0)); FQD_CUDA(cudaEventCreate(&t.e1));
            FQD_CUDA(cudaEventRecord(t.e0, s));
            FQD_TRY(reset_counters(ctx));
            if (ctx->arena.off > W.shared_off) { set_error("internal: arena offsets of the ranks disagree"); return FQD_ERR_CUDA; }
            ctx->arena.off = W.shared_off;
            bool inside = true;
            auto shared = [&](auto **p, size_t count) -> int {
                using Elem = std::remove_pointer_t<std::remove_pointer_t<decltype(p)>>;
                FQD_TRY(arena(ctx, count, p));
                inside = inside && in_slab(sh, *p, std::max<size_t>(count * sizeof(Elem), 16));
                return FQD_OK;
            };
            FQD_TRY(shared(&t.tiles0, (size_t)nparts0 * region * RW));
            FQD_TRY(shared(&t.spill, (size_t)spill_cap * RW));
            if (emit_next) FQD_TRY(shared(&t.tilesN, (size_t)npartsN * region * RW));
            FQD_TRY(shared(&t.edges, (size_t)cap_e));
            if (cap_adj) FQD_TRY(shared(&t.adj, (size_t)cap_adj));
            FQD_TRY(shared(&t.cand, (size_t)cap_c * RW));
            FQD_TRY(shared(&t.cand_root, (size_t)cap_c));
            FQD_TRY(shared(&t.bitmap, (size_t)bm_words));
            FQD_TRY(arena(ctx, (size_t)nparts0, &t.cursor0));
            FQD_TRY(arena(ctx, (size_t)nparts0 + G, &t.cnt0));
            FQD_TRY(arena(ctx, (size_t)std::max(nparts0, npartsN) + G, &t.a2a_send));
            FQD_CUDA(cudaEventCreateWithFlags(&t.ev_fork, cudaEventDisableTiming));
            FQD_CUDA(cudaEventCreateWithFlags(&t.ev_join, cudaEventDisableTiming));
            if (emit_next) {
                FQD_TRY(arena(ctx, (size_t)npartsN, &t.cursorN));
                FQD_TRY(arena(ctx, (size_t)npartsN + G, &t.cntN));
            }
            FQD_TRY(arena(ctx, (size_t)CTR_WORDS, &t.ctrs));
            FQD_TRY(arena(ctx, (size_t)nper0, &t.oversize));
            FQD_TRY(arena(ctx, (size_t)4 + 2 * G, &t.gath_in));   // [0..3] this rank's numbers, [4..] edge counts at the fork
            FQD_TRY(arena(ctx, (size_t)4 * G * 2, &t.gath));
            FQD_TRY(arena(ctx, (size_t)STATUS_WORDS * 2, &t.stat_in));
            FQD_TRY(arena(ctx, (size_t)STATUS_WORDS * 2 * G, &t.stat_all));
            FQD_TRY(arena(ctx, (size_t)cap_u * KW, &sh.local.ukey));
            FQD_TRY(arena(ctx, (size_t)cap_u, &sh.local.ucount));
            FQD_TRY(arena(ctx, (size_t)cap_u, &sh.local.ufirst));
            FQD_CUDA(cudaMemsetAsync(t.cursor0, 0, (size_t)nparts0 * 4, s));
            if (emit_next) FQD_CUDA(cudaMemsetAsync(t.cursorN, 0, (size_t)npartsN * 4, s));
            FQD_CUDA(cudaMemsetAsync(t.ctrs, 0, CTR_WORDS * 4, s));
            // the slab estimate was too small: this rank's (zeroed) counters keep the peers' tile kernels harmless
            // until the ranks have agreed to hand the job over
            return inside ? FQD_OK : RC_FALLBACK_REPLICATED;
        }();
    }

    // =====================================================================================================
    // phase 2: every rank partitions its records into its own tile regions
    // =====================================================================================================
    for (int i = 0; i < L; i++) {
        Shard &sh = S[i];
        TileRank &t = T[i];
        if (t.rc != FQD_OK) continue;
        t.rc = [&]() -> int {
            FQD_CUDA(cudaSetDevice(sh.ctx->device));
            const DeviceJob &job = sh.job;
            IngestParams ip{};
            ip.n = job.n;
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
            ip.index_base = sh.index_base;
            ip.sharded = 0;      // (a filtered record travels with weight 0 in partition mode anyway)
            ip.ctr = sh.ctx->d_ctr;
            ip.codec = codec;
            const PartParams part{t.tiles0, t.cursor0, nparts0, t.spill, t.ctrs + CTR_SPILL, spill_cap, region};
            if (job.n) FQD_TRY((launch_partition<K, PW>(sh.ctx, job, codec, ip, part, fused ? (uint32_t)d + 1u : 0u, sh.index_base, sh.tt)));
            return FQD_OK;
        }();
    }
    lap("partition");

    // =====================================================================================================
    // phase 3: fill counters to the tile owners; phase 4: dedupe (+ pass 0, + tiles of pass 1) on the owners
    // =====================================================================================================
    FQD_TRY(exchange_counters([&](int i) { return T[i].cursor0; }, [&](int i) { return T[i].cnt0; }, nper0));
    lap("counters a2a");
    for (int i = 0; i < L; i++) {
        Shard &sh = S[i];
        TileRank &t = T[i];
        if (t.rc != FQD_OK) continue;
        t.rc = [&]() -> int {
            FQD_CUDA(cudaSetDevice(sh.ctx->device));
            cudaStream_t s = sh.ctx->stream;
            const int r = rank_of(i);
            FQD_CUDA(cudaMemsetAsync(t.bitmap, 0, (size_t)bm_words * 4, s));   // (every rank has entered this job by now)
            TileSource src{};
            for (int g = 0; g < G; g++) src.buf[g] = reinterpret_cast<const uint32_t *>(peer_ptr(sh, g, t.tiles0));
            src.cnt = t.cnt0; src.cnt_stride = nper0 + 1; src.G = (uint32_t)G; src.self = (uint32_t)r; src.first_tile = (uint32_t)r * nper0; src.ntiles = nper0;
            src.region = region; src.peer_ldg = peer_ldg;
            DedupeOut out{sh.local.ukey, sh.local.ucount, sh.local.ufirst, t.ctrs + CTR_UNIQUE, t.oversize, t.ctrs + CTR_OVERSIZE, 0,
                          cap_u, t.ctrs + CTR_UNIQUE_OVER};
            PassParams p0{};
            EdgeSink sink0{};
            if (fused) {
                p0.d = d; p0.edit = 0; p0.varlen = sh.job.varlen ? 1 : 0; p0.method = method;
                p0.max_len = sh.job.max_len; p0.pad_code = codec.pad_code; p0.V = 1; p0.world = 1;
                p0.pass_j = 0;
                p0.fix_st = 0;
                p0.fix_bl = block_start(sh.job.max_len, 1u, (uint32_t)d + 1u);
                p0.ctr = sh.ctx->d_ctr;
                p0.edge_flags = 1;
                sink0 = EdgeSink{t.edges, t.ctrs + CTR_EDGES, cap_e, t.ctrs + CTR_EDGE_OVER};
                NextPass nx{};
                if (emit_next) {
                    nx.next = PartParams{t.tilesN, t.cursorN, npartsN, nullptr, nullptr, 0, region};
                    nx.pass_j = 1;
                    nx.st = block_start(sh.job.max_len, 1u, (uint32_t)d + 1u);
                    nx.bl = block_start(sh.job.max_len, 2u, (uint32_t)d + 1u) - nx.st;
                }
                dedupe_tile_kernel<K, PW, true><<<nper0, TILE_THREADS, 0, s>>>(src, out, p0, sink0, nx);
            } else {
                dedupe_tile_kernel<K, PW, false><<<nper0, TILE_THREADS, 0, s>>>(src, out, p0, sink0, NextPass{});
            }
            sh.tt.launches++;
            FQD_CUDA(cudaGetLastError());
            if (i == 0) lap("dedupe tiles");
            return FQD_OK;
        }();
    }

    // ---- agreement 1: status, input errors, unique counts, skew ----
    {
        const int rc = agree(no_tail);
        if (rc != FQD_OK) return rc;
    }
    uint64_t n_disc = 0, n_seq = 0, phred = ~0ull, skew = 0, any_over = 0;
    bool any_unknown = false;
    std::vector<uint32_t> n_spill(G), n_over_of(G), U_tiles_of(G);
    for (int g = 0; g < G; g++) {
        const uint64_t *v = &all[(size_t)g * STATUS_WORDS];
        U_tiles_of[g] = (uint32_t)v[1];
        n_over_of[g] = (uint32_t)v[2];
        n_spill[g] = (uint32_t)v[3];
        any_over += v[2] + v[3];
        skew |= v[4];
        phred = std::min(phred, v[5]);
        n_disc += v[6];
        n_seq += j0.weights ? v[7] : (W.base[g + 1] - W.base[g]) - v[6];
        for (int k = 0; k < 8; k++) { unknown_out[k] |= (uint32_t)v[8 + k]; any_unknown |= v[8 + k] != 0; }
    }
    if (phred != ~0ull) {
        for (auto &sh : S) { sh.st->bad_record = phred >> 8; sh.st->bad_char = (uint32_t)(phred & 0xFF); }
        set_error("Character %c outside of valid phred range ('%c' to '%c')", (int)(phred & 0xFF), (int)j0.phred_offset, 126);
        return FQD_ERR_PHRED;
    }
    for (auto &sh : S) sh.st->bad_record = ~0ull;
    if (any_unknown) return RC_RETRY_ALPHABET;
    if (skew) {
        if (trace) fprintf(stderr, "[fqd trace] tiles: a spill / unique / edge buffer overflowed -> replicated-set plan\n");
        return RC_FALLBACK_REPLICATED;
    }
    lap("agree");
    // From here on nothing waits for the host until the final agreement: where a count is still being produced on the
    // device (the spill path below adds uniques) launches are sized by an upper bound every rank can compute from
    // the agreed numbers, and the kernels read the count itself.
    uint64_t spill_total = 0;
    for (int g = 0; g < G; g++) spill_total += n_spill[g];
    auto spill_records_of = [&](int g) { return (uint64_t)n_over_of[g] * G * region + spill_total; };   // what rank g's spill path may see
    uint32_t U_max = 0;
    uint64_t U_bound_total = 0;
    std::vector<uint32_t> U_bound_of(G);
    for (int g = 0; g < G; g++) {
        U_bound_of[g] = (uint32_t)std::min<uint64_t>(cap_u, U_tiles_of[g] + (any_over ? spill_records_of(g) : 0));
        U_max = std::max(U_max, U_bound_of[g]);
        U_bound_total += U_bound_of[g];
    }
    for (int i = 0; i < L; i++) { T[i].U_tiles = U_tiles_of[rank_of(i)]; T[i].U = U_bound_of[rank_of(i)]; }

    // =====================================================================================================
    // phase 5: oversize tiles (a key family larger than a tile) through the single-table insert on their owner
    // =====================================================================================================
    if (any_over) {
        // (scratch of this phase has rank-dependent sizes: released afterwards so that the buffers allocated later
        // sit at the same arena offsets on every rank again)
        std::vector<size_t> marks(L);
        for (int i = 0; i < L; i++) marks[i] = arena_mark(S[i].ctx);
        for (int i = 0; i < L; i++) {
            Shard &sh = S[i];
            TileRank &t = T[i];
            t.rc = [&]() -> int {
                FQD_CUDA(cudaSetDevice(sh.ctx->device));
                cudaStream_t s = sh.ctx->stream;
                const int r = rank_of(i);
                const uint32_t n_over = n_over_of[r];
                const uint64_t n_rec = spill_records_of(r);
                if (n_rec == 0) return FQD_OK;
                const uint64_t capacity = n_rec + (n_rec >> 1) + 1024;
                uint32_t *table, *uslot;
                FQD_TRY(arena(sh.ctx, capacity * RW, &table));
                FQD_TRY(arena(sh.ctx, std::max<uint64_t>(n_rec, 1), &uslot));
                FQD_CUDA(cudaMemsetAsync(table, 0xFF, capacity * RW * 4, s));
                const TableRef tr{table, capacity, uslot, sh.ctx->d_ctr};
                TileSource src{};
                for (int g = 0; g < G; g++) src.buf[g] = reinterpret_cast<const uint32_t *>(peer_ptr(sh, g, t.tiles0));
                src.cnt = t.cnt0; src.cnt_stride = nper0 + 1; src.G = (uint32_t)G; src.self = (uint32_t)r; src.first_tile = (uint32_t)r * nper0; src.ntiles = nper0;
                src.region = region; src.peer_ldg = peer_ldg;
                if (n_over)
                    spill_insert_tiles_kernel<K, PW><<<n_over * (uint32_t)G * (TILE_R / 256), 256, 0, s>>>(src, t.oversize, tr, t.ctrs + CTR_CLAIMED);
                for (int g = 0; g < G; g++)
                    if (n_spill[g])
                        spill_insert_owned_kernel<K, PW><<<cdiv(n_spill[g], 256), 256, 0, s>>>(
                            reinterpret_cast<const uint32_t *>(peer_ptr(sh, g, t.spill)), n_spill[g], fused ? (uint32_t)d + 1u : 0u, (uint32_t)d + 1u,
                            nparts0, (uint32_t)r * nper0, nper0, sh.job.max_len, codec.pad_code, sh.job.varlen ? 1 : 0, tr, t.ctrs + CTR_CLAIMED);
                // (the number of claimed slots stays on the device: the gather is sized by its upper bound)
                gather_nonzero_kernel<K, PW><<<cdiv(n_rec, 256), 256, 0, s>>>((uint32_t)std::min<uint64_t>(n_rec, 0xFFFFFFF0u), table, uslot,
                                                                                   sh.local.ukey, sh.local.ucount, sh.local.ufirst,
                                                                                   t.ctrs + CTR_UNIQUE, 0, cap_u, t.ctrs + CTR_UNIQUE_OVER,
                                                                                   t.ctrs + CTR_CLAIMED);
                FQD_CUDA(cudaGetLastError());
                sh.tt.launches += 2 + G;
                // the uniques of oversize tiles were not compared inside a tile, and they are not in the tiles of pass 1
                // yet: pass 0 among themselves (their bucket mates took the same path, on this rank), then emit them
                if (fused || emit_next) {
                    PassParams sp{};
                    sp.U = t.U; sp.U_dev = t.ctrs + CTR_UNIQUE; sp.u_lo = t.U_tiles; sp.ukey = sh.local.ukey; sp.ucount = sh.local.ucount;
                    sp.d = d; sp.edit = 0; sp.varlen = sh.job.varlen ? 1 : 0; sp.method = method;
                    sp.max_len = sh.job.max_len; sp.pad_code = codec.pad_code; sp.V = 1; sp.world = 1;
                    sp.ctr = sh.ctx->d_ctr;
                    sp.edge_flags = 1; sp.id_mul = (uint32_t)G; sp.id_add = (uint32_t)r;
                    for (int k = 0; k < 256; k++) sp.rank_of_code[k] = codec.rank[k];
                    const uint32_t ns = t.U - t.U_tiles;   // (upper bound)
                    if (fused && ns) {
                        const uint32_t np = tile_partitions(ns);
                        uint32_t *sb, *sc;
                        FQD_TRY(arena(sh.ctx, (size_t)np * TILE_R * RW, &sb));
                        FQD_TRY(arena(sh.ctx, (size_t)np, &sc));
                        FQD_CUDA(cudaMemsetAsync(sc, 0, (size_t)np * 4, s));
                        sp.pass_j = 0;
                        sp.fix_st = 0;
                        sp.fix_bl = block_start(sh.job.max_len, 1u, (uint32_t)d + 1u);
                        const PartParams qp{sb, sc, np, nullptr, nullptr, 0};
                        bucket_partition_kernel<K, PW><<<cdiv(ns, 256 * BP_ROWS), 256, 0, s>>>(sp, qp);
                        const EdgeSink sink{t.edges, t.ctrs + CTR_EDGES, cap_e, t.ctrs + CTR_EDGE_OVER};
                        bucket_tile_kernel<K, PW><<<np, TILE_THREADS, 0, s>>>(single_source(qp), sp, sink);
                        sh.tt.launches += 2;
                    }
                    if (emit_next && ns) {
                        sp.pass_j = 1;
                        sp.fix_st = block_start(sh.job.max_len, 1u, (uint32_t)d + 1u);
                        sp.fix_bl = block_start(sh.job.max_len, 2u, (uint32_t)d + 1u) - sp.fix_st;
                        const PartParams qn{t.tilesN, t.cursorN, npartsN, nullptr, nullptr, 0, region};
                        bucket_partition_kernel<K, PW><<<cdiv(ns, 256 * BP_ROWS), 256, 0, s>>>(sp, qn);
                        sh.tt.launches++;
                    }
                    FQD_CUDA(cudaGetLastError());
                }
                return FQD_OK;
            }();
        }
        for (int i = 0; i < L; i++) arena_release(S[i].ctx, marks[i]);
        lap("oversize tiles");
    }
    const uint64_t U_total_bound = U_bound_total;
    const uint64_t ids64 = (uint64_t)U_max * G;          // job-wide id space; slots are rank-major with stride U_max
    if (ids64 >= EDGE_ID) { set_error("too many unique keys for the tile-sharded plan (%llu)", (unsigned long long)U_total_bound); return FQD_ERR_UNSUPPORTED; }
    const uint32_t n_ids = (uint32_t)std::max<uint64_t>(ids64, 1);

    // =====================================================================================================
    // phase 6: forests and flags over the job-wide id space (every rank holds all of them)
    // =====================================================================================================
    const int npass_all = (d > 0 && U_total_bound > 1) ? d + 1 : 0;
    const int first_pass = fused ? 1 : 0;
    const bool use_emitted = emit_next && U_total_bound <= guessU;
    const bool need_pass_buffers = npass_all > first_pass + (use_emitted ? 1 : 0);
    const uint32_t nperP = cdiv(tile_partitions(std::max<uint64_t>(U_total_bound, 1)), (uint32_t)G), npartsP = nperP * (uint32_t)G;
    for (int i = 0; i < L; i++) {
        Shard &sh = S[i];
        TileRank &t = T[i];
        t.rc = [&]() -> int {
            FQD_CUDA(cudaSetDevice(sh.ctx->device));
            cudaStream_t s = sh.ctx->stream;
            bool inside = true;
            if (need_pass_buffers) {
                for (int b = 0; b < 2; b++) {
                    FQD_TRY(arena(sh.ctx, (size_t)npartsP * region * RW, &t.tilesP[b]));
                    inside = inside && in_slab(sh, t.tilesP[b], (size_t)npartsP * region * RW * 4);
                }
                FQD_TRY(arena(sh.ctx, (size_t)npartsP, &t.cursorP));
                FQD_TRY(arena(sh.ctx, (size_t)npartsP + G, &t.cntP));
                if (npartsP > std::max(nparts0, npartsN)) FQD_TRY(arena(sh.ctx, (size_t)npartsP + G, &t.a2a_send));
            }
            if (!inside) return RC_FALLBACK_REPLICATED;
            Forest &f = sh.f;
            FQD_TRY(arena(sh.ctx, (size_t)n_ids, &f.parent_full));
            FQD_TRY(arena(sh.ctx, (size_t)std::max<uint32_t>(t.U, 1), &f.selected));
            FQD_TRY(arena(sh.ctx, (size_t)std::max<uint32_t>(t.U, 1), &t.root_of));
            FQD_TRY(arena(sh.ctx, (size_t)std::max<uint32_t>(t.U, 1), &t.loc_of));
            const size_t own = std::max<uint32_t>(t.U, 1);   // flags: this rank's own keys only
            if (method == METHOD_DIRECTIONAL) {
                FQD_TRY(arena(sh.ctx, (size_t)n_ids, &f.parent_one));
                FQD_TRY(arena(sh.ctx, own, &f.dominated));
                FQD_TRY(arena(sh.ctx, own, &f.dead));
                FQD_TRY(arena(sh.ctx, (size_t)n_ids, &f.deadroot));
                FQD_CUDA(cudaMemsetAsync(f.dominated, 0, own, s));
                FQD_CUDA(cudaMemsetAsync(f.dead, 0, own, s));
                FQD_CUDA(cudaMemsetAsync(f.deadroot, 0, n_ids, s));
            }
            if (method != METHOD_ADJACENCY) {
                FQD_TRY(arena(sh.ctx, (size_t)n_ids, &f.best));
                FQD_TRY(arena(sh.ctx, own, &t.linked));
                FQD_CUDA(cudaMemsetAsync(f.best, 0xFF, (size_t)n_ids * 4, s));
                FQD_CUDA(cudaMemsetAsync(t.linked, 0, own, s));
            }
            init_forest_kernel<<<cdiv(n_ids, 256), 256, 0, s>>>(n_ids, f.parent_full, f.parent_one, nullptr);
            sh.tt.launches++;
            FQD_CUDA(cudaGetLastError());
            FQD_CUDA(cudaMemsetAsync(&sh.ctx->d_ctr->n_merges, 0, 4, s));
            return FQD_OK;
        }();
    }

    // =====================================================================================================
    // phase 7: the remaining pigeonhole passes, tile by tile on the tile owners
    // =====================================================================================================
    // fused jobs: the edges of pass 0 can be applied while pass 1 runs.  Measured on 8 B200 (100 M reads): SLOWER -- the
    // pass stretches from 0.39 to 1.13 ms while the two edge applications together only drop from 1.33 to 0.88 ms --
    // so it is a switch (FQD_APPLY_OVERLAP=1), off by default.
    const bool overlap_apply = fused && npass_all > first_pass && getenv("FQD_APPLY_OVERLAP");
    std::vector<const uint32_t *> early_first(L, nullptr);
    uint32_t early_stride = 0;
    for (int j = first_pass; j < npass_all; j++) {
        const bool emitted = use_emitted && j == 1;
        const uint32_t nper = emitted ? nperN : nperP, nparts = emitted ? npartsN : npartsP;
        auto pass_params = [&](int i) {
            Shard &sh = S[i];
            PassParams pp{};
            pp.U = T[i].U; pp.U_dev = T[i].ctrs + CTR_UNIQUE; pp.u_lo = 0; pp.ukey = sh.local.ukey; pp.ucount = sh.local.ucount;
            pp.d = d; pp.edit = 0; pp.varlen = sh.job.varlen ? 1 : 0; pp.method = method;
            pp.max_len = sh.job.max_len; pp.pad_code = codec.pad_code; pp.V = 1; pp.world = 1;
            pp.ctr = sh.ctx->d_ctr;
            pp.edges = T[i].adj; pp.edge_cap = cap_adj;
            pp.edge_flags = 1; pp.id_mul = (uint32_t)G; pp.id_add = (uint32_t)rank_of(i);
            pp.pass_j = j;
            pp.fix_st = block_start(sh.job.max_len, (uint32_t)j, (uint32_t)d + 1u);
            pp.fix_bl = block_start(sh.job.max_len, (uint32_t)j + 1u, (uint32_t)d + 1u) - pp.fix_st;
            for (int k = 0; k < 256; k++) pp.rank_of_code[k] = codec.rank[k];
            return pp;
        };
        if (!emitted) {
            for (int i = 0; i < L; i++) {
                if (T[i].rc != FQD_OK) continue;
                T[i].rc = [&]() -> int {
                    FQD_CUDA(cudaSetDevice(S[i].ctx->device));
                    cudaStream_t s = S[i].ctx->stream;
                    FQD_CUDA(cudaMemsetAsync(T[i].cursorP, 0, (size_t)npartsP * 4, s));
                    const PartParams qp{T[i].tilesP[j & 1], T[i].cursorP, npartsP, nullptr, nullptr, 0, region};
                    if (T[i].U) bucket_partition_kernel<K, PW><<<cdiv(T[i].U, 256 * BP_ROWS), 256, 0, s>>>(pass_params(i), qp);
                    S[i].tt.launches++;
                    FQD_CUDA(cudaGetLastError());
                    return FQD_OK;
                }();
            }
        }
        FQD_TRY(exchange_counters([&](int i) { return emitted ? T[i].cursorN : T[i].cursorP; },
                                  [&](int i) { return emitted ? T[i].cntN : T[i].cntP; }, nper));
        for (int i = 0; i < L; i++) {
            if (T[i].rc != FQD_OK) continue;
            T[i].rc = [&]() -> int {
                Shard &sh = S[i];
                TileRank &t = T[i];
                FQD_CUDA(cudaSetDevice(sh.ctx->device));
                const int r = rank_of(i);
                if (j == first_pass && overlap_apply) {
                    // every rank's edges of the passes so far (pass 0, done inside the dedupe tiles) are complete, and the
                    // exchange just told how many there are: hook them on a second stream while this pass runs
                    const uint32_t *cnt = emitted ? t.cntN : t.cntP;
                    // (the counter block is reused by later passes: keep the G numbers)
                    copy_strided_kernel<<<1, 32, 0, sh.ctx->stream>>>(t.gath_in + 4, cnt + nper, nper + 1, (uint32_t)G);
                    FQD_CUDA(cudaGetLastError());
                    FQD_CUDA(cudaEventRecord(t.ev_fork, sh.ctx->stream));
                    FQD_CUDA(cudaStreamWaitEvent(sh.ctx->copy_stream, t.ev_fork, 0));
                    EdgeSource es{};
                    for (int g = 0; g < G; g++) { es.edges[g] = reinterpret_cast<const uint2 *>(peer_ptr(sh, g, t.edges)); es.cap[g] = cap_e; }
                    es.n_edges = t.gath_in + 4; es.n_stride = 1; es.G = (uint32_t)G; es.self = (uint32_t)r; es.id_stride = U_max;
                    EdgeFlags ef{sh.f.dominated, sh.f.dead, t.linked, method == METHOD_HIGHEST ? 1 : 0};
                    apply_edges_kernel<<<sh.ctx->sm_count * 4, 256, 0, sh.ctx->copy_stream>>>(es, sh.f.parent_full, sh.f.parent_one, ef,
                                                                                           sh.ctx->d_ctr);
                    FQD_CUDA(cudaGetLastError());
                    FQD_CUDA(cudaEventRecord(t.ev_join, sh.ctx->copy_stream));
                    sh.tt.launches++;
                    early_first[i] = t.gath_in + 4;
                    early_stride = 1;
                }
                TileSource src{};
                const uint32_t *mine = emitted ? t.tilesN : t.tilesP[j & 1];
                for (int g = 0; g < G; g++) src.buf[g] = reinterpret_cast<const uint32_t *>(peer_ptr(sh, g, mine));
                src.cnt = emitted ? t.cntN : t.cntP; src.cnt_stride = nper + 1;
                src.G = (uint32_t)G; src.self = (uint32_t)r; src.first_tile = (uint32_t)r * nper; src.ntiles = nper;
                src.region = region; src.peer_ldg = peer_ldg;
                const EdgeSink sink{t.edges, t.ctrs + CTR_EDGES, cap_e, t.ctrs + CTR_EDGE_OVER};
                bucket_tile_kernel<K, PW><<<nper, TILE_THREADS, 0, sh.ctx->stream>>>(src, pass_params(i), sink);
                sh.tt.launches++;
                FQD_CUDA(cudaGetLastError());
                return FQD_OK;
            }();
        }
        (void)nparts;
    }
    lap("forest init + passes");

    // =====================================================================================================
    // phase 8: all edges of all ranks into every rank's forests (edge lists read from the peers' HBM)
    // =====================================================================================================
    auto gather4 = [&](int slot) -> int {   // {edge count, candidate count, adjacency edge count lo, hi} of every rank
        for (int i = 0; i < L; i++) {
            if (T[i].rc != FQD_OK) continue;
            FQD_CUDA(cudaSetDevice(S[i].ctx->device));
            cudaStream_t s = S[i].ctx->stream;
            FQD_CUDA(cudaMemcpyAsync(T[i].gath_in, T[i].ctrs + CTR_EDGES, 4, cudaMemcpyDeviceToDevice, s));
            FQD_CUDA(cudaMemcpyAsync(T[i].gath_in + 1, T[i].ctrs + CTR_CAND, 4, cudaMemcpyDeviceToDevice, s));
            FQD_CUDA(cudaMemcpyAsync(T[i].gath_in + 2, &S[i].ctx->d_ctr->n_edges, 8, cudaMemcpyDeviceToDevice, s));
        }
        return allgather_u32([&](int i) { return T[i].gath_in; }, [&](int i) { return T[i].gath + (size_t)slot * 4 * G; }, 4);
    };
    FQD_TRY(gather4(0));
    lap("edge counts ag");
    for (int i = 0; i < L; i++) {
        if (T[i].rc != FQD_OK) continue;
        T[i].rc = [&]() -> int {
            Shard &sh = S[i];
            TileRank &t = T[i];
            FQD_CUDA(cudaSetDevice(sh.ctx->device));
            EdgeSource es{};
            for (int g = 0; g < G; g++) { es.edges[g] = reinterpret_cast<const uint2 *>(peer_ptr(sh, g, t.edges)); es.cap[g] = cap_e; }
            es.n_edges = t.gath; es.n_stride = 4; es.G = (uint32_t)G; es.self = (uint32_t)rank_of(i); es.id_stride = U_max;
            es.first = early_first[i]; es.first_stride = early_stride;   // (what the overlapped launch has taken care of)
            EdgeFlags ef{sh.f.dominated, sh.f.dead, t.linked, method == METHOD_HIGHEST ? 1 : 0};
            if (npass_all)
                apply_edges_kernel<<<sh.ctx->sm_count * 8, 256, 0, sh.ctx->stream>>>(es, sh.f.parent_full, sh.f.parent_one, ef, sh.ctx->d_ctr);
            if (early_first[i]) FQD_CUDA(cudaStreamWaitEvent(sh.ctx->stream, t.ev_join, 0));
            sh.tt.launches++;
            FQD_CUDA(cudaGetLastError());
            return FQD_OK;
        }();
    }
    lap("apply edges");

    // =====================================================================================================
    // phase 9: what a component's answer needs from several members
    // =====================================================================================================
    std::vector<uint8_t *> adj_state(L, nullptr);
    if (method == METHOD_ADJACENCY) {
        // (higher, lower) edges of all ranks, fetched once; then the rounds of the single-GPU plan over the id space
        // the counts are needed on the host (launch sizes of the rounds): one more agreement point
        const int rc = agree(no_tail);
        if (rc != FQD_OK) return rc;
        std::vector<uint32_t> off(G + 1, 0);
        for (int g = 0; g < G; g++) {
            const uint64_t ne = all[(size_t)g * STATUS_WORDS + 19];
            if (ne > cap_adj || all[(size_t)g * STATUS_WORDS + 4]) return RC_FALLBACK_REPLICATED;
            off[g + 1] = off[g] + (uint32_t)ne;
        }
        const uint32_t n_adj = off[G];
        for (int i = 0; i < L; i++) {
            T[i].rc = [&]() -> int {
                Shard &sh = S[i];
                TileRank &t = T[i];
                FQD_CUDA(cudaSetDevice(sh.ctx->device));
                cudaStream_t s = sh.ctx->stream;
                uint2 *alle;
                uint32_t *d_off;
                FQD_TRY(arena(sh.ctx, (size_t)std::max<uint32_t>(n_adj, 1), &alle));
                FQD_TRY(arena(sh.ctx, (size_t)G + 1, &d_off));
                FQD_CUDA(cudaMemcpyAsync(d_off, off.data(), (size_t)(G + 1) * 4, cudaMemcpyHostToDevice, s));
                EdgeSource es{};
                for (int g = 0; g < G; g++) { es.edges[g] = reinterpret_cast<const uint2 *>(peer_ptr(sh, g, t.adj)); es.cap[g] = (uint32_t)cap_adj; }
                es.n_edges = t.gath + 2; es.n_stride = 4; es.G = (uint32_t)G; es.self = (uint32_t)rank_of(i); es.id_stride = U_max;
                if (n_adj) adjacency_copy_kernel<<<sh.ctx->sm_count * 4, 256, 0, s>>>(es, alle, d_off);
                FQD_CUDA(cudaGetLastError());
                FQD_CUDA(cudaStreamSynchronize(s));   // `off` is host memory
                SelectParams sp{};
                sp.U = n_ids;
                sp.method = method; sp.ctr = sh.ctx->d_ctr;
                FQD_TRY(arena(sh.ctx, (size_t)n_ids, &sp.state));
                FQD_TRY(arena(sh.ctx, (size_t)n_ids, &sp.stamp));
                FQD_CUDA(cudaMemsetAsync(sp.state, 0, n_ids, s));
                FQD_CUDA(cudaMemsetAsync(sp.stamp, 0, (size_t)n_ids * 4, s));
                sp.edges = alle;
                sp.n_edges = n_adj;
                for (uint32_t round = 1;; round++) {
                    sp.round = round;
                    FQD_CUDA(cudaMemsetAsync(&sh.ctx->d_ctr->undecided, 0, 4, s));
                    if (sp.n_edges) { adj_edge_kernel<<<cdiv(sp.n_edges, 256), 256, 0, s>>>(sp); sh.tt.launches++; }
                    adj_node_kernel<<<cdiv(n_ids, 256), 256, 0, s>>>(sp);
                    sh.tt.launches++;
                    FQD_TRY(fetch_counters(sh.ctx));
                    if (sh.ctx->h_ctr->undecided == 0) break;
                    if (round > n_ids + 2) { set_error("internal: adjacency rounds did not converge"); return FQD_ERR_CUDA; }
                }
                adj_state[i] = sp.state;
                return FQD_OK;
            }();
        }
    } else {
        for (int i = 0; i < L; i++) {
            if (T[i].rc != FQD_OK) continue;
            T[i].rc = [&]() -> int {
                Shard &sh = S[i];
                TileRank &t = T[i];
                FQD_CUDA(cudaSetDevice(sh.ctx->device));
                cudaStream_t s = sh.ctx->stream;
                CandParams cp{};
                cp.U = t.U; cp.U_dev = t.ctrs + CTR_UNIQUE; cp.G = (uint32_t)G; cp.self = (uint32_t)rank_of(i); cp.id_stride = U_max;
                cp.ukey = sh.local.ukey; cp.ucount = sh.local.ucount;
                cp.forest = method == METHOD_DIRECTIONAL ? sh.f.parent_one : sh.f.parent_full;
                cp.dead = sh.f.dead; cp.linked = t.linked;
                cp.root_of = t.root_of; cp.loc_of = t.loc_of;
                cp.cand = t.cand; cp.cand_root = t.cand_root; cp.n_cand = t.ctrs + CTR_CAND; cp.cap = cap_c;
                cp.method = method;
                if (t.U) candidates_kernel<K, PW><<<cdiv(t.U, 256), 256, 0, s>>>(cp);
                sh.tt.launches++;
                FQD_CUDA(cudaGetLastError());
                return FQD_OK;
            }();
        }
        FQD_TRY(gather4(1));
        for (int i = 0; i < L; i++) {
            if (T[i].rc != FQD_OK) continue;
            T[i].rc = [&]() -> int {
                Shard &sh = S[i];
                TileRank &t = T[i];
                FQD_CUDA(cudaSetDevice(sh.ctx->device));
                BestParams bp{};
                for (int g = 0; g < G; g++) {
                    bp.cand[g] = reinterpret_cast<const uint32_t *>(peer_ptr(sh, g, t.cand));
                    bp.cand_root[g] = reinterpret_cast<const uint32_t *>(peer_ptr(sh, g, t.cand_root));
                    bp.cap[g] = cap_c;
                }
                bp.n_cand = t.gath + (size_t)4 * G + 1; bp.n_stride = 4; bp.G = (uint32_t)G;
                bp.best = sh.f.best;
                bp.deadroot = sh.f.deadroot;
                for (int k = 0; k < 256; k++) bp.rank_of_code[k] = codec.rank[k];
                best_candidate_kernel<K, PW><<<sh.ctx->sm_count * 4, 256, 0, sh.ctx->stream>>>(bp);
                sh.tt.launches++;
                FQD_CUDA(cudaGetLastError());
                return FQD_OK;
            }();
        }
    }
    lap("candidates");

    // =====================================================================================================
    // phase 10: the keys of a rank decide; keep bits go to the rank that holds the record
    // =====================================================================================================
    for (int i = 0; i < L; i++) {
        if (T[i].rc != FQD_OK) continue;
        T[i].rc = [&]() -> int {
            Shard &sh = S[i];
            TileRank &t = T[i];
            FQD_CUDA(cudaSetDevice(sh.ctx->device));
            cudaStream_t s = sh.ctx->stream;
            SelectOwnParams sp{};
            sp.U = t.U; sp.U_dev = t.ctrs + CTR_UNIQUE; sp.G = (uint32_t)G; sp.self = (uint32_t)rank_of(i); sp.id_stride = U_max;
            sp.ucount = sh.local.ucount; sp.ufirst = sh.local.ufirst;
            sp.root_of = t.root_of; sp.loc_of = t.loc_of; sp.best = sh.f.best;
            sp.dominated = sh.f.dominated; sp.dead = sh.f.dead; sp.linked = t.linked; sp.deadroot = sh.f.deadroot;
            sp.state = adj_state[i];
            sp.selected = sh.f.selected;
            sp.bitmap = t.bitmap; sp.bit_base = (uint32_t)W.base[0];
            sp.method = method; sp.ctr = sh.ctx->d_ctr;
            FQD_CUDA(cudaMemsetAsync(&sh.ctx->d_ctr->n_selected, 0, 4, s));
            if (t.U) select_own_kernel<<<cdiv(t.U, 256), 256, 0, s>>>(sp);
            sh.tt.launches++;
            FQD_CUDA(cudaGetLastError());
            return FQD_OK;
        }();
    }

    // ---- final agreement (also the barrier behind the remote keep bits): totals and late overflows; the keep bitmap is
    //      handed to the caller behind the exchange, ahead of the one host synchronisation ----
    {
        const int rc = agree([&](int i) -> int {
            // behind the exchange every rank's bitmap is complete: OR the peers' slices of this rank's records together
            if (S[i].job.bitmap && S[i].job.n) {
                MergeParams mp{};
                for (int g = 0; g < G; g++) mp.src[g] = reinterpret_cast<const uint32_t *>(peer_ptr(S[i], g, T[i].bitmap));
                mp.G = (uint32_t)G;
                mp.bit_lo = (uint32_t)(W.base[rank_of(i)] - W.base[0]);
                mp.n = (uint32_t)S[i].job.n;
                mp.src_words = bm_words;
                mp.dst = S[i].job.bitmap;
                merge_bitmaps_kernel<<<cdiv(cdiv(S[i].job.n, 32), 256), 256, 0, S[i].ctx->stream>>>(mp);
                FQD_CUDA(cudaGetLastError());
                S[i].tt.launches++;
            }
            FQD_CUDA(cudaEventRecord(T[i].e1, S[i].ctx->stream));
            return FQD_OK;
        });
        if (rc != FQD_OK) return rc;
        // nobody may reuse its arena (the next job, of any kind) while a peer still merges: one more rendezvous, on the
        // streams only
        FQD_TRY(allgather_u32([&](int i) { return T[i].gath_in; }, [&](int i) { return T[i].gath; }, 1));
    }
    uint64_t n_sel = 0, n_cand_pairs = 0, merges = 0, late = 0, U_total = 0;
    for (int g = 0; g < G; g++) {
        const uint64_t *v = &all[(size_t)g * STATUS_WORDS];
        U_total += v[1];
        late |= v[4] | v[20];
        n_sel += v[16];
        n_cand_pairs += v[17];
        merges = v[18];          // every rank applied every edge: the same number everywhere
    }
    if (late) {
        if (trace) fprintf(stderr, "[fqd trace] tiles: a pass tile / the unique, edge or candidate list overflowed -> replicated-set plan\n");
        return RC_FALLBACK_REPLICATED;
    }
    lap("select + agree");
    lap_report();
    for (int i = 0; i < L; i++) {
        Shard &sh = S[i];
        TileRank &t = T[i];
        FQD_CUDA(cudaSetDevice(sh.ctx->device));
        t.U = (uint32_t)all[(size_t)rank_of(i) * STATUS_WORDS + 1];   // the count itself (so far an upper bound)
        fqd_cluster_stats *st = sh.st;
        st->total_records = N;
        st->discarded_records = n_disc;
        st->number_of_sequences = n_seq;
        st->number_of_uniques = U_total;
        st->number_of_clusters = U_total - merges;
        st->number_selected = n_sel;
        st->candidate_pairs = n_cand_pairs;
        st->n_passes = npass_all;
        st->own_uniques = t.U;
        cudaEventElapsedTime(&st->ms_total, t.e0, t.e1);
        st->launches = sh.tt.launches;
        st->plan_flags = FQD_PLAN_DEDUPE_PARTITIONED | FQD_PLAN_PASSES_PARTITIONED | FQD_PLAN_SHARD_TILES |
                         (fused ? FQD_PLAN_PASS0_FUSED : 0u) | (use_emitted && npass_all > 1 ? FQD_PLAN_PASS1_TILES_EMITTED : 0u);
        sh.local.U = t.U;
        publish_result(sh.ctx, sh.local, sh.f, N, n_sel);
        sh.ctx->res.id_mul = 1;                                    // slot of own unique u: rank * stride + u
        sh.ctx->res.id_add = (uint32_t)rank_of(i) * U_max;
        sh.ctx->res.roots_only = true;
    }
    return FQD_OK;
    }
}
