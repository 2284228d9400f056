// partitioned.cuh -- the streaming plan for large jobs (included at the end of pipeline.cuh).
//
// Random probes into an HBM-sized hash table (the single-table plan) pay one 128-byte DRAM line
// per record and sit at ~30 % of HBM bandwidth; a counting sort of the bucket entries pays
// random 32-byte scatter writes.  For large jobs both stages are therefore rebuilt around ONE
// idea: PARTITION the 32-byte records by the top hash bits so finely (TILE_R = 512 records per
// region, ~60 % full) that one partition is one SHARED-MEMORY TILE, then give every partition
// to one thread block that stages it with coalesced streaming loads and does all the random
// work -- hash probes, key compares, counting -- inside shared memory.  HBM only ever sees
// streaming reads and writes; there is no table in HBM or L2 at all and no inter-block
// synchronisation.
//
//   ingest_kernel (partition mode)  records -> nparts regions of {key, weight, record index}
//                          (quality filter + packing + hashing fused; one atomic on the
//                          partition cursor + one 256-bit store per record; replaces
//                          average_error_rate + Trie.add_sequence, reference
//                          _fastqmodule.c:38-76, _triemodule.c:222-288)
//   dedupe_tile_kernel     one partition: open-addressing table of RECORD INDICES in shared
//                          memory (the keys stay in the staged tile, so a slot is claimed with
//                          one 32-bit CAS and is readable at once -- no lock, no fence); the
//                          winners accumulate count and smallest index, then leave as the
//                          dense unique arrays
//   bucket_partition_kernel  uniques -> regions of {key, count, uid} by the hash of pigeonhole
//                          block j
//   bucket_tile_kernel     one partition: the same index table used as a MULTIMAP keyed by the
//                          block hash: the inserting thread walks the probe sequence and tests
//                          every entry it passes with XOR+POPC Hamming, so bucket build and
//                          compare are one pass over a tile (replaces TrieNode_FindNearest /
//                          pop_cluster, _triemodule.c:380-495, :778-897).  Hits set the
//                          directional flags at once; the union-find hooks (random HBM accesses)
//                          are deferred to apply_edges_kernel, which runs them at full occupancy.
//
// Skew: a partition that outgrows its region ("oversize": one key with thousands of copies)
// sends the surplus to a spill buffer; oversize partitions and the spill are deduplicated by
// the single-table kernels into the same dense arrays.  A Hamming pass whose buckets outgrow a
// tile (few distinct block values) is redone by the counting-sort plan.
#pragma once

namespace fqd {

constexpr int TILE_T = 2 * TILE_R;         // table entries per tile (load <= 0.5)
constexpr int TILE_THREADS = 128;
constexpr int TILE_E = 256;                // edges a tile buffers in shared memory before hooking inline
static_assert(TILE_R <= 1024, "tile-local indices are packed into 10 bits");
constexpr uint32_t TILE_EMPTY = 0xFFFFFFFFu;
// An edge is (x, y | state): ids use the low 29 bits of y, the state says what the edge means to the dissection
// (closed forms).  One GPU writes the directional flag bytes at once and only ever stores
// NONE / ONE; a sharded job cannot (x and y belong to other ranks), so there every consequence of an edge
// travels in its state and is applied where the edge lists of all ranks are merged (apply_edges_kernel).
constexpr uint32_t EDGE_SHIFT = 29;
constexpr uint32_t EDGE_ID = (1u << EDGE_SHIFT) - 1u;
constexpr uint32_t EDGE_NONE = 0u << EDGE_SHIFT;
constexpr uint32_t EDGE_ONE = 1u << EDGE_SHIFT;      // both ends have count 1: the edge also joins the count-1 forest
constexpr uint32_t EDGE_DEAD_X = 2u << EDGE_SHIFT;   // x has count 1 and touches a key with count >= 2
constexpr uint32_t EDGE_DEAD_Y = 3u << EDGE_SHIFT;
constexpr uint32_t EDGE_DOM_X = 4u << EDGE_SHIFT;    // x (count >= 2) has a neighbour with count >= 2 * count_x - 1
constexpr uint32_t EDGE_DOM_Y = 5u << EDGE_SHIFT;
constexpr uint32_t EDGE_STATE = 7u << EDGE_SHIFT;

// Where the records of a tile come from.  One GPU: one partition buffer, one fill counter per tile.  A job
// sharded over G ranks: every rank partitions ITS records into its own buffer (all tiles, each ~1/G full);
// tile t belongs to rank t / ntiles_per_rank, and the owner's tile kernel fetches the G fragments of the tile
// straight from the peers' HBM (NVLink peer memory) with one bulk copy each -- the all-to-all of the records IS
// the staging of the tile, there is no exchange buffer and no second partition pass.
constexpr int MAX_RANKS = 16;
struct TileSource {
    const uint32_t *buf[MAX_RANKS];   // partition buffers (peer-mapped for the other ranks), regions indexed by GLOBAL tile
    const uint32_t *cnt;              // fill counters of the owned tiles gathered from all ranks: cnt[g * cnt_stride + j]
    uint32_t cnt_stride;
    uint32_t G, self;
    uint32_t first_tile;              // global index of owned tile 0
    uint32_t ntiles;                  // owned tiles
    uint32_t region;                  // records per region of every rank's buffer (PartParams::region)
    int peer_ldg;                     // measurement switch: fetch peer fragments with 16-byte loads instead of bulk copies
    // Split launches (one GPU): tiles are handled by two instances of a tile kernel, one whose shared-memory tile
    // holds TILE_R_SMALL records (more resident blocks per SM: the tile kernels are latency-bound) for the tiles
    // that fit it, one with full-size tiles for the rest.  A launch handles the tiles with cnt_lo < fill <= cnt_hi
    // (cnt_hi == 0: no window); oversize tiles belong to the launch whose window reaches TILE_R.  `list` (with its
    // device-side length) names the tiles of a launch explicitly: blocks then loop over it.
    uint32_t cnt_lo = 0, cnt_hi = 0;
    const uint32_t *list = nullptr, *n_list = nullptr;
};
constexpr int TILE_R_SMALL = 384;    // mean fill 307 +- ~50: nine tiles in ten fit

__device__ __forceinline__ void tile_load_rec(const uint32_t *recs, uint32_t i, uint32_t (&e)[PART_RW])
{
    const uint4 *r4 = reinterpret_cast<const uint4 *>(recs + (size_t)i * PART_RW);
    const uint4 a = r4[0], b = r4[1];
    e[0] = a.x; e[1] = a.y; e[2] = a.z; e[3] = a.w; e[4] = b.x; e[5] = b.y; e[6] = b.z; e[7] = b.w;
}

// ONE bulk asynchronous copy (TMA engine: cp.async.bulk global -> shared, completion on an mbarrier,
// evict-first in the L2: the data is read exactly once) of `bytes` bytes (16-byte aligned, a multiple of
// 16).  Called by every thread of the block; `between` runs while the copy is in flight; returns with
// the data (and whatever `between` wrote to shared memory) visible to the whole block.
template <typename Between>
__device__ __forceinline__ void bulk_load_to_shared(void *dst, const void *src, uint32_t bytes, Between between)
{
    __shared__ __align__(8) uint64_t mbar;
    const uint32_t mbar_a = (uint32_t)__cvta_generic_to_shared(&mbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar_a), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t policy;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_a), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                     ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"(mbar_a), "l"(policy)
                     : "memory");
    }
    between();
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(mbar_a), "r"(0) : "memory");
    } while (!done);
    __syncthreads();
}
__device__ __forceinline__ void bulk_load_to_shared(void *dst, const void *src, uint32_t bytes)
{
    bulk_load_to_shared(dst, src, bytes, [] {});
}

// Stage owned tile j: its G fragments (one bulk copy each, issued by the first G lanes, all completing on one
// mbarrier) land back to back in `recs` while all threads clear `tab_n` table entries.  Returns the number of
// records staged; 0 = empty tile, > TILE_R = oversize tile (nothing was staged).
__device__ __forceinline__ uint32_t stage_tile(uint32_t *recs, uint32_t *tab, uint32_t tab_n, const TileSource &S, uint32_t j)
{
    __shared__ __align__(8) uint64_t mbar;
    const uint32_t tid = threadIdx.x;
    uint32_t total = 0, mine = 0, before = 0, local = 0;
    bool over = false;
    for (uint32_t g = 0; g < S.G; g++) {
        const uint32_t c = S.cnt[(size_t)g * S.cnt_stride + j];
        if (g == S.self) local = c;
        over |= c > S.region;               // the rank's own region overflowed (the surplus is in its spill buffer)
        if (g < tid) before += c;
        if (g == tid) mine = c;
        total += c;
    }
    if (total == 0) return 0;
    const uint32_t hi = S.cnt_hi ? S.cnt_hi : (uint32_t)TILE_R;
    if (over || total > (uint32_t)TILE_R) return hi >= (uint32_t)TILE_R ? TILE_R + 1 : 0;
    if (total <= S.cnt_lo || total > hi) return 0;   // another launch's tile
    const uint32_t mbar_a = (uint32_t)__cvta_generic_to_shared(&mbar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar_a), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t bulk_bytes = (S.peer_ldg ? local : total) * (PART_RW * 4u);
    if (tid == 0 && bulk_bytes)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_a), "r"(bulk_bytes) : "memory");
    if (tid < S.G && mine && !(S.peer_ldg && tid != S.self)) {
        const uint32_t *src = S.buf[tid] + (size_t)(S.first_tile + j) * S.region * PART_RW;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(recs + (size_t)before * PART_RW);
        if (tid == S.self) {   // local HBM: read exactly once, evict-first in the L2
            uint64_t policy;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                         ::"r"(dst), "l"(src), "r"(mine * (PART_RW * 4u)), "r"(mbar_a), "l"(policy) : "memory");
        } else {               // a peer's HBM over NVLink
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(src), "r"(mine * (PART_RW * 4u)), "r"(mbar_a) : "memory");
        }
    }
    for (uint32_t i = tid; i < tab_n; i += TILE_THREADS) tab[i] = TILE_EMPTY;
    if (S.peer_ldg) {
        // the peers' fragments by plain 16-byte loads, four in flight per thread
        uint32_t off = 0;
        for (uint32_t g = 0; g < S.G; g++) {
            const uint32_t c = S.cnt[(size_t)g * S.cnt_stride + j];
            if (g != S.self && c) {
                const uint4 *src = reinterpret_cast<const uint4 *>(S.buf[g] + (size_t)(S.first_tile + j) * S.region * PART_RW);
                uint4 *dst = reinterpret_cast<uint4 *>(recs + (size_t)off * PART_RW);
                const uint32_t n16 = c * (PART_RW / 4);
                for (uint32_t i = tid; i < n16; i += 4 * TILE_THREADS) {
                    uint4 v[4];
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (i + u * TILE_THREADS < n16) v[u] = __ldcs(src + i + u * TILE_THREADS);
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (i + u * TILE_THREADS < n16) dst[i + u * TILE_THREADS] = v[u];
                }
            }
            off += c;
        }
    }
    if (bulk_bytes) {
        uint32_t done;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(mbar_a), "r"(0) : "memory");
        } while (!done);
    }
    __syncthreads();
    if (tid == 0) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(mbar_a) : "memory");   // (a block may stage again)
    return total;
}

// ---- the partition pass of the common large job -----------------------------------------------------
//
// Fixed-stride ACGTN keys of exactly 4*NW symbols, rows back to back, no quality filter, no
// multiplicities: everything the general ingest_kernel decides at run time is known here, which
// removes a third of its instructions (the pass is instruction-bound: ~1.7 G warp instructions
// for 100 M records).  Same record format, same partition function.
//
// What bounds it (ncu, profiles/r02_ncu_full_cfg5_summary.json): every record costs one cursor atomic and one 32-byte
// store to its own tile, i.e. two warp instructions with 32 distinct sectors each, and the nine shared-memory loads
// of a row queue behind them in the same load/store pipe (mio_throttle + short_scoreboard = 36 % of the stall
// samples, the atomics' round trip 25 %).  Measured and dropped in round 2: a persistent variant (two staging
// buffers, the bulk copy of chunk i+2 in flight while chunk i is packed, 64-bit shared loads, the stores of a chunk
// issued behind the atomics of the next one): 2.83 ms against 2.25 ms -- its 62 registers halve the resident warps
// and the load/store pipe, not latency, is the limit; 12 % fewer instructions (validity by byte permute) changed
// nothing either.
constexpr int LEAN_ROWS = 2;   // records per thread: both cursor atomics are in flight before the first store needs its position

template <int PW, int NW>
static __global__ void __launch_bounds__(256) partition_dna_kernel(const __grid_constant__ IngestParams P)
{
    constexpr int K = 3, KW = K * PW, ROWS = LEAN_ROWS;
    static_assert(slot_words(KW) == PART_RW, "partitioned plan: 32-byte records");
    __shared__ __align__(16) uint32_t stage[256 * ROWS * NW];
    const uint32_t tid = threadIdx.x;
    const uint64_t t0 = (uint64_t)blockIdx.x * (256u * ROWS);
    const uint32_t nblk = (uint32_t)min((uint64_t)(256u * ROWS), P.n - t0);
    const uint32_t nwords = nblk * NW;
    const uint32_t *src = reinterpret_cast<const uint32_t *>(P.keys + t0 * (4u * NW));
    if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0 && (nwords & 3u) == 0) {
        // one bulk asynchronous copy (TMA engine) instead of a load/store loop: the pass is instruction-bound
        bulk_load_to_shared(stage, src, nwords * 4u);
    } else {
        for (uint32_t i = tid; i < nwords; i += 256) stage[i] = __ldcs(src + i);
        __syncthreads();
    }
    uint32_t e[ROWS][PART_RW], part[ROWS], pos[ROWS];
    bool go[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
        const uint32_t row = r * 256u + tid;
        go[r] = row < nblk;
        if (!go[r]) continue;
        Key<K, PW> key;
        if (!pack_key_acgtn_fixed<PW, NW>(stage + row * NW, key)) {
            // report every unknown byte of this key so one retry with a grown alphabet suffices
            const uint8_t *kb = reinterpret_cast<const uint8_t *>(stage + row * NW);
            for (uint32_t i = 0; i < 4u * NW; i++) {
                const uint32_t c = kb[i];
                if (P.codec.lut[c] == 0xFF) atomicOr(&P.ctr->unknown[c >> 5], 1u << (c & 31));
            }
            go[r] = false;
            continue;
        }
        // (records partitioned by their pigeonhole block 0 never need the hash of the whole key)
        const uint64_t h = P.part_blocks ? block0_hash(key, block_start(4u * NW, 1, P.part_blocks), PART_SALT | (4u * NW)) : hash_key(key);
#pragma unroll
        for (int i = 0; i < PART_RW; i++) e[r][i] = 0;
#pragma unroll
        for (int i = 0; i < KW; i++) e[r][i] = key.w[i];
        e[r][KW] = 1u;
        e[r][KW + 1] = P.index_base + (uint32_t)(t0 + row);
        part[r] = part_of(h, P.part.nparts);
        pos[r] = atomicAdd(P.part.cursor + part[r], 1u);
    }
#pragma unroll
    for (int r = 0; r < ROWS; r++)
        if (go[r]) part_place(P.part, part[r], pos[r], e[r]);
}

// The same pass over keys the HOST has already packed (host_pack.cpp, pack_planes_parallel): the rows of a chunk back
// to back form one stream of symbols, and the chunk arrives as three bit streams -- bit t of plane p is code bit p of
// symbol t.  Row r owns bits [r * L, (r + 1) * L) of each stream: two 64-bit loads and one funnel shift per plane give
// the plane-major key, 13.5 bytes instead of 36 for the 36-nt configs is what crossed PCIe, and the host did no
// per-row work at all.  The bytes were validated by the packer.
template <int PW, int NW>
static __global__ void __launch_bounds__(256) partition_planes_kernel(const __grid_constant__ IngestParams P)
{
    constexpr int K = 3, KW = K * PW, ROWS = LEAN_ROWS;
    constexpr uint32_t L = 4u * NW;
    static_assert(L <= 64u && L <= 32u * PW, "a row's plane bits fit one 64-bit word and the key");
    static_assert(slot_words(KW) == PART_RW, "partitioned plan: 32-byte records");
    const uint32_t tid = threadIdx.x;
    const uint64_t t0 = (uint64_t)blockIdx.x * (256u * ROWS);
    const uint64_t *planes = reinterpret_cast<const uint64_t *>(P.keys);
    constexpr uint64_t row_mask = L >= 64u ? ~0ull : ((1ull << L) - 1ull);
    uint32_t e[ROWS][PART_RW], part[ROWS], pos[ROWS];
    bool go[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
        const uint64_t row = t0 + r * 256u + tid;
        go[r] = row < P.n;
        if (!go[r]) continue;
        const uint64_t bit = row * L, w = bit >> 6;
        const uint32_t sh = (uint32_t)(bit & 63u);
        Key<K, PW> key;
#pragma unroll
        for (int p = 0; p < K; p++) {
            const uint64_t *pl = planes + (size_t)p * P.plane_words + w;
            const uint64_t lo = __ldg(pl), hi = __ldg(pl + 1);   // (a zero word follows the stream)
            const uint64_t bits = (sh ? (lo >> sh) | (hi << (64u - sh)) : lo) & row_mask;
            key.w[p * PW] = (uint32_t)bits;
            if constexpr (PW >= 2) key.w[p * PW + 1] = (uint32_t)(bits >> 32);
#pragma unroll
            for (int i = 2; i < PW; i++) key.w[p * PW + i] = 0u;
        }
        const uint64_t h = P.part_blocks ? block0_hash(key, block_start(L, 1, P.part_blocks), PART_SALT | L) : hash_key(key);
#pragma unroll
        for (int i = 0; i < PART_RW; i++) e[r][i] = 0;
#pragma unroll
        for (int i = 0; i < KW; i++) e[r][i] = key.w[i];
        e[r][KW] = 1u;
        e[r][KW + 1] = P.index_base + (uint32_t)row;
        part[r] = part_of(h, P.part.nparts);
        pos[r] = atomicAdd(P.part.cursor + part[r], 1u);
    }
#pragma unroll
    for (int r = 0; r < ROWS; r++)
        if (go[r]) part_place(P.part, part[r], pos[r], e[r]);
}

// pass edges: (ui, uj | state) pairs waiting for apply_edges_kernel
struct EdgeSink {
    uint2 *edges;          // (ui, uj | state) pairs of this pass
    uint32_t *n_edges;
    uint32_t cap;
    uint32_t *overflow;    // set when a partition outgrew its tile: the pass is redone by the counting-sort plan
};

// What a tile does with two entries within the distance (closed forms: DESIGN.md "dissection").
template <int K, int PW>
__device__ __forceinline__ uint32_t tile_edge_flags(const PassParams &P, uint32_t ui, uint32_t uj, uint32_t ci, uint32_t cj,
                                                    const Key<K, PW> &ki, const Key<K, PW> &kj)
{
    uint32_t flag = EDGE_NONE;
    if (P.method == METHOD_DIRECTIONAL) {
        // closed form of reference __init__.py:60-91
        if (P.edge_flags) {
            // sharded job: ui / uj are global ids of other ranks' keys -- the consequence travels with the edge
            // (the five cases exclude each other: two counts >= 2 cannot dominate each other)
            if (ci == 1 && cj == 1) flag = EDGE_ONE;
            else if (ci == 1) flag = EDGE_DEAD_X;
            else if (cj == 1) flag = EDGE_DEAD_Y;
            else if ((unsigned long long)cj >= 2ull * ci - 1ull) flag = EDGE_DOM_X;
            else if ((unsigned long long)ci >= 2ull * cj - 1ull) flag = EDGE_DOM_Y;
            return flag;
        }
        if (ci >= 2 && (unsigned long long)cj >= 2ull * ci - 1ull) P.dominated[ui] = 1;
        if (cj >= 2 && (unsigned long long)ci >= 2ull * cj - 1ull) P.dominated[uj] = 1;
        if (ci == 1 && cj == 1) flag = EDGE_ONE;
        else if (ci == 1) P.dead[ui] = 1;
        else if (cj == 1) P.dead[uj] = 1;
    } else if (P.method == METHOD_ADJACENCY) {
        const bool i_less = prio_less<K, PW>(ci, ki, cj, kj, P.rank_of_code);
        const unsigned long long pos = aggregated_inc64(&P.ctr->n_edges);
        if (pos < P.edge_cap) P.edges[pos] = i_less ? make_uint2(uj, ui) : make_uint2(ui, uj);
    }
    return flag;
}

// Job statistics of a tile kernel: one atomic per warp, spread over STAT_SPREAD counters by block
// (per-warp atomics on ONE address serialise at ~0.5 ns each and end up bounding the kernel; a
// block-wide reduction costs three barriers per tile).  The host folds the counters.
__device__ __forceinline__ void tile_stats(DevCounters *ctr, uint32_t merges, uint32_t cand)
{
    for (int o = 16; o; o >>= 1) {
        merges += __shfl_xor_sync(WARP_FULL, merges, o);
        cand += __shfl_xor_sync(WARP_FULL, cand, o);
    }
    if ((threadIdx.x & 31) == 0) {
        const uint32_t k = (blockIdx.x * 8u + (threadIdx.x >> 5)) % STAT_SPREAD;
        if (merges) atomicAdd(&ctr->merge_spread[k], merges);
        if (cand) atomicAdd(&ctr->cand_spread[k], (unsigned long long)cand);
    }
}

// One pigeonhole pass over `n` entries of a staged tile (entry k is record index_of(k); its
// record carries {key, count, unique id}).  The table (`tsize` entries, a power of two) is used
// as a MULTIMAP keyed by the block hash: the probe sequence of an entry starts at the hash of
// its pigeonhole block, so all entries of one bucket share it.  Of two entries of a bucket the
// one that claims its table entry later has walked over the other one (entries are never
// released), so every in-bucket pair is tested exactly once, with XOR+POPC Hamming; entries of
// other buckets that happen to lie on the walk are tested too, which is harmless (a hit is a
// true edge).  Hits set the directional flags at once; their union-find hooks are buffered in
// `s_edge` (as tile-local index pairs) and leave the tile as one contiguous run of the edge list.
// (Measured slower on B200: recording the met pairs and verifying them in a second, dense loop
// [+10 %]; claiming all table entries first and walking the own span after a barrier [+12 %].)
template <int K, int PW, bool CAN_UNION, typename IndexOf, typename SlotOf>
__device__ __forceinline__ void tile_bucket_pass(const uint32_t *recs, uint32_t *tab, uint32_t tsize, uint32_t *s_edge,
                                                 uint32_t cap_e, uint32_t n, IndexOf index_of, SlotOf slot_of, const PassParams &P,
                                                 const EdgeSink &E, uint32_t &merges, uint32_t &cand)
{
    constexpr int KW = K * PW;
    __shared__ uint32_t sc[2];   // [0] edges recorded, [1] base of the edges in the global list
    const uint32_t tid = threadIdx.x, tmask = tsize - 1;
    for (uint32_t i = tid; i < tsize; i += TILE_THREADS) tab[i] = TILE_EMPTY;
    if (tid == 0) sc[0] = 0;
    __syncthreads();
#pragma unroll 1
    for (uint32_t k = tid; k < n; k += TILE_THREADS) {
        const uint32_t i = index_of(k);
        uint32_t e[PART_RW];
        tile_load_rec(recs, i, e);
        Key<K, PW> ki;
#pragma unroll
        for (int w = 0; w < KW; w++) ki.w[w] = e[w];
        const uint32_t ci = e[KW], ui = e[KW + 1];
        uint32_t s = slot_of(ki) & tmask;   // a function of the entry's pigeonhole block only
        for (;;) {
            uint32_t cur = *reinterpret_cast<volatile uint32_t *>(tab + s);
            if (cur == TILE_EMPTY) {
                cur = atomicCAS(tab + s, TILE_EMPTY, i);
                if (cur == TILE_EMPTY) break;
            }
            uint32_t f[PART_RW];
            tile_load_rec(recs, cur, f);
            Key<K, PW> kj;
#pragma unroll
            for (int w = 0; w < KW; w++) kj.w[w] = f[w];
            cand++;
            if (hamming_within<K, PW>(ki, kj, P.d, P.varlen != 0, P.pad_code)) {
                const uint32_t flag = tile_edge_flags<K, PW>(P, ui, f[KW + 1], ci, f[KW], ki, kj);
                const uint32_t epos = atomicAdd(&sc[0], 1u);
                if (epos < cap_e) {
                    s_edge[epos] = i | (cur << 10) | flag;
                } else {   // dense tile (a big family): this edge goes to the global list on its own
                    const uint32_t gpos = atomicAdd(E.n_edges, 1u);
                    if (gpos < E.cap) {
                        E.edges[gpos] = make_uint2(ui, f[KW + 1] | flag);
                    } else if (CAN_UNION && !P.edge_flags) {
                        if (uf_union(P.parent_full, ui, f[KW + 1])) merges++;
                        if (flag == EDGE_ONE) uf_union(P.parent_one, ui, f[KW + 1]);
                    } else {
                        *E.overflow = 1u;           // no forest to hook into (yet / on this rank): the caller redoes the pass
                    }
                }
            }
            s = (s + 1) & tmask;
        }
    }
    __syncthreads();
    const uint32_t ne = min(sc[0], cap_e);
    if (tid == 0) sc[1] = ne ? atomicAdd(E.n_edges, ne) : 0u;
    __syncthreads();
    for (uint32_t k = tid; k < ne; k += TILE_THREADS) {
        const uint32_t pos = sc[1] + k;
        const uint32_t ed = s_edge[k];
        const uint32_t ui = recs[(size_t)(ed & 1023u) * PART_RW + KW + 1], uj = recs[(size_t)((ed >> 10) & 1023u) * PART_RW + KW + 1];
        if (pos < E.cap) {
            E.edges[pos] = make_uint2(ui, uj | (ed & EDGE_STATE));
        } else if (CAN_UNION && !P.edge_flags) {
            if (uf_union(P.parent_full, ui, uj)) merges++;
            if ((ed & EDGE_STATE) == EDGE_ONE) uf_union(P.parent_one, ui, uj);
        } else {
            *E.overflow = 1u;
        }
    }
}

// ---- stage A: exact dedupe of one partition --------------------------------------------------------

struct DedupeOut {
    uint32_t *ukey, *ucount, *ufirst;
    uint32_t *n_unique;     // dense output cursor
    uint32_t *oversize;     // list of partitions that outgrew their region (handled by the spill path)
    uint32_t *n_oversize;
    int keep_zero;          // replicated-set plan: keep keys whose every record was filtered (weight 0)
    uint32_t cap;           // entries of the dense arrays (a tile that would write past them sets *overflow)
    uint32_t *overflow;
};

// FUSED: the records were partitioned by the hash of pigeonhole block 0, so a tile holds whole
// pass-0 buckets: once the uniques of the tile have their dense ids, the same table is reused
// as the block multimap (see bucket_tile_kernel) and pass 0 never touches HBM again.  The
// forest does not exist yet (U is unknown), so hits can only be deferred: when a buffer is
// too small the overflow flag makes the caller run pass 0 the ordinary way as well.
// EMIT (with FUSED): the uniques leave a second time as {key, count, id} entries of the NEXT pass,
// appended to the tiles of its block hash (X.next) -- that pass then starts with its tiles in
// place instead of reading the unique set back and partitioning it.
struct NextPass {
    PartParams next;        // next.buf == null: nothing to emit
    uint32_t st, bl;        // block of the next pass for keys of max_len symbols (cf. PassParams::fix_st)
    int pass_j;
};

template <int K, int PW, bool FUSED, int TR>
__device__ __forceinline__ void dedupe_tile_body(const uint32_t p, const TileSource &S, const DedupeOut &O, const PassParams &P,
                                                 const EdgeSink &E, const NextPass &X)
{
    constexpr int KW = K * PW, RW = slot_words(KW);
    static_assert(RW == PART_RW, "partitioned plan: 32-byte records");
    static_assert(TR % TILE_THREADS == 0 && TR <= TILE_R, "tile capacity");
    __shared__ __align__(16) uint32_t recs[TR * PART_RW];
    __shared__ uint32_t tab[TILE_T];
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t total, s_base;
    const uint32_t tid = threadIdx.x;   // p: owned tile
    const uint32_t cnt = stage_tile(recs, tab, TILE_T, S, p);
    if (cnt == 0) return;
    if (cnt > (uint32_t)TILE_R) {
        if (tid == 0) O.oversize[atomicAdd(O.n_oversize, 1u)] = p;
        return;
    }

    // A record either wins an empty table entry (it becomes the representative of its key) or
    // meets the representative and adds its weight / lowers the first index there.
    uint32_t rep = 0;   // bit r: my record of round r is a representative
#pragma unroll
    for (int r = 0; r < TR / TILE_THREADS; r++) {
        const uint32_t i = r * TILE_THREADS + tid;
        if (i >= cnt) break;
        uint32_t e[PART_RW];
        tile_load_rec(recs, i, e);
        Key<K, PW> key;
#pragma unroll
        for (int j = 0; j < KW; j++) key.w[j] = e[j];
        uint32_t s = hash_key32(key) & (TILE_T - 1);
        for (;;) {
            uint32_t cur = *reinterpret_cast<volatile uint32_t *>(tab + s);
            if (cur == TILE_EMPTY) {
                cur = atomicCAS(tab + s, TILE_EMPTY, i);
                if (cur == TILE_EMPTY) { rep |= 1u << r; break; }
            }
            const uint32_t *o = recs + (size_t)cur * PART_RW;
            uint32_t diff = 0;
#pragma unroll
            for (int j = 0; j < KW; j++) diff |= o[j] ^ e[j];
            if (diff == 0) {
                if (e[KW]) atomicAdd(recs + (size_t)cur * PART_RW + KW, e[KW]);
                atomicMin(recs + (size_t)cur * PART_RW + KW + 1, e[KW + 1]);
                break;
            }
            s = (s + 1) & (TILE_T - 1);
        }
    }
    __syncthreads();

    // representatives -> dense unique arrays (keys whose every record was filtered are dropped).
    // Their tile-local indices are compacted first so that the copy-out (and the fused pass 0)
    // run with full warps and neighbouring threads write neighbouring unique ids.
    __shared__ uint16_t replist[TR];
    uint32_t nval = 0;
#pragma unroll
    for (int r = 0; r < TR / TILE_THREADS; r++) {
        if (!((rep >> r) & 1u)) continue;
        const uint32_t i = r * TILE_THREADS + tid;
        if (recs[(size_t)i * PART_RW + KW] != 0 || O.keep_zero) nval++;
        else rep &= ~(1u << r);
    }
    uint32_t off = block_exclusive_scan(nval, &total, warp_sums);
    if (tid == 0) s_base = total ? atomicAdd(O.n_unique, total) : 0u;
#pragma unroll
    for (int r = 0; r < TR / TILE_THREADS; r++)
        if ((rep >> r) & 1u) replist[off++] = (uint16_t)(r * TILE_THREADS + tid);
    __syncthreads();
    const uint32_t nrep = total, base = s_base;
    if (base + nrep > O.cap) {   // (only a sharded job sizes the arrays below the record count)
        if (tid == 0) *O.overflow = 1u;
        return;
    }
    for (uint32_t k = tid; k < nrep; k += TILE_THREADS) {
        const uint32_t i = replist[k], pos = base + k;
        uint32_t e[PART_RW];
        tile_load_rec(recs, i, e);
#pragma unroll
        for (int j = 0; j < KW; j++) O.ukey[(size_t)pos * KW + j] = e[j];
        O.ucount[pos] = e[KW];
        O.ufirst[pos] = e[KW + 1];
        if constexpr (FUSED) {
            const uint32_t gid = pos * S.G + S.self;    // unique id in the job-wide id space (ranks interleaved)
            recs[(size_t)i * PART_RW + KW + 1] = gid;   // the record now carries its unique id
            if (X.next.buf) {
                Key<K, PW> ki;
#pragma unroll
                for (int j = 0; j < KW; j++) ki.w[j] = e[j];
                uint32_t st = X.st, bl = X.bl, len = P.max_len;
                if (P.varlen) {
                    len = key_length(ki, P.pad_code, P.max_len);
                    st = block_start(len, (uint32_t)X.pass_j, (uint32_t)P.d + 1u);
                    bl = block_start(len, (uint32_t)X.pass_j + 1u, (uint32_t)P.d + 1u) - st;
                }
                const uint64_t sig = block_hash(ki, st, bl, ((uint64_t)X.pass_j << 32) | len);   // == pass_variant of that pass
                e[KW + 1] = gid;
                part_append(X.next, part_of(sig, X.next.nparts), e);
            }
        }
    }

    if constexpr (FUSED) {
        __syncthreads();   // ids are in place, the dedupe probes are over: the table is free for pass 0
        uint32_t merges = 0, cand = 0;
        // the uniques of a tile need half the table; its other half buffers the edges.  (Measured and dropped in round 2:
        // the unions of pass 0 inside the tile -- 16-bit tile-local forests in that half, every unique leaving with its
        // parent links -- save init_forest + the pass-0 apply_edges, 0.2 ms, and cost the tile kernel the same 0.2 ms.)
        tile_bucket_pass<K, PW, false>(
            recs, tab, TILE_T / 2, tab + TILE_T / 2, TILE_T / 2, nrep, [&](uint32_t k) { return (uint32_t)replist[k]; },
            [&](const Key<K, PW> &ki) {   // pass 0: the leading block
                const uint32_t len = P.varlen ? key_length(ki, P.pad_code, P.max_len) : P.max_len;
                return (uint32_t)block0_hash(ki, block_start(len, 1, (uint32_t)P.d + 1u), (uint64_t)len);
            },
            P, E, merges, cand);
        tile_stats(P.ctr, 0, cand);
    }
}

// TR: records the shared-memory tile holds; MINB: resident blocks per SM the register allocation must allow
template <int K, int PW, bool FUSED, int TR = TILE_R, int MINB = 10>
static __global__ void __launch_bounds__(TILE_THREADS, MINB) dedupe_tile_kernel(const __grid_constant__ TileSource S,
                                                                                const __grid_constant__ DedupeOut O,
                                                                                const __grid_constant__ PassParams P,
                                                                                const __grid_constant__ EdgeSink E,
                                                                                const __grid_constant__ NextPass X)
{
    if (S.list) {
        const uint32_t n = *S.n_list;
        for (uint32_t t = blockIdx.x; t < n; t += gridDim.x) {
            dedupe_tile_body<K, PW, FUSED, TR>(S.list[t], S, O, P, E, X);
            __syncthreads();   // the next tile reuses the shared memory
        }
    } else {
        dedupe_tile_body<K, PW, FUSED, TR>(blockIdx.x, S, O, P, E, X);
    }
}

// Tiles whose fill exceeds `small` (they need the full-size launch), as a list; one GPU only (fill = the cursor).
static __global__ void __launch_bounds__(256) classify_tiles_kernel(const uint32_t *__restrict__ cursor, uint32_t nparts, uint32_t small,
                                                                    uint32_t *__restrict__ list, uint32_t *n_list)
{
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    const bool big = i < nparts && cursor[i] > small;
    const uint32_t pos = block_reserve(big, n_list);
    if (big) list[pos] = i;
}

// Spill path: the records of the oversize partitions (their full regions: TILE_R / 256 consecutive
// blocks per partition, all on blockIdx.x -- a skewed 100 M-read job can have more oversize partitions
// than gridDim.y allows) or of the spill buffer (list == null) go through the single-table insert.
template <int K, int PW>
static __global__ void __launch_bounds__(256) spill_insert_kernel(const uint32_t *__restrict__ recs, const uint32_t *__restrict__ list,
                                                                  uint32_t n, const __grid_constant__ TableRef tab,
                                                                  uint32_t *n_claimed)
{
    constexpr int KW = K * PW, RW = slot_words(KW);
    static_assert(RW == PART_RW, "partitioned plan: 32-byte records");
    constexpr uint32_t BPP = TILE_R / 256;   // blocks per oversize partition
    const uint32_t i = (list ? blockIdx.x % BPP : blockIdx.x) * 256u + threadIdx.x;
    const uint32_t *src = list ? recs + (size_t)list[blockIdx.x / BPP] * TILE_R * PART_RW : recs;
    uint32_t claimed = NO_CLAIM;
    if (i < n) {
        uint32_t e[PART_RW];
        load_rec_stream(src + (size_t)i * PART_RW, e);
        Key<K, PW> key;
#pragma unroll
        for (int j = 0; j < KW; j++) key.w[j] = e[j];
        claimed = table_insert<K, PW>(tab, key, hash_key(key), e[KW + 1], e[KW]);
    }
    const uint32_t pos = block_reserve(claimed != NO_CLAIM, n_claimed);
    if (claimed != NO_CLAIM) tab.uslot[pos] = claimed;
}

// ---- stage B: one Hamming pigeonhole pass -------------------------------------------------------------

constexpr int BP_ROWS = 2;   // uniques per thread: the loads and the cursor atomics of both are in flight together

template <int K, int PW>
static __global__ void __launch_bounds__(256) bucket_partition_kernel(const __grid_constant__ PassParams P,
                                                                      const __grid_constant__ PartParams Q)
{
    constexpr int KW = K * PW, RW = fat_words(KW);
    static_assert(RW == PART_RW, "partitioned plan: 32-byte records");
    uint32_t e[BP_ROWS][RW], part[BP_ROWS], pos[BP_ROWS];
    bool go[BP_ROWS];
#pragma unroll
    for (int r = 0; r < BP_ROWS; r++) {
        const uint64_t u = (uint64_t)P.u_lo + ((uint64_t)blockIdx.x * BP_ROWS + r) * 256u + threadIdx.x;
        go[r] = u < (P.U_dev ? min(P.U, *P.U_dev) : P.U);
        if (!go[r]) continue;
        Key<K, PW> key;
        load_key_stream<K, PW>(P.ukey, (uint32_t)u, key);
#pragma unroll
        for (int i = 0; i < RW; i++) e[r][i] = 0;
#pragma unroll
        for (int i = 0; i < KW; i++) e[r][i] = key.w[i];
        e[r][KW] = __ldcs(P.ucount + u);
        e[r][KW + 1] = (uint32_t)u * P.id_mul + P.id_add;   // job-wide id (sharded: ranks interleaved)
    }
#pragma unroll
    for (int r = 0; r < BP_ROWS; r++) {
        if (!go[r]) continue;
        Key<K, PW> key;
#pragma unroll
        for (int i = 0; i < KW; i++) key.w[i] = e[r][i];
        const uint32_t len = P.varlen ? key_length(key, P.pad_code, P.max_len) : P.max_len;
        uint64_t sig;
        bool build;
        pass_variant<K, PW>(key, len, P, 0, sig, build);
        go[r] = !(P.world > 1 && !P.edge_flags && (uint32_t)(sig >> 32) % (uint32_t)P.world != (uint32_t)P.my_rank);
        part[r] = part_of(sig, Q.nparts);
        if (go[r]) pos[r] = atomicAdd(Q.cursor + part[r], 1u);
    }
#pragma unroll
    for (int r = 0; r < BP_ROWS; r++)
        if (go[r] && pos[r] < Q.region) store_rec_stream(Q.buf + ((size_t)part[r] * Q.region + pos[r]) * PART_RW, e[r]);
}

template <int K, int PW, int TR>
__device__ __forceinline__ void bucket_tile_body(const uint32_t p, const TileSource &S, const PassParams &P, const EdgeSink &E)
{
    constexpr int KW = K * PW, RW = fat_words(KW);
    static_assert(RW == PART_RW, "partitioned plan: 32-byte records");
    __shared__ __align__(16) uint32_t recs[TR * PART_RW];
    __shared__ uint32_t tab[TILE_T];
    __shared__ uint32_t s_edges[TILE_E];
    const uint32_t tid = threadIdx.x;   // p: owned tile
    const uint32_t cnt = stage_tile(recs, tab, 0, S, p);   // (the pass clears the table itself)
    if (cnt == 0) return;
    if (cnt > (uint32_t)TILE_R) {
        if (tid == 0) *E.overflow = 1u;
        return;
    }
    uint32_t merges = 0, cand = 0;
    tile_bucket_pass<K, PW, true>(   // (starts with a barrier)
        recs, tab, TILE_T, s_edges, TILE_E, cnt, [](uint32_t k) { return k; },
        [&](const Key<K, PW> &ki) {
            const uint32_t len = P.varlen ? key_length(ki, P.pad_code, P.max_len) : P.max_len;
            uint64_t sig;
            bool build;
            pass_variant<K, PW>(ki, len, P, 0, sig, build);
            return (uint32_t)sig;
        },
        P, E, merges, cand);
    tile_stats(P.ctr, merges, cand);
}

template <int K, int PW, int TR = TILE_R, int MINB = 10>
static __global__ void __launch_bounds__(TILE_THREADS, MINB) bucket_tile_kernel(const __grid_constant__ TileSource S,
                                                                                const __grid_constant__ PassParams P,
                                                                                const __grid_constant__ EdgeSink E)
{
    if (S.list) {
        const uint32_t n = *S.n_list;
        for (uint32_t t = blockIdx.x; t < n; t += gridDim.x) {
            bucket_tile_body<K, PW, TR>(S.list[t], S, P, E);
            __syncthreads();   // the next tile reuses the shared memory
        }
    } else {
        bucket_tile_body<K, PW, TR>(blockIdx.x, S, P, E);
    }
}

// All pairs among the uniques [lo, hi) (the handful that left the dedupe stage through the spill
// path and so were not compared inside a tile by the fused pass 0).
template <int K, int PW>
static __global__ void __launch_bounds__(256) range_pairs_kernel(const __grid_constant__ PassParams P, uint32_t lo, uint32_t hi)
{
    const uint32_t i = lo + blockIdx.x * 256u + threadIdx.x;
    const uint32_t j0 = lo + blockIdx.y * 256u, j1 = min(hi, j0 + 256u);   // this block's slice of partners
    uint32_t merges = 0, cand = 0;
    if (i < hi && j1 > i + 1) {
        Key<K, PW> ki;
        load_key<K, PW>(P.ukey, i, ki);
        const uint32_t ci = P.ucount[i];
        for (uint32_t j = max(i + 1, j0); j < j1; j++) {
            Key<K, PW> kj;
            load_key<K, PW>(P.ukey, j, kj);
            cand++;
            if (hamming_within<K, PW>(ki, kj, P.d, P.varlen != 0, P.pad_code))
                process_edge<K, PW>(P, i, j, ci, P.ucount[j], ki, kj, merges);
        }
    }
    for (int o = 16; o; o >>= 1) {
        merges += __shfl_xor_sync(WARP_FULL, merges, o);
        cand += __shfl_xor_sync(WARP_FULL, cand, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (merges) atomicAdd(&P.ctr->n_merges, merges);
        if (cand) atomicAdd(&P.ctr->n_candidates, (unsigned long long)cand);
    }
}

// The edge lists of all ranks (one list on one GPU).  A sharded job never gathers them: every rank's
// apply_edges_kernel streams the peers' lists out of their HBM over NVLink while it hooks -- the all-gather of
// the edges is fused into the union-find that consumes them.
struct EdgeSource {
    const uint2 *edges[MAX_RANKS];
    uint32_t cap[MAX_RANKS];
    const uint32_t *n_edges;    // [G * n_stride] edge counters of all ranks (gathered; a local copy)
    uint32_t n_stride;          // counter of rank g: n_edges[g * n_stride]
    const uint32_t *first;      // optional: the lists are read from first[g * first_stride] on (what an earlier launch,
    uint32_t first_stride;      // overlapped with the next pass, has already applied)
    uint32_t G, self;
    uint32_t id_stride;         // slot of id (u * G + rank): rank * id_stride + u, see slot_of_id
};

// Ids travel as (local id * G + rank): a tile kernel can number its uniques without knowing how many the other
// ranks have.  The forests and flag arrays are indexed rank-major instead (slot = rank * stride + local id, the
// stride agreed once the unique counts are known): the edges of one rank's tiles then touch one contiguous
// block, as they do on one GPU -- interleaved slots would drag every 32-byte sector of the arrays through the
// caches once per rank.
__device__ __forceinline__ uint32_t slot_of_id(uint32_t id, uint32_t G, uint32_t stride)
{
    return G == 1 ? id : (id % G) * stride + id / G;
}

// Per-key state of the dissection that edges set: byte arrays over THIS RANK'S uniques (index = local unique id; on
// one GPU that is the id itself).  A rank applies every edge of every rank to its forests, but only the flags of its
// own keys: what the other members of a component contribute reaches it through the candidate lists.
struct EdgeFlags {
    uint8_t *dominated;   // directional
    uint8_t *dead;        // directional: the count-1 key touches a key with count >= 2
    uint8_t *linked;      // the key has an edge of the kind the dissection reduces over (directional: count-1 edges;
                          // highest_count: any edge) -- keys without one are their own component and need no exchange
    int any_edge;         // highest_count: `linked` for every edge
};

static __global__ void __launch_bounds__(256) apply_edges_kernel(const __grid_constant__ EdgeSource E, uint32_t *parent_full,
                                                                 uint32_t *parent_one, const __grid_constant__ EdgeFlags F,
                                                                 DevCounters *ctr)
{
    uint32_t merges = 0;
    const uint32_t stride = gridDim.x * 256u;
    for (uint32_t k = 0; k < E.G; k++) {
        const uint32_t g = (k + blockIdx.x) % E.G;   // blocks start on different peers: all NVLink ports busy at once
        const uint32_t n = min(E.n_edges[(size_t)g * E.n_stride], E.cap[g]);
        const uint32_t lo = E.first ? min(E.first[(size_t)g * E.first_stride], n) : 0u;
        const uint2 *edges = E.edges[g];
        // The list may live in a peer's HBM (2-3 us per load over NVLink): a thread keeps the loads of its next
        // APPLY_AHEAD edges in flight while it hooks the current one, in list order (the edges of a pass-0 tile are
        // neighbours in the forest; batching the fetches instead broke that locality and was slower).
        constexpr int APPLY_AHEAD = 2;
        uint32_t i0 = lo + blockIdx.x * 256u + threadIdx.x;
        uint2 ahead[APPLY_AHEAD];
#pragma unroll
        for (int a = 0; a < APPLY_AHEAD; a++)
            if ((uint64_t)i0 + (uint64_t)a * stride < n) ahead[a] = __ldcs(edges + i0 + a * stride);
        for (; i0 < n; i0 += stride) {
            const uint2 ed = ahead[0];
#pragma unroll
            for (int a = 0; a + 1 < APPLY_AHEAD; a++) ahead[a] = ahead[a + 1];
            if ((uint64_t)i0 + (uint64_t)APPLY_AHEAD * stride < n) ahead[APPLY_AHEAD - 1] = __ldcs(edges + i0 + APPLY_AHEAD * stride);
            {
                const uint32_t st = ed.y & EDGE_STATE, idx = ed.x, idy = ed.y & EDGE_ID;
                const uint32_t x = slot_of_id(idx, E.G, E.id_stride), y = slot_of_id(idy, E.G, E.id_stride);
                if (uf_union(parent_full, x, y)) merges++;
                if (st == EDGE_ONE && parent_one) uf_union(parent_one, x, y);
                if (st == EDGE_NONE && !F.any_edge) continue;
                // flags of this rank's own keys (local id = id / G)
                const bool x_mine = idx % E.G == E.self, y_mine = idy % E.G == E.self;
                const uint32_t ux = idx / E.G, uy = idy / E.G;
                if (st == EDGE_ONE || F.any_edge) {
                    if (F.linked && x_mine) F.linked[ux] = 1;
                    if (F.linked && y_mine) F.linked[uy] = 1;
                } else if (st == EDGE_DEAD_X) {
                    if (x_mine) F.dead[ux] = 1;
                } else if (st == EDGE_DEAD_Y) {
                    if (y_mine) F.dead[uy] = 1;
                } else if (st == EDGE_DOM_X) {
                    if (x_mine) F.dominated[ux] = 1;
                } else if (st == EDGE_DOM_Y) {
                    if (y_mine) F.dominated[uy] = 1;
                }
            }
        }
    }
    for (int o = 16; o; o >>= 1) merges += __shfl_xor_sync(WARP_FULL, merges, o);
    if ((threadIdx.x & 31) == 0 && merges) atomicAdd(&ctr->n_merges, merges);
}

// ---- dissection of a tile-sharded job ---------------------------------------------------------------
//
// Every rank holds the whole forests (it applied all edges) but only ITS keys and their flags.  Flags and roots
// are enough for most keys; where a component's answer depends on several members (the largest key of a count-1
// component and whether any member is dead; the largest (count, key) of a cluster) the members that matter
// become CANDIDATES -- {key, count, id} records in each rank's candidate list -- and every rank reduces all
// candidate lists (read from the peers' HBM) into best[root] / deadroot[root].  A dead member travels as a record
// with count 0.

// One status block per rank and agreement point, packed on the device (no host round trip before the exchange):
// the plan's counters and the job counters the ranks have to agree on or add up.
constexpr int STATUS_WORDS = 24;   // uint64 each
struct StatusParams {
    long long rc;
    const uint32_t *ctrs;          // the plan's counters (sharded_tiles.cuh CTR_*)
    const DevCounters *ctr;
    uint32_t cap_u, spill_cap, cap_e, cap_c;
    unsigned long long *out;
};
static __global__ void pack_status_kernel(const __grid_constant__ StatusParams P)
{
    if (threadIdx.x != 0) return;
    const uint32_t *c = P.ctrs;
    const DevCounters &d = *P.ctr;
    unsigned long long *o = P.out;
    o[0] = (unsigned long long)P.rc;
    o[1] = min(c[2], P.cap_u);
    o[2] = c[3];
    o[3] = min(c[0], P.spill_cap);
    o[4] = (c[1] | c[7] | c[6] | (c[5] > P.cap_e) | (c[8] > P.cap_c)) ? 1ull : 0ull;
    o[5] = d.phred_err;
    o[6] = d.n_discarded;
    o[7] = d.sum_weights;
    for (int k = 0; k < 8; k++) o[8 + k] = d.unknown[k];
    unsigned long long cand = d.n_candidates, merges = d.n_merges;
    for (uint32_t k = 0; k < STAT_SPREAD; k++) { cand += d.cand_spread[k]; merges += d.merge_spread[k]; }
    o[16] = d.n_selected;
    o[17] = cand;
    o[18] = merges;
    o[19] = d.n_edges;
    o[20] = d.table_full;
    for (int k = 21; k < STATUS_WORDS; k++) o[k] = 0;
}

struct CandParams {
    uint32_t U;               // this rank's uniques (an upper bound when U_dev is set)
    const uint32_t *U_dev;    // the count on the device (the spill path may still be adding to it when the launch is sized)
    uint32_t G, self, id_stride;
    const uint32_t *ukey, *ucount;
    uint32_t *forest;         // parent_one (directional) / parent_full (highest_count)
    const uint8_t *dead, *linked;   // per own unique
    uint32_t *root_of;        // per own unique: root of its component in `forest`
    uint32_t *loc_of;         // per own unique: (self << 28 | index) of its candidate record, or NO_CLAIM
    uint32_t *cand;           // candidate records {key[KW], count, id} of PART_RW words
    uint32_t *cand_root;
    uint32_t *n_cand;
    uint32_t cap;
    int method;
};

template <int K, int PW>
static __global__ void __launch_bounds__(256) candidates_kernel(const __grid_constant__ CandParams P)
{
    constexpr int KW = K * PW;
    static_assert(slot_words(KW) == PART_RW, "candidate records are 32 bytes");
    const uint32_t u = blockIdx.x * 256u + threadIdx.x;
    const uint32_t U = P.U_dev ? min(P.U, *P.U_dev) : P.U;
    bool emit = false;
    uint32_t root = 0, gid = 0, c = 0;
    if (u < U) {
        gid = P.self * P.id_stride + u;   // slot of the key in the forests
        c = P.ucount[u];
        root = gid;
        if (P.linked[u] && (P.method != METHOD_DIRECTIONAL || c == 1)) {
            root = uf_find(P.forest, gid);
            emit = true;
            if (P.method == METHOD_DIRECTIONAL && P.dead[u]) c = 0;   // a dead member: kills the component, never wins it
        }
        P.root_of[u] = root;
    }
    const uint32_t pos = block_reserve(emit, P.n_cand);
    if (u < U) P.loc_of[u] = (emit && pos < P.cap) ? ((P.self << 28) | pos) : NO_CLAIM;
    if (emit && pos < P.cap) {
        uint32_t e[PART_RW];
#pragma unroll
        for (int i = 0; i < PART_RW; i++) e[i] = 0;
        Key<K, PW> key;
        load_key_stream<K, PW>(P.ukey, u, key);
#pragma unroll
        for (int i = 0; i < KW; i++) e[i] = key.w[i];
        e[KW] = c;
        e[KW + 1] = gid;
        store_rec_stream(P.cand + (size_t)pos * PART_RW, e);
        P.cand_root[pos] = root;
    }
}

struct BestParams {
    const uint32_t *cand[MAX_RANKS];
    const uint32_t *cand_root[MAX_RANKS];
    uint32_t cap[MAX_RANKS];
    const uint32_t *n_cand;   // [G * n_stride] (gathered)
    uint32_t n_stride;
    uint32_t G;
    uint32_t *best;           // per root: (rank << 28 | index) of the best candidate so far; NO_CLAIM = none
    uint8_t *deadroot;        // per root: a member of the count-1 component is dead (directional)
    uint8_t rank_of_code[256];
};

// Per root keep the candidate with the largest (count, key) -- the head of the reference's descending sort
// (__init__.py:99-101) resp. the survivor of a count-1 chain (:72, :91).  Replicated on every rank.
template <int K, int PW>
static __global__ void __launch_bounds__(256) best_candidate_kernel(const __grid_constant__ BestParams P)
{
    constexpr int KW = K * PW;
    for (uint32_t k = 0; k < P.G; k++) {
        const uint32_t g = (k + blockIdx.x) % P.G;
        const uint32_t n = min(P.n_cand[(size_t)g * P.n_stride], P.cap[g]);
        // (a peer's list: the loads of the next record are in flight while this one is reduced)
        const uint32_t step = gridDim.x * 256u;
        uint32_t i = blockIdx.x * 256u + threadIdx.x;
        uint32_t ne[PART_RW], nr = 0;
        if (i < n) { load_rec_stream(P.cand[g] + (size_t)i * PART_RW, ne); nr = __ldcs(P.cand_root[g] + i); }
        for (; i < n; i += step) {
            uint32_t e[PART_RW];
#pragma unroll
            for (int w = 0; w < PART_RW; w++) e[w] = ne[w];
            const uint32_t r = nr;
            if ((uint64_t)i + step < n) { load_rec_stream(P.cand[g] + (size_t)(i + step) * PART_RW, ne); nr = __ldcs(P.cand_root[g] + i + step); }
            if (e[KW] == 0) { P.deadroot[r] = 1; continue; }   // (only directional jobs send count 0)
            const uint32_t me = (g << 28) | i;
            Key<K, PW> km;
#pragma unroll
            for (int w = 0; w < KW; w++) km.w[w] = e[w];
            uint32_t cur = ld_relaxed_u32(P.best + r);
            for (;;) {
                if (cur != NO_CLAIM) {
                    uint32_t f[PART_RW];
                    tile_load_rec(P.cand[cur >> 28], cur & 0x0FFFFFFFu, f);
                    Key<K, PW> kc;
#pragma unroll
                    for (int w = 0; w < KW; w++) kc.w[w] = f[w];
                    if (!prio_less<K, PW>(f[KW], kc, e[KW], km, P.rank_of_code)) break;   // cur >= me
                }
                const uint32_t old = atomicCAS(P.best + r, cur, me);
                if (old == cur) break;
                cur = old;
            }
        }
    }
}

struct SelectOwnParams {
    uint32_t U, G, self, id_stride;
    const uint32_t *U_dev;
    const uint32_t *ucount, *ufirst;
    const uint32_t *root_of, *loc_of, *best;
    const uint8_t *dominated, *dead, *linked;   // per own unique
    const uint8_t *deadroot, *state;            // per slot
    uint8_t *selected;        // per own unique
    uint32_t *bitmap;         // this rank's keep bits over ALL records of the job (bit = global record index - bit_base)
    uint32_t bit_base;
    int method;
    DevCounters *ctr;
};

// The keys of this rank decide.  The keep bit of a selected key belongs to the rank that holds the key's first
// record (reference pass 2, __init__.py:201-206): every rank marks its selected keys in a bitmap over the records
// of the whole job, and merge_bitmaps_kernel on each rank ORs together the peers' slices of its own records.
static __global__ void __launch_bounds__(256) select_own_kernel(const __grid_constant__ SelectOwnParams P)
{
    const uint32_t u = blockIdx.x * 256u + threadIdx.x;
    const uint32_t U = P.U_dev ? min(P.U, *P.U_dev) : P.U;
    bool sel = false;
    if (u < U) {
        const uint32_t gid = P.self * P.id_stride + u;   // slot of the key in the forests / flag arrays
        const uint32_t c = __ldcs(P.ucount + u);
        if (P.method == METHOD_DIRECTIONAL) {
            if (c >= 2) sel = !P.dominated[u];
            else if (P.dead[u]) sel = false;
            else if (!P.linked[u]) sel = true;
            else sel = !P.deadroot[P.root_of[u]] && P.best[P.root_of[u]] == P.loc_of[u];
        } else if (P.method == METHOD_HIGHEST) {
            sel = !P.linked[u] || P.best[P.root_of[u]] == P.loc_of[u];
        } else {
            sel = P.state[gid] == 1;
        }
        P.selected[u] = sel ? 1 : 0;
        if (sel) {
            const uint32_t t = __ldcs(P.ufirst + u) - P.bit_base;
            atomicOr(P.bitmap + (t >> 5), 1u << (t & 31));
        }
    }
    block_add(sel ? 1u : 0u, &P.ctr->n_selected);
}

// Send buffer of the counter exchange: G blocks of (nper fill counters of the receiver's tiles, this rank's current
// edge count) -- the receiver learns, for free, how far every list is complete at this point of the job.
static __global__ void __launch_bounds__(256) pack_counters_kernel(const uint32_t *__restrict__ cursor, uint32_t nper, uint32_t G,
                                                                   const uint32_t *__restrict__ extra, uint32_t *__restrict__ out)
{
    const uint32_t i = blockIdx.x * 256u + threadIdx.x, block = nper + 1;
    if (i >= G * block) return;
    const uint32_t g = i / block, j = i % block;
    out[i] = j < nper ? cursor[(size_t)g * nper + j] : *extra;
}

static __global__ void copy_strided_kernel(uint32_t *dst, const uint32_t *src, uint32_t stride, uint32_t n)
{
    if (threadIdx.x < n) dst[threadIdx.x] = src[(size_t)threadIdx.x * stride];
}

// Keep bitmap of this rank's records [lo, lo + n): OR of that bit range of every rank's job-wide bitmap (fetched
// from the peers' HBM); output bit t = job bit (lo - bit_base + t).
struct MergeParams {
    const uint32_t *src[MAX_RANKS];
    uint32_t G;
    uint32_t bit_lo;      // lo - bit_base
    uint32_t n;           // records of this rank
    uint32_t src_words;   // words of a job-wide bitmap
    uint32_t *dst;
};
static __global__ void __launch_bounds__(256) merge_bitmaps_kernel(const __grid_constant__ MergeParams P)
{
    const uint32_t w = blockIdx.x * 256u + threadIdx.x, nw = (P.n + 31) / 32;
    if (w >= nw) return;
    const uint32_t k = (P.bit_lo >> 5) + w, sh = P.bit_lo & 31u;
    uint32_t acc = 0;
    for (uint32_t g = 0; g < P.G; g++) {
        const uint32_t lo = __ldcs(P.src[g] + k);
        const uint32_t hi = (sh && k + 1 < P.src_words) ? __ldcs(P.src[g] + k + 1) : 0u;
        acc |= sh ? ((lo >> sh) | (hi << (32u - sh))) : lo;
    }
    const uint32_t rem = P.n - 32u * w;
    if (rem < 32u) acc &= (1u << rem) - 1u;
    P.dst[w] = acc;
}

// Oversize tiles of a sharded job: the fragments of the listed owned tiles (all ranks' regions) go through the
// single-table insert on the owner.  grid.x = n_listed * G * (TILE_R / 256).
template <int K, int PW>
static __global__ void __launch_bounds__(256) spill_insert_tiles_kernel(const __grid_constant__ TileSource S,
                                                                        const uint32_t *__restrict__ list,
                                                                        const __grid_constant__ TableRef tab, uint32_t *n_claimed)
{
    constexpr int KW = K * PW;
    constexpr uint32_t BPP = TILE_R / 256;
    const uint32_t j = list[blockIdx.x / (S.G * BPP)], g = (blockIdx.x / BPP) % S.G;
    const uint32_t i = (blockIdx.x % BPP) * 256u + threadIdx.x;
    const uint32_t n = min(S.cnt[(size_t)g * S.cnt_stride + j], S.region);
    uint32_t claimed = NO_CLAIM;
    if (i < n) {
        uint32_t e[PART_RW];
        load_rec_stream(S.buf[g] + ((size_t)(S.first_tile + j) * S.region + i) * PART_RW, e);
        Key<K, PW> key;
#pragma unroll
        for (int w = 0; w < KW; w++) key.w[w] = e[w];
        claimed = table_insert<K, PW>(tab, key, hash_key(key), e[KW + 1], e[KW]);
    }
    const uint32_t pos = block_reserve(claimed != NO_CLAIM, n_claimed);
    if (claimed != NO_CLAIM) tab.uslot[pos] = claimed;
}

// ... and the records a rank had to spill (its own region of a tile was full): every owner scans every rank's
// spill buffer and takes the records of its tiles.
template <int K, int PW>
static __global__ void __launch_bounds__(256) spill_insert_owned_kernel(const uint32_t *__restrict__ spill, uint32_t n,
                                                                        uint32_t part_blocks, uint32_t d1, uint32_t nparts,
                                                                        uint32_t first_tile, uint32_t ntiles, uint32_t max_len,
                                                                        uint32_t pad_code, int varlen,
                                                                        const __grid_constant__ TableRef tab, uint32_t *n_claimed)
{
    constexpr int KW = K * PW;
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    uint32_t claimed = NO_CLAIM;
    if (i < n) {
        uint32_t e[PART_RW];
        load_rec_stream(spill + (size_t)i * PART_RW, e);
        Key<K, PW> key;
#pragma unroll
        for (int w = 0; w < KW; w++) key.w[w] = e[w];
        const uint32_t len = varlen ? key_length(key, pad_code, max_len) : max_len;
        const uint64_t h = part_blocks ? block0_hash(key, block_start(len, 1, d1), PART_SALT | len) : hash_key(key);
        const uint32_t t = part_of(h, nparts);
        if (t >= first_tile && t - first_tile < ntiles) claimed = table_insert<K, PW>(tab, key, hash_key(key), e[KW + 1], e[KW]);
    }
    const uint32_t pos = block_reserve(claimed != NO_CLAIM, n_claimed);
    if (claimed != NO_CLAIM) tab.uslot[pos] = claimed;
}

// adjacency runs several rounds over the (higher, lower) edges of all ranks: fetch them from the peers once
static __global__ void __launch_bounds__(256) adjacency_copy_kernel(const __grid_constant__ EdgeSource E, uint2 *dst,
                                                                    const uint32_t *__restrict__ offsets)
{
    for (uint32_t g = 0; g < E.G; g++) {
        const uint32_t n = min(E.n_edges[(size_t)g * E.n_stride], E.cap[g]);
        const uint32_t off = offsets[g];
        for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < n; i += gridDim.x * 256u) {
            const uint2 ed = __ldcs(E.edges[g] + i);
            dst[off + i] = make_uint2(slot_of_id(ed.x, E.G, E.id_stride), slot_of_id(ed.y, E.G, E.id_stride));
        }
    }
}

}  // namespace fqd
