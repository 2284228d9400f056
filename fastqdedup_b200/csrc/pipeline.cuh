// pipeline.cuh -- device kernels of the clustering path (sm_100a): the single-table plan (small
// jobs, long keys, Levenshtein passes, fallback), the dissection, the sharding helpers, and the
// shared pieces of the streaming plan, whose kernels are in partitioned.cuh (included at the end).
//
// Stages of the single-table plan (DESIGN.md has the data layout and the per-kernel roofline):
//   ingest   : per record quality filter (bit-exact double sum), bit-plane packing and
//              exact dedupe with counts + first index in an open-addressing table in HBM
//              (replaces Trie.add_sequence, reference _triemodule.c:222-288, and
//              average_error_rate, _fastqmodule.c:38-76)
//   passes   : pigeonhole block bucketing (counting sort of 8-byte {tag, uid} entries by
//              block hash) + in-bucket verification with XOR/popc Hamming or Myers
//              bit-vector Levenshtein; every verified pair is an edge
//              (replaces TrieNode_FindNearest / Trie.pop_cluster, _triemodule.c:380-495,
//              :778-897, and distances.h)
//   select   : lock-free union-find components + the three dissections in closed /
//              round-based form (replaces __init__.py:60-122)
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "key.cuh"
#define FQD_LUT_QUALIFIER static __constant__
#include "phred_lut.h"

namespace cg = cooperative_groups;

namespace fqd {

constexpr uint32_t SLOT_EMPTY = 0xFFFFFFFFu;
constexpr uint32_t SLOT_LOCKED = 0xFFFFFFFEu;
constexpr uint32_t ENT_LAST = 0x80000000u;
constexpr uint32_t ENT_BUILD = 0x40000000u;
constexpr uint32_t ENT_UID = 0x3FFFFFFFu;
constexpr uint32_t RANK_INVALID = 0xFFFFFFFFu;

constexpr int METHOD_HIGHEST = 0, METHOD_ADJACENCY = 1, METHOD_DIRECTIONAL = 2;

__host__ __device__ constexpr int round_up4(int x) { return (x + 3) & ~3; }

struct DevCounters {
    unsigned long long phred_err;     // min over (record << 8 | byte); ~0 = none
    unsigned long long n_candidates;
    unsigned long long n_edges;
    unsigned long long sum_weights;   // kept reads when records carry multiplicities
    uint32_t n_unique;
    uint32_t n_discarded;
    uint32_t n_merges;
    uint32_t n_selected;
    uint32_t table_full;
    uint32_t undecided;
    uint32_t unknown[8];              // bitmap of key bytes outside the alphabet
    uint32_t len_min, len_max;
    // n_candidates / n_merges contributions of the tile kernels, spread to avoid one hot address;
    // fetch_counters folds them into the two totals
    unsigned long long cand_spread[64];
    uint32_t merge_spread[64];
};
constexpr uint32_t STAT_SPREAD = 64;

// ---- small device helpers -----------------------------------------------------------

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// one atomicAdd per converged group instead of one per thread
__device__ __forceinline__ uint32_t aggregated_inc(uint32_t *ctr)
{
    cg::coalesced_group g = cg::coalesced_threads();
    uint32_t base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(ctr, g.size());
    return g.shfl(base, 0) + g.thread_rank();
}
__device__ __forceinline__ unsigned long long aggregated_inc64(unsigned long long *ctr)
{
    cg::coalesced_group g = cg::coalesced_threads();
    unsigned long long base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(ctr, (unsigned long long)g.size());
    return g.shfl(base, 0) + g.thread_rank();
}

// A single hot address takes ~0.5 ns per atomic on B200: 36 M appends through one counter cost
// more than the whole table probe.  These helpers issue ONE atomic per 256-thread block.
// Every thread of the block must call them (they contain __syncthreads).
__device__ __forceinline__ uint32_t block_reserve(bool flag, uint32_t *counter)
{
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_base;
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, flag);
    if (lane == 0) s_warp[wid] = __popc(bal);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t total = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) { const uint32_t c = s_warp[i]; s_warp[i] = total; total += c; }
        s_base = total ? atomicAdd(counter, total) : 0u;
    }
    __syncthreads();
    const uint32_t pos = s_base + s_warp[wid] + __popc(bal & ((1u << lane) - 1u));
    __syncthreads();   // s_warp / s_base may be reused by the next call
    return pos;
}

__device__ __forceinline__ void block_add(uint32_t v, uint32_t *counter)
{
    __shared__ uint32_t s_sum;
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_sum, v);
    __syncthreads();
    if (threadIdx.x == 0 && s_sum) atomicAdd(counter, s_sum);
    __syncthreads();
}

__device__ __forceinline__ void block_add64(uint32_t v, unsigned long long *counter)
{
    __shared__ uint32_t s_sum64;
    if (threadIdx.x == 0) s_sum64 = 0;
    __syncthreads();
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_sum64, v);
    __syncthreads();
    if (threadIdx.x == 0 && s_sum64) atomicAdd(counter, (unsigned long long)s_sum64);
    __syncthreads();
}

template <int K, int PW>
__device__ __forceinline__ void load_key(const uint32_t *__restrict__ ukey, uint32_t u,
                                         Key<K, PW> &k)
{
    const uint32_t *p = ukey + (size_t)u * (K * PW);
#pragma unroll
    for (int i = 0; i < K * PW; i++) k.w[i] = __ldg(p + i);
}

// streaming variant for kernels that read every key once (keeps the bucket counters in L2)
template <int K, int PW>
__device__ __forceinline__ void load_key_stream(const uint32_t *__restrict__ ukey, uint32_t u,
                                                Key<K, PW> &k)
{
    const uint32_t *p = ukey + (size_t)u * (K * PW);
    if constexpr ((K * PW) % 2 == 0) {
        const uint2 *p2 = reinterpret_cast<const uint2 *>(p);
#pragma unroll
        for (int i = 0; i < K * PW / 2; i++) { const uint2 v = __ldcs(p2 + i); k.w[2 * i] = v.x; k.w[2 * i + 1] = v.y; }
    } else {
#pragma unroll
        for (int i = 0; i < K * PW; i++) k.w[i] = __ldcs(p + i);
    }
}

// ---- partition buffers of the streaming plan (partitioned.cuh) ---------------------------------

constexpr int PART_RW = 8;                 // the plan handles 32-byte records (keys up to 6 words)
constexpr int TILE_R = 512;                // records per partition region = one shared-memory tile
constexpr uint32_t WARP_FULL = 0xFFFFFFFFu;
// salt of the hash that spreads records over the tiles by their pigeonhole block 0: it must not be the
// pass-0 signature itself, or the uniques of a few tiles would land in a few tiles again when a pass
// re-partitions just them (the spill path)
constexpr uint64_t PART_SALT = 0x5bd1e995ull << 32;

struct PartParams {
    uint32_t *buf;         // nparts regions of TILE_R records of PART_RW words; null = not partitioning
    uint32_t *cursor;      // nparts fill counters (a counter may exceed TILE_R: the partition is "oversize")
    uint32_t nparts;
    uint32_t *spill;       // records that did not fit their region (oversize partitions only)
    uint32_t *spill_cnt;   // [0] records appended to spill, [1] set when spill itself overflowed
    uint32_t spill_cap;
    // records per region.  One GPU: TILE_R (a region is a whole tile).  Sharded over G ranks a rank's region only
    // ever receives ~1/G of a tile, so it is sized for that (dense buffers: 1/G of the footprint, and the owner's
    // peer fetch reads a contiguous run instead of the head of a mostly empty 16 KB region)
    uint32_t region = TILE_R;
};

__device__ __forceinline__ uint32_t part_of(uint64_t h, uint32_t nparts)
{
    return __umulhi((uint32_t)(h >> 32), nparts);   // top hash bits; the tile's table uses the low ones
}

__device__ __forceinline__ void load_rec_stream(const uint32_t *p, uint32_t (&w)[PART_RW])
{
    asm volatile("ld.global.cs.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                 : "l"(p) : "memory");
}
__device__ __forceinline__ void store_rec_stream(uint32_t *p, const uint32_t (&w)[PART_RW])
{
    asm volatile("st.global.cs.v8.b32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};"
                 :: "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "l"(p)
                 : "memory");
}

// Append one 32-byte record to partition `part`: one atomic on the partition's cursor and one
// 256-bit store; the L2 assembles the nparts write streams into full lines.
__device__ __forceinline__ void part_place(const PartParams &Q, uint32_t part, uint32_t pos, const uint32_t (&e)[PART_RW]);
__device__ __forceinline__ void part_append(const PartParams &Q, uint32_t part, const uint32_t (&e)[PART_RW])
{
    part_place(Q, part, atomicAdd(Q.cursor + part, 1u), e);
}
// the store half: `pos` is what the cursor atomic returned
__device__ __forceinline__ void part_place(const PartParams &Q, uint32_t part, uint32_t pos, const uint32_t (&e)[PART_RW])
{
    if (pos < Q.region) {
        store_rec_stream(Q.buf + ((size_t)part * Q.region + pos) * PART_RW, e);
    } else if (Q.spill) {   // (without a spill buffer the tile kernel sees cursor > TILE_R and flags the overflow)
        const uint32_t sp = atomicAdd(Q.spill_cnt, 1u);
        if (sp < Q.spill_cap) store_rec_stream(Q.spill + (size_t)sp * PART_RW, e);
        else Q.spill_cnt[1] = 1u;   // extreme skew: the caller falls back to the other plan
    }
}

// ---- ingest: filter + pack + exact dedupe ----------------------------------------------

// The exact-dedupe table in HBM.
struct TableRef {
    uint32_t *table;           // open addressing: {key[KW], count, state} per slot;
                               // state = EMPTY | LOCKED | smallest record index seen so far
    uint64_t capacity;
    uint32_t *uslot;           // claimed slots in claim order (= unique ids)
    DevCounters *ctr;
};

struct IngestParams {
    uint64_t n;
    const uint8_t *keys;
    const uint64_t *key_off;
    const uint32_t *key_lens;
    uint32_t key_stride, key_len;
    const uint8_t *quals;
    const uint64_t *qual_off;
    const uint32_t *qual_lens;
    uint32_t qual_stride, qual_len;
    uint32_t max_len;          // padded key length
    uint32_t stage_bytes;      // >0: fixed-stride rows are staged through shared memory
    int filter_on;
    int phase;                 // 0: insert kept records, 1: first-index fix-up of discarded ones
    double max_err;
    uint32_t phred_offset;
    uint32_t pad_code;
    TableRef tab;
    PartParams part;           // part.buf != null: append {key, weight, index} to the partition of the
                               // key hash instead of inserting into `tab` (streaming plan)
    uint32_t part_blocks;      // > 0: partition by the hash of pigeonhole block 0 of part_blocks instead
                               // (keys sharing that block share a tile: pass 0 runs inside the dedupe tiles)
    uint32_t *keepmask;        // bit per record: passed the filter (null in partition mode)
    const uint32_t *weights;   // optional multiplicity per record
    uint32_t index_base;       // global index of record 0 of this shard
    int sharded;               // 1: filtered records are inserted with weight 0 (their first
                               //    index must reach the key's owner rank), no fix-up phase
    DevCounters *ctr;
    Codec codec;
    uint64_t plane_words = 0;  // partition_planes_kernel: 64-bit words per plane stream (`keys` points at plane 0)
};

// One probe reads a whole table record.  A 32-byte record is exactly one L2 sector and is
// fetched by ONE 256-bit load (LDG.E.ENL2.256.STRONG.GPU): the state word and the key words
// come from the same sector snapshot, so a published state implies the published key.
// Larger records read the state word first and the key words after it (the loads are
// control-dependent on the state, and the writer fences between key and state).
template <int RW>
__device__ __forceinline__ void load_record256(const uint32_t *rec, uint32_t (&w)[RW])
{
    static_assert(RW == 8, "256-bit record load");
    asm volatile("ld.relaxed.gpu.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                 : "l"(rec) : "memory");
}
__device__ __forceinline__ uint4 ld_relaxed_v4(const uint32_t *p)
{
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

// Slot layout: KW key words, count, state; padded to a multiple of 4 words (32 bytes = one
// L2 sector for the 36- and 48-nt configs).
__host__ __device__ constexpr int slot_words(int kw) { return round_up4(kw + 2); }

// Whole slot in as few loads as its size allows (see load_record256).
template <int RW>
__device__ __forceinline__ void load_slot(const uint32_t *rec, uint32_t (&w)[RW])
{
    if constexpr (RW == 8) {
        load_record256<RW>(rec, w);
    } else {
#pragma unroll
        for (int c = 0; c < RW / 4; c++) {
            const uint4 v = ld_relaxed_v4(rec + 4 * c);
            w[4 * c] = v.x; w[4 * c + 1] = v.y; w[4 * c + 2] = v.z; w[4 * c + 3] = v.w;
        }
    }
}

// Exact dedupe with counts (replaces Trie.add_sequence, reference _triemodule.c:222-288).
// The first record of a key claims a slot (CAS EMPTY -> LOCKED), writes key and count and
// publishes its record index in the state word.  Every later record of the key finds the
// slot with one sector read and adds to the count in that same (L2-hot) sector; the state
// word keeps the smallest record index, which is rarely lowered because records arrive
// roughly in index order.
constexpr uint32_t NO_CLAIM = 0xFFFFFFFFu;

// Returns the slot this call claimed (the record was the first of its key) or NO_CLAIM; the
// caller appends claimed slots to the unique list with one atomic per block.
template <int K, int PW>
__device__ __forceinline__ uint32_t table_insert(const TableRef &P, const Key<K, PW> &key,
                                                 uint64_t h, uint32_t t, uint32_t weight)
{
    constexpr int KW = K * PW, RW = slot_words(KW);
    uint64_t s = __umul64hi(h, P.capacity);
    for (uint64_t probes = 0; probes < P.capacity;) {
        uint32_t *rec = P.table + s * RW;
        uint32_t w[RW];
        uint32_t st;
        if constexpr (RW <= 8) {
            load_slot<RW>(rec, w);      // one sector: state and key from the same snapshot
            st = w[KW + 1];
        } else {
            st = ld_relaxed_u32(rec + KW + 1);
        }
        if (st == SLOT_EMPTY) {
            const uint32_t old = atomicCAS(rec + KW + 1, SLOT_EMPTY, SLOT_LOCKED);
            if (old == SLOT_EMPTY) {
                // (a single 256-bit store of the whole sector was measured 20 % slower than
                // scalar stores + release on B200: it collides with the CAS in flight)
#pragma unroll
                for (int i = 0; i < KW; i++) rec[i] = key.w[i];
                rec[KW] = weight;
                st_release_u32(rec + KW + 1, t);   // orders the stores above before the index
                return (uint32_t)s;
            }
            continue;   // somebody else claimed it first: look at the slot again
        }
        if (st == SLOT_LOCKED) continue;   // being published right now: look again
        if constexpr (RW > 8) load_slot<RW>(rec, w);
        uint32_t diff = 0;
#pragma unroll
        for (int i = 0; i < KW; i++) diff |= w[i] ^ key.w[i];
        if (diff == 0) {
            atomicAdd(rec + KW, weight);
            if (t < st) atomicMin(rec + KW + 1, t);
            return NO_CLAIM;
        }
        s = (s + 1 == P.capacity) ? 0 : s + 1;
        probes++;
    }
    P.ctr->table_full = 1u;
    return NO_CLAIM;
}

// Records that failed the filter still define "first occurrence" (pass 2 of the reference
// does not re-apply the filter, __init__.py:201-206): lower `first` of an existing key.
template <int K, int PW>
__device__ __forceinline__ void table_touch_first(const TableRef &P, const Key<K, PW> &key,
                                                  uint64_t h, uint32_t t)
{
    constexpr int KW = K * PW, RW = slot_words(KW);
    uint64_t s = __umul64hi(h, P.capacity);
    for (uint64_t probes = 0; probes < P.capacity; probes++) {
        uint32_t *rec = P.table + s * RW;
        uint32_t w[RW];
        load_slot<RW>(rec, w);
        const uint32_t st = w[KW + 1];
        if (st == SLOT_EMPTY) return;
        uint32_t diff = 0;
#pragma unroll
        for (int i = 0; i < KW; i++) diff |= w[i] ^ key.w[i];
        if (diff == 0) {
            if (t < st) atomicMin(rec + KW + 1, t);
            return;
        }
        s = (s + 1 == P.capacity) ? 0 : s + 1;
    }
}

// cooperative copy of `bytes` bytes into shared memory (16-byte loads when aligned)
__device__ __forceinline__ void stage_rows(uint8_t *stage, const uint8_t *src, uint64_t bytes, int tid)
{
    if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
        const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
        uint4 *d4 = reinterpret_cast<uint4 *>(stage);
        for (uint32_t i = tid; i < bytes / 16; i += 256) d4[i] = __ldg(s4 + i);
        for (uint32_t i = (uint32_t)(bytes & ~15ull) + tid; i < bytes; i += 256) stage[i] = __ldg(src + i);
    } else {
        for (uint32_t i = tid; i < bytes; i += 256) stage[i] = __ldg(src + i);
    }
}

// Records per thread.  2 (pack the second key while the first key's prefetched table sector is
// on its way from HBM) was measured 26 % SLOWER on B200 than 1: the kernel lives on occupancy
// (32 registers, 8 blocks/SM), not on per-thread memory-level parallelism.
constexpr int INGEST_ROWS = 1;

template <int K, int PW>
static __global__ void __launch_bounds__(256) ingest_kernel(const __grid_constant__ IngestParams P)
{
    constexpr int ROWS = INGEST_ROWS;
    constexpr int KW = K * PW, RW = slot_words(KW);
    extern __shared__ __align__(16) uint8_t smem[];
    double *lut_d = reinterpret_cast<double *>(smem);   // 128 doubles
    uint8_t *lut = smem + 1024;                         // 256 bytes
    uint8_t *stage = smem + 1280;                       // stage_bytes
    const int tid = threadIdx.x;
    if (tid < 128) lut_d[tid] = __longlong_as_double((long long)FQD_PHRED_LUT_BITS[tid]);
    lut[tid] = P.codec.lut[tid];
    const uint64_t t0 = (uint64_t)blockIdx.x * (256u * ROWS);
    const uint32_t nblk = (uint32_t)min((uint64_t)(256u * ROWS), P.n - t0);
    uint64_t t[ROWS];
    bool active[ROWS], keep[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
        const uint32_t lr = r * 256u + tid;
        t[r] = t0 + lr;
        active[r] = lr < nblk;
        keep[r] = true;
        if (P.phase == 1) keep[r] = active[r] && !((P.keepmask[t[r] >> 5] >> (t[r] & 31)) & 1u);
    }
    __syncthreads();

    // ---- quality filter (reference _fastqmodule.c:58-75) ----
    if (P.filter_on && P.phase == 0) {
        const bool staged = !P.qual_off && P.stage_bytes;
        if (staged) {
            stage_rows(stage, P.quals + t0 * P.qual_stride, (uint64_t)nblk * P.qual_stride, tid);
            __syncthreads();
        }
#pragma unroll
        for (int r = 0; r < ROWS; r++) {
            if (!active[r]) continue;
            const uint8_t *q;
            uint32_t qlen;
            if (P.qual_off) {
                q = P.quals + P.qual_off[t[r]];
                qlen = (uint32_t)(P.qual_off[t[r] + 1] - P.qual_off[t[r]]);
            } else {
                q = staged ? stage + (size_t)(r * 256u + tid) * P.qual_stride : P.quals + t[r] * P.qual_stride;
                qlen = P.qual_lens ? P.qual_lens[t[r]] : P.qual_len;
            }
            double total = 0.0;
            const uint32_t max_score = (126u - P.phred_offset) & 0xFFu;
            bool bad = false;
            for (uint32_t i = 0; i < qlen; i++) {
                const uint32_t c = q[i];
                const uint32_t score = (c - P.phred_offset) & 0xFFu;   // uint8 wrap (:62)
                if (score > max_score) {
                    atomicMin(&P.ctr->phred_err, (unsigned long long)(((t[r] + P.index_base) << 8) | c));
                    bad = true;
                    break;
                }
                total = __dadd_rn(total, lut_d[score]);                // left to right (:72)
            }
            const double avg = __ddiv_rn(total, (double)qlen);         // (:74), 0/0 = NaN
            keep[r] = !bad && !(avg > P.max_err);                      // strict; NaN keeps
        }
        __syncthreads();   // stage is reused for the keys
    }
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
        if (!active[r]) keep[r] = false;
        if (P.phase == 0) {
            const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, keep[r]);
            if ((tid & 31) == 0 && t[r] < P.n && P.keepmask) P.keepmask[t[r] >> 5] = ballot;
            if (P.filter_on && active[r] && !keep[r]) aggregated_inc(&P.ctr->n_discarded);
        }
    }

    // ---- pack (bit planes), hash, prefetch the home slot ----
    const bool kstaged = !P.key_off && P.stage_bytes;
    if (kstaged) {
        stage_rows(stage, P.keys + t0 * P.key_stride, (uint64_t)nblk * P.key_stride, tid);
        __syncthreads();
    }
    Key<K, PW> key[ROWS];
    uint64_t hash[ROWS];
    uint32_t klens[ROWS];
    bool go[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
        klens[r] = 0;
        // phase 0: a filtered record stops here (unless sharded); phase 1: `keep` means "was filtered"
        go[r] = active[r] && (keep[r] || ((P.sharded || P.part.buf) && P.phase == 0));
        if (!go[r]) continue;
        const uint8_t *kb;
        uint32_t klen;
        if (P.key_off) {
            kb = P.keys + P.key_off[t[r]];
            klen = (uint32_t)(P.key_off[t[r] + 1] - P.key_off[t[r]]);
        } else {
            kb = kstaged ? stage + (size_t)(r * 256u + tid) * P.key_stride : P.keys + t[r] * P.key_stride;
            klen = P.key_lens ? P.key_lens[t[r]] : P.key_len;
        }
        if (klen > P.max_len) klen = P.max_len;
        klens[r] = klen;
        uint32_t badbyte = 0;
        bool packed = false;
        if constexpr (K == 3) {
            // table-free DNA packing from the 4-byte aligned staged row (shared-memory loads)
            if (P.codec.swar && kstaged && (P.key_stride & 3u) == 0) {
                const uint32_t *row = reinterpret_cast<const uint32_t *>(stage) + (size_t)(r * 256u + tid) * (P.key_stride >> 2);
                const uint32_t nw = (!P.key_lens && !P.codec.varlen && (klen & 3u) == 0) ? klen >> 2 : 0u;   // uniform
                if (nw == 9 && PW >= 2) { if constexpr (PW >= 2) packed = pack_key_acgtn_fixed<PW, 9>(row, key[r]); }
                else if (nw == 12 && PW >= 2) { if constexpr (PW >= 2) packed = pack_key_acgtn_fixed<PW, 12>(row, key[r]); }
                else if (nw == 6) packed = pack_key_acgtn_fixed<PW, 6>(row, key[r]);
                else if (nw == 3) packed = pack_key_acgtn_fixed<PW, 3>(row, key[r]);
                else packed = pack_key_acgtn<PW>(row, klen, P.codec.varlen ? P.max_len : klen, key[r]);
            }
        }
        if (!packed &&
            !pack_key<K, PW>(kb, klen, P.codec.varlen ? P.max_len : klen, lut, P.pad_code, key[r], &badbyte)) {
            // report every unknown byte of this key so one retry with a grown alphabet suffices
            for (uint32_t i = 0; i < klen; i++) {
                const uint32_t c = kb[i];
                if (lut[c] == 0xFF) atomicOr(&P.ctr->unknown[c >> 5], 1u << (c & 31));
            }
            go[r] = false;
            continue;
        }
        // (records partitioned by their pigeonhole block 0 never need the hash of the whole key)
        hash[r] = (P.part.buf && P.part_blocks) ? block0_hash(key[r], block_start(klen, 1, P.part_blocks), PART_SALT | klen)
                                                : hash_key(key[r]);
        if constexpr (ROWS > 1) {
            const uint32_t *home = P.tab.table + __umul64hi(hash[r], P.tab.capacity) * RW;
            asm volatile("prefetch.global.L2 [%0];" :: "l"(home));
        }
    }

    // ---- streaming plan: hand the record to the partition of its hash ----
    if (P.part.buf) {
        if constexpr (RW == PART_RW) {
#pragma unroll
            for (int r = 0; r < ROWS; r++) {
                if (!go[r]) continue;
                // a filtered record travels with weight 0: it may hold its key's first index
                const uint32_t weight = keep[r] ? (P.weights ? P.weights[t[r]] : 1u) : 0u;
                if (P.weights && keep[r]) atomicAdd(&P.ctr->sum_weights, (unsigned long long)weight);
                uint32_t e[RW];
#pragma unroll
                for (int i = 0; i < RW; i++) e[i] = 0;
#pragma unroll
                for (int i = 0; i < KW; i++) e[i] = key[r].w[i];
                e[KW] = weight;
                e[KW + 1] = P.index_base + (uint32_t)t[r];
                part_append(P.part, part_of(hash[r], P.part.nparts), e);
            }
        }
        return;
    }

    // ---- exact dedupe ----
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
        uint32_t claimed = NO_CLAIM;
        if (go[r]) {
            const uint32_t tg = P.index_base + (uint32_t)t[r];
            if (P.phase == 0) {
                // a filtered record only gets here in sharded mode: weight 0 carries its first index
                const uint32_t weight = keep[r] ? (P.weights ? P.weights[t[r]] : 1u) : 0u;
                if (P.weights && keep[r]) atomicAdd(&P.ctr->sum_weights, (unsigned long long)weight);
                claimed = table_insert<K, PW>(P.tab, key[r], hash[r], tg, weight);
            } else {
                table_touch_first<K, PW>(P.tab, key[r], hash[r], tg);
            }
        }
        if (P.phase == 0) {   // uniform: the claimed slots of the block join the unique list
            const uint32_t pos = block_reserve(claimed != NO_CLAIM, &P.ctr->n_unique);
            if (claimed != NO_CLAIM) P.tab.uslot[pos] = claimed;
        }
    }
}

// min / max of the key lengths (decides PW and whether PAD is needed)
static __global__ void length_range_kernel(uint64_t n, const uint64_t *off, const uint32_t *lens,
                                    DevCounters *ctr)
{
    uint32_t lo = 0xFFFFFFFFu, hi = 0;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n;
         t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t l = off ? (uint32_t)(off[t + 1] - off[t]) : lens[t];
        lo = min(lo, l);
        hi = max(hi, l);
    }
    for (int o = 16; o; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xFFFFFFFFu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&ctr->len_min, lo);
        atomicMax(&ctr->len_max, hi);
    }
}

// ---- gather: claimed slots -> dense unique arrays + forest init --------------------------------

template <int K, int PW>
static __global__ void __launch_bounds__(256) gather_kernel(uint32_t U, const uint32_t *__restrict__ table,
                                                     const uint32_t *__restrict__ uslot,
                                                     uint32_t *__restrict__ ukey,
                                                     uint32_t *__restrict__ ucount,
                                                     uint32_t *__restrict__ ufirst,
                                                     uint32_t *__restrict__ parent_a,
                                                     uint32_t *__restrict__ parent_b,
                                                     uint32_t *__restrict__ best)
{
    constexpr int KW = K * PW, RW = slot_words(KW);
    const uint32_t u = blockIdx.x * 256u + threadIdx.x;
    if (u >= U) return;
    const uint4 *rec = reinterpret_cast<const uint4 *>(table + (size_t)uslot[u] * RW);
    uint32_t w[RW];
#pragma unroll
    for (int c = 0; c < RW / 4; c++) {
        const uint4 v = __ldcs(rec + c);
        w[4 * c] = v.x; w[4 * c + 1] = v.y; w[4 * c + 2] = v.z; w[4 * c + 3] = v.w;
    }
#pragma unroll
    for (int i = 0; i < KW; i++) ukey[(size_t)u * KW + i] = w[i];
    ucount[u] = w[KW];
    ufirst[u] = w[KW + 1];
    if (parent_a) parent_a[u] = u;
    if (parent_b) parent_b[u] = u;
    if (best) best[u] = u;
}

static __global__ void __launch_bounds__(256) init_forest_kernel(uint32_t U, uint32_t *parent_a,
                                                                 uint32_t *parent_b, uint32_t *best)
{
    const uint32_t u = blockIdx.x * 256u + threadIdx.x;
    if (u >= U) return;
    parent_a[u] = u;
    if (parent_b) parent_b[u] = u;
    if (best) best[u] = u;
}

static __global__ void __launch_bounds__(256) iota_kernel(uint32_t *p, uint32_t n)
{
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i < n) p[i] = i;
}

// ---- union-find (roots have the smallest id of their set => deterministic labels) ------

__device__ __forceinline__ uint32_t uf_find(uint32_t *parent, uint32_t x)
{
    uint32_t p = ld_relaxed_u32(parent + x);
    while (p != x) {
        const uint32_t gp = ld_relaxed_u32(parent + p);
        if (gp != p) atomicMin(parent + x, gp);   // path halving; parents only ever decrease
        x = p;
        p = gp;
    }
    return x;
}

__device__ __forceinline__ bool uf_union(uint32_t *parent, uint32_t a, uint32_t b)
{
    for (;;) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return false;
        if (a < b) { const uint32_t t = a; a = b; b = t; }   // hook the larger root under the smaller
        const uint32_t old = atomicCAS(parent + a, a, b);
        if (old == a) return true;
    }
}

// ---- pigeonhole passes -------------------------------------------------------------------

struct PassParams {
    uint32_t U;
    uint32_t u_lo;          // streaming plan: the pass covers the uniques [u_lo, U) only
    const uint32_t *ukey;
    const uint32_t *ucount;
    int d, edit, varlen, method;
    uint32_t max_len, pad_code;
    int pass_j, V;
    uint32_t fix_st, fix_bl;   // Hamming block of pass_j for keys of max_len symbols (set with pass_j)
    int my_rank, world;     // replicated-set plan: buckets are owned by rank (sig >> 32) % world
    // tile-sharded plan: ids are job-wide (local id * id_mul + id_add, ranks interleaved) and the consequences of
    // an edge travel in its state bits instead of being written to the flag bytes (partitioned.cuh, EDGE_*)
    uint32_t id_mul = 1, id_add = 0;
    int edge_flags = 0;
    const uint32_t *U_dev = nullptr;   // when set: the unique count on the device (U is then an upper bound)
    uint32_t nb_mask;
    uint32_t *cnt;          // NB+1 counters -> exclusive offsets after the scan
    uint32_t *fill;         // thin entries: 2 * NB cursors -- builds fill a bucket from the front, probes from the back
    uint32_t *rank;         // U*V
    uint2 *entries;         // thin {tag, uid|flags} entries (Levenshtein passes)
    uint32_t *fat;          // fat {key, count, uid|flags} entries (Hamming passes)
    uint32_t n_entries;
    uint32_t *parent_full;
    uint32_t *parent_one;
    uint8_t *dominated;
    uint8_t *dead;
    uint2 *edges;
    unsigned long long edge_cap;
    DevCounters *ctr;
    uint8_t rank_of_code[256];
};

// Variant v of unique `key` in pass j: which substring is hashed, and whether the entry
// is the key's own (canonical) block.  Hamming: one variant, block j of d+1.
// Levenshtein: the probe side enumerates every build length la = len - delta and every
// shift s the <= d edits allow: s in [-floor((d-delta)/2), floor((d+delta)/2)].
template <int K, int PW>
__device__ __forceinline__ bool pass_variant(const Key<K, PW> &key, uint32_t len,
                                             const PassParams &P, int v, uint64_t &sig,
                                             bool &build)
{
    const uint32_t nb = (uint32_t)P.d + 1u;
    const uint32_t j = (uint32_t)P.pass_j;
    if (!P.edit) {
        uint32_t st, bl;
        if (!P.varlen) {   // one length: the block bounds were computed on the host
            st = P.fix_st; bl = P.fix_bl;
        } else {
            st = block_start(len, j, nb);
            bl = block_start(len, j + 1, nb) - st;
        }
        sig = block_hash(key, st, bl, ((uint64_t)j << 32) | len);
        build = true;
        return true;
    }
    const int li = v / (P.d + 1), si = v % (P.d + 1);
    const int delta = P.varlen ? li - P.d : 0;
    const int la = (int)len - delta;
    if (la < 0 || la > (int)P.max_len) return false;
    const int smin = -((P.d - delta) / 2), smax = (P.d + delta) / 2;
    const int s = smin + si;
    if (s > smax) return false;
    const uint32_t st = block_start((uint32_t)la, j, nb);
    const uint32_t bl = block_start((uint32_t)la, j + 1, nb) - st;
    const int pos = (int)st + s;
    if (pos < 0 || pos + (int)bl > (int)len) return false;
    sig = block_hash(key, (uint32_t)pos, bl, ((uint64_t)j << 32) | (uint32_t)la);
    build = (delta == 0 && s == 0);
    return true;
}

template <int K, int PW>
static __global__ void __launch_bounds__(256) sig_count_kernel(const __grid_constant__ PassParams P)
{
    const uint32_t u = blockIdx.x * 256u + threadIdx.x;
    if (u >= P.U) return;
    Key<K, PW> key;
    load_key_stream<K, PW>(P.ukey, u, key);
    const uint32_t len = P.varlen ? key_length(key, P.pad_code, P.max_len) : P.max_len;
    for (int v = 0; v < P.V; v++) {
        uint64_t sig;
        bool build;
        uint32_t r = RANK_INVALID;
        if (pass_variant<K, PW>(key, len, P, v, sig, build) &&
            (P.world <= 1 || (uint32_t)(sig >> 32) % (uint32_t)P.world == (uint32_t)P.my_rank))
            r = atomicAdd(P.cnt + ((uint32_t)sig & P.nb_mask), 1u);
        __stcs(P.rank + (size_t)u * P.V + v, r);
    }
}

template <int K, int PW>
static __global__ void __launch_bounds__(256) scatter_kernel(const __grid_constant__ PassParams P)
{
    const uint32_t u = blockIdx.x * 256u + threadIdx.x;
    if (u >= P.U) return;
    Key<K, PW> key;
    load_key_stream<K, PW>(P.ukey, u, key);
    const uint32_t len = P.varlen ? key_length(key, P.pad_code, P.max_len) : P.max_len;
    for (int v = 0; v < P.V; v++) {
        const uint32_t r = __ldcs(P.rank + (size_t)u * P.V + v);
        if (r == RANK_INVALID) continue;
        uint64_t sig;
        bool build;
        pass_variant<K, PW>(key, len, P, v, sig, build);
        const uint32_t b = (uint32_t)sig & P.nb_mask;
        const uint32_t lo = P.cnt[b], hi = P.cnt[b + 1];
        // builds first: a pair needs at least one build entry, so behind a probe entry nothing of its bucket pairs
        // with it -- the compare kernels drop probe rows at once and the lanes of a warp that work are (nearly) all
        // build rows with every column behind them live (ncu before: 11 of 32 lanes per instruction)
        const uint32_t pos = build ? lo + atomicAdd(P.fill + 2u * b, 1u) : hi - 1u - atomicAdd(P.fill + 2u * b + 1u, 1u);
        uint32_t meta = u | (build ? ENT_BUILD : 0u) | (pos + 1 == hi ? ENT_LAST : 0u);
        P.entries[pos] = make_uint2((uint32_t)(sig >> 32), meta);
    }
}

template <int K, int PW>
__device__ __forceinline__ void process_edge(const PassParams &P, uint32_t ui, uint32_t uj,
                                             uint32_t ci, uint32_t cj, const Key<K, PW> &ki,
                                             const Key<K, PW> &kj, uint32_t &merges)
{
    if (uf_union(P.parent_full, ui, uj)) merges++;
    if (P.method == METHOD_DIRECTIONAL) {
        // closed form of reference __init__.py:60-91 (DESIGN.md "directional")
        if (ci >= 2 && (unsigned long long)cj >= 2ull * ci - 1ull) P.dominated[ui] = 1;
        if (cj >= 2 && (unsigned long long)ci >= 2ull * cj - 1ull) P.dominated[uj] = 1;
        if (ci == 1 && cj == 1) uf_union(P.parent_one, ui, uj);
        else if (ci == 1) P.dead[ui] = 1;
        else if (cj == 1) P.dead[uj] = 1;
    } else if (P.method == METHOD_ADJACENCY) {
        const bool i_less = prio_less<K, PW>(ci, ki, cj, kj, P.rank_of_code);
        const unsigned long long pos = aggregated_inc64(&P.ctr->n_edges);
        if (pos < P.edge_cap) P.edges[pos] = i_less ? make_uint2(uj, ui) : make_uint2(ui, uj);
    }
}

// "Levenshtein distance <= d" of two candidates of a bucket.  Keys of up to 63 symbols at d <= 3 take the
// furthest-reaching-diagonals form (key.cuh lv_within_d): exact, ~200 integer operations at d = 2 and no
// data-dependent loop, so a warp verifies 32 candidates in lockstep.  Longer keys: Myers' bit-vector algorithm
// behind the shifted-Hamming filter (almost every candidate of a block bucket is a chance hit; the filter turns
// most of them away for a seventh of the cost of the verify).
template <int K, int PW>
__device__ __forceinline__ bool verify_edit(const Key<K, PW> &ki, uint32_t li, const Key<K, PW> &kj, uint32_t lj, int d)
{
    if constexpr (PW <= 2) {
        if (d <= LV_MAX_D && li <= 63u && lj <= 63u) return edit_within<K, PW>(ki, li, kj, lj, d);
    }
    return (d > 4 || shifted_hamming_maybe_within<K, PW>(ki, li, kj, lj, d)) && myers_within<K, PW>(ki, li, kj, lj, d);
}

template <int K, int PW>
static __global__ void __launch_bounds__(256) compare_kernel(const __grid_constant__ PassParams P)
{
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    uint32_t merges = 0, cand = 0;
    if (i < P.cnt[P.nb_mask + 1]) {   // cnt[NB] = number of entries after the scan
        const uint2 e = P.entries[i];
        if (!(e.y & ENT_LAST) && (e.y & ENT_BUILD)) {   // (builds come first in a bucket: behind a probe entry only probes)
            const uint32_t ui = e.y & ENT_UID;
            bool loaded = false;
            Key<K, PW> ki;
            uint32_t li = 0, ci = 0;
            for (uint32_t j = i + 1;; j++) {
                const uint2 f = P.entries[j];
                if (f.x == e.x) {
                    const uint32_t uj = f.y & ENT_UID;
                    if (uj != ui && ((e.y | f.y) & ENT_BUILD)) {
                        if (!loaded) {
                            load_key<K, PW>(P.ukey, ui, ki);
                            li = P.varlen ? key_length(ki, P.pad_code, P.max_len) : P.max_len;
                            ci = P.ucount[ui];
                            loaded = true;
                        }
                        Key<K, PW> kj;
                        load_key<K, PW>(P.ukey, uj, kj);
                        cand++;
                        bool ok;
                        if (P.edit) {
                            const uint32_t lj = P.varlen ? key_length(kj, P.pad_code, P.max_len) : P.max_len;
                            ok = verify_edit<K, PW>(ki, li, kj, lj, P.d);
                        } else {
                            ok = hamming_within<K, PW>(ki, kj, P.d, P.varlen != 0, P.pad_code);
                        }
                        if (ok) process_edge<K, PW>(P, ui, uj, ci, P.ucount[uj], ki, kj, merges);
                    }
                }
                if (f.y & ENT_LAST) break;
            }
        }
    }
    // per-warp reduction of the two counters
    for (int o = 16; o; o >>= 1) {
        merges += __shfl_xor_sync(0xFFFFFFFFu, merges, o);
        cand += __shfl_xor_sync(0xFFFFFFFFu, cand, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (merges) atomicAdd(&P.ctr->n_merges, merges);
        if (cand) atomicAdd(&P.ctr->n_candidates, (unsigned long long)cand);
    }
}

// The same compare as dense tiles (the default for the Levenshtein passes; FQD_COMPARE_V1=1 selects the kernel above).
// Short pigeonhole blocks make big buckets -- config 4 at d = 2 has 8-symbol blocks, ~73 entries per bucket and 2e8
// candidate pairs -- and the one-thread-per-entry walk above then spends its time on a dependent key load per
// candidate and on warps whose lanes walk ranges of different lengths (31 ms for 1.6 M keys).  Here one warp owns 32
// consecutive entries (rows).  It streams the entries behind them in chunks of 32 columns: every lane fetches ONE
// column (entry + key, the only random loads: one per entry and chunk instead of one per candidate) into the warp's
// shared-memory tile, then all lanes test their row against column t, t = 0..31, read by broadcast.  A row is done
// at the ENT_LAST mark of its bucket; the warp stops when all its rows are.  Bucket boundaries only bound the work:
// a pair is reported iff the verify passes, so whatever else lies in a tile is harmless.
template <int KW> __host__ __device__ constexpr int dense_warps() { return KW <= 12 ? 8 : 4; }

template <int K, int PW>
static __global__ void __launch_bounds__(dense_warps<K * PW>() * 32) compare_dense_kernel(const __grid_constant__ PassParams P)
{
    constexpr int KW = K * PW, WARPS = dense_warps<KW>(), CW = (KW + 2) | 1;   // odd stride: conflict-free column writes
    __shared__ uint32_t col[WARPS][32][CW];
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const uint32_t E = P.cnt[P.nb_mask + 1];   // cnt[NB] = number of entries after the scan
    const uint32_t row0 = (blockIdx.x * WARPS + wid) * 32u;
    uint32_t merges = 0, cand = 0;
    if (row0 < E) {   // (warp-uniform)
        const uint32_t i = row0 + lane;
        const uint2 e = i < E ? P.entries[i] : make_uint2(0u, ENT_LAST);
        const uint32_t ui = e.y & ENT_UID;
        // nothing follows the last entry of a bucket, and (builds first) nothing behind a probe entry pairs with it
        bool ended = (e.y & ENT_LAST) != 0 || !(e.y & ENT_BUILD);
        Key<K, PW> ki;
#pragma unroll
        for (int w = 0; w < KW; w++) ki.w[w] = 0;
        if (i < E) load_key<K, PW>(P.ukey, ui, ki);
        const uint32_t li = P.varlen ? key_length(ki, P.pad_code, P.max_len) : P.max_len;
        uint32_t (*tile)[CW] = col[wid];
        for (uint32_t c0 = row0; __any_sync(WARP_FULL, !ended); c0 += 32u) {
            __syncwarp();
            if (c0 == row0) {
#pragma unroll
                for (int w = 0; w < KW; w++) tile[lane][w] = ki.w[w];
                tile[lane][KW] = e.x;
                tile[lane][KW + 1] = e.y;
            } else {
                const uint32_t j = c0 + lane;
                const uint2 f = j < E ? P.entries[j] : make_uint2(0u, ENT_LAST);   // past the end: stops every row
                Key<K, PW> kf;
#pragma unroll
                for (int w = 0; w < KW; w++) kf.w[w] = 0;
                if (j < E) load_key<K, PW>(P.ukey, f.y & ENT_UID, kf);
#pragma unroll
                for (int w = 0; w < KW; w++) tile[lane][w] = kf.w[w];
                tile[lane][KW] = f.x;
                tile[lane][KW + 1] = f.y;
            }
            __syncwarp();
#pragma unroll 2
            for (uint32_t t = 0; t < 32u; t++) {
                const uint32_t fx = tile[t][KW], fy = tile[t][KW + 1];
                const bool live = !ended && c0 + t > i;
                if (live && fx == e.x && (fy & ENT_UID) != ui && ((e.y | fy) & ENT_BUILD)) {
                    Key<K, PW> kj;
#pragma unroll
                    for (int w = 0; w < KW; w++) kj.w[w] = tile[t][w];
                    cand++;
                    bool ok;
                    if (P.edit) {
                        const uint32_t lj = P.varlen ? key_length(kj, P.pad_code, P.max_len) : P.max_len;
                        ok = verify_edit<K, PW>(ki, li, kj, lj, P.d);
                    } else {
                        ok = hamming_within<K, PW>(ki, kj, P.d, P.varlen != 0, P.pad_code);
                    }
                    if (ok) {
                        const uint32_t uj = fy & ENT_UID;
                        process_edge<K, PW>(P, ui, uj, P.ucount[ui], P.ucount[uj], ki, kj, merges);
                    }
                }
                if (live && (fy & ENT_LAST)) ended = true;
                if (!__any_sync(WARP_FULL, !ended)) break;
            }
        }
    }
    for (int o = 16; o; o >>= 1) {
        merges += __shfl_xor_sync(WARP_FULL, merges, o);
        cand += __shfl_xor_sync(WARP_FULL, cand, o);
    }
    if (lane == 0) {
        if (merges) atomicAdd(&P.ctr->n_merges, merges);
        if (cand) atomicAdd(&P.ctr->n_candidates, (unsigned long long)cand);
    }
}

// ---- Hamming passes: bucket entries carry the key --------------------------------------------
//
// With one variant per key (Hamming) the bucket-ordered array holds {key, count, uid|LAST}
// records of fat_words() 32-bit words (32 bytes for keys up to 6 words): the compare kernel
// then streams the array once and never dereferences a unique id.

__host__ __device__ constexpr int fat_words(int kw) { return round_up4(kw + 2); }

template <int K, int PW>
static __global__ void __launch_bounds__(256) scatter_fat_kernel(const __grid_constant__ PassParams P)
{
    constexpr int KW = K * PW, FW = fat_words(KW);
    const uint32_t u = blockIdx.x * 256u + threadIdx.x;
    if (u >= P.U) return;
    Key<K, PW> key;
    load_key_stream<K, PW>(P.ukey, u, key);
    const uint32_t len = P.varlen ? key_length(key, P.pad_code, P.max_len) : P.max_len;
    const uint32_t r = __ldcs(P.rank + u);
    if (r == RANK_INVALID) return;   // bucket owned by another rank
    uint64_t sig;
    bool build;
    pass_variant<K, PW>(key, len, P, 0, sig, build);
    const uint32_t b = (uint32_t)sig & P.nb_mask;
    const uint32_t lo = P.cnt[b], hi = P.cnt[b + 1];
    const uint32_t pos = lo + r;
    uint32_t e[FW];
#pragma unroll
    for (int i = 0; i < FW; i++) e[i] = 0;
#pragma unroll
    for (int i = 0; i < KW; i++) e[i] = key.w[i];
    e[KW] = __ldcs(P.ucount + u);
    e[KW + 1] = u | (pos + 1 == hi ? ENT_LAST : 0u);
    uint4 *dst = reinterpret_cast<uint4 *>(P.fat + (size_t)pos * FW);
#pragma unroll
    for (int c = 0; c < FW / 4; c++) __stcs(dst + c, make_uint4(e[4 * c], e[4 * c + 1], e[4 * c + 2], e[4 * c + 3]));
}

template <int K, int PW>
__device__ __forceinline__ void load_fat(const uint32_t *__restrict__ fat, uint32_t i, Key<K, PW> &k,
                                         uint32_t &count, uint32_t &meta)
{
    constexpr int KW = K * PW, FW = fat_words(KW);
    const uint4 *src = reinterpret_cast<const uint4 *>(fat + (size_t)i * FW);
    uint32_t e[FW];
#pragma unroll
    for (int c = 0; c < FW / 4; c++) {
        const uint4 v = __ldg(src + c);
        e[4 * c] = v.x; e[4 * c + 1] = v.y; e[4 * c + 2] = v.z; e[4 * c + 3] = v.w;
    }
#pragma unroll
    for (int j = 0; j < KW; j++) k.w[j] = e[j];
    count = e[KW];
    meta = e[KW + 1];
}

template <int K, int PW>
static __global__ void __launch_bounds__(256) compare_fat_kernel(const __grid_constant__ PassParams P)
{
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    uint32_t merges = 0, cand = 0;
    if (i < P.cnt[P.nb_mask + 1]) {
        Key<K, PW> ki;
        uint32_t ci, mi;
        load_fat<K, PW>(P.fat, i, ki, ci, mi);
        if (!(mi & ENT_LAST)) {
            const uint32_t ui = mi & ENT_UID;
            for (uint32_t j = i + 1;; j++) {
                Key<K, PW> kj;
                uint32_t cj, mj;
                load_fat<K, PW>(P.fat, j, kj, cj, mj);
                cand++;
                if (hamming_within<K, PW>(ki, kj, P.d, P.varlen != 0, P.pad_code))
                    process_edge<K, PW>(P, ui, mj & ENT_UID, ci, cj, ki, kj, merges);
                if (mj & ENT_LAST) break;
            }
        }
    }
    for (int o = 16; o; o >>= 1) {
        merges += __shfl_xor_sync(0xFFFFFFFFu, merges, o);
        cand += __shfl_xor_sync(0xFFFFFFFFu, cand, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (merges) atomicAdd(&P.ctr->n_merges, merges);
        if (cand) atomicAdd(&P.ctr->n_candidates, (unsigned long long)cand);
    }
}

// ---- exclusive scan over the bucket counters (3 kernels, 4096 items per block) ---------

constexpr int SCAN_ITEMS = 16, SCAN_THREADS = 256, SCAN_TILE = SCAN_ITEMS * SCAN_THREADS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total, uint32_t *warp_sums)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0;
        uint32_t winc = w;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, winc, o);
            if (lane >= o) winc += n;
        }
        warp_sums[lane] = winc - w;
        if (lane == 31) *total = winc;
    }
    __syncthreads();
    return inc - v + warp_sums[wid];
}

static __global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const uint32_t *in, uint32_t n,
                                                                   uint32_t *block_sums)
{
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t total;
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) s += (base + k < n) ? in[base + k] : 0u;
    block_exclusive_scan(s, &total, warp_sums);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: exclusive scan of up to 1024*SCAN_ITEMS block sums in place
static __global__ void __launch_bounds__(1024) scan_sums_kernel(uint32_t *block_sums, uint32_t nblocks,
                                                         uint32_t *grand_total)
{
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t total;
    const uint32_t base = threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = (base + k < nblocks) ? block_sums[base + k] : 0u; s += v[k]; }
    uint32_t off = block_exclusive_scan(s, &total, warp_sums);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < nblocks) block_sums[base + k] = off; off += v[k]; }
    if (threadIdx.x == 0) *grand_total = total;
}

// writes exclusive offsets in place; element n receives the grand total
static __global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(uint32_t *data, uint32_t n,
                                                                  const uint32_t *block_sums,
                                                                  const uint32_t *grand_total)
{
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t total;
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = (base + k < n) ? data[base + k] : 0u; s += v[k]; }
    uint32_t off = block_exclusive_scan(s, &total, warp_sums) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < n) data[base + k] = off; off += v[k]; }
    if (blockIdx.x == 0 && threadIdx.x == 0) data[n] = *grand_total;
}

// ---- selection ----------------------------------------------------------------------------

struct SelectParams {
    uint32_t U;
    const uint32_t *ukey;
    const uint32_t *ucount;
    const uint32_t *ufirst;
    uint32_t *parent_full;
    uint32_t *parent_one;
    uint32_t *best;          // per root: uid of the best member so far
    uint32_t *root;          // per unique: root used by the dissection
    uint8_t *dominated;
    uint8_t *dead;
    uint8_t *deadroot;
    uint8_t *selected;
    uint8_t *state;          // adjacency: 0 undecided, 1 selected, 2 removed
    uint32_t *stamp;
    const uint2 *edges;
    unsigned long long n_edges;
    uint32_t round;
    uint32_t *bitmap;        // bits for records [bitmap_base, bitmap_base + bitmap_n) only
    uint32_t bitmap_base, bitmap_n;
    int own_only;
    uint32_t *minfirst;
    int method;
    DevCounters *ctr;
    uint8_t rank_of_code[256];
};

// Per root keep the member with the largest (count, key): the head of the reference's
// descending sort (__init__.py:99-101) resp. the survivor of a singleton chain.
template <int K, int PW>
static __global__ void __launch_bounds__(256) root_best_kernel(const __grid_constant__ SelectParams P, int only_singletons)
{
    const uint32_t u = blockIdx.x * 256u + threadIdx.x;
    if (u >= P.U) return;
    const uint32_t cu = P.ucount[u];
    uint32_t *parent = only_singletons ? P.parent_one : P.parent_full;
    if (only_singletons && cu != 1) { P.root[u] = u; return; }
    const uint32_t r = uf_find(parent, u);
    P.root[u] = r;
    if (only_singletons && P.dead[u]) P.deadroot[r] = 1;
    Key<K, PW> ku;
    bool loaded = false;
    uint32_t cur = ld_relaxed_u32(P.best + r);
    while (cur != u) {
        if (!loaded) { load_key<K, PW>(P.ukey, u, ku); loaded = true; }
        Key<K, PW> kc;
        load_key<K, PW>(P.ukey, cur, kc);
        const uint32_t cc = P.ucount[cur];
        if (!prio_less<K, PW>(cc, kc, cu, ku, P.rank_of_code)) break;   // cur >= u
        const uint32_t old = atomicCAS(P.best + r, cur, u);
        if (old == cur) break;
        cur = old;
    }
}

static __global__ void __launch_bounds__(256) select_kernel(const __grid_constant__ SelectParams P)
{
    const uint32_t u = blockIdx.x * 256u + threadIdx.x;
    bool sel = false;
    if (u < P.U) {
        // independent loads first (the kernel is latency-bound), dependent ones after
        const uint32_t f = __ldcs(P.ufirst + u) - P.bitmap_base;   // wraps for records of other shards
        const uint32_t c = __ldcs(P.ucount + u);
        const uint32_t r = P.method != METHOD_ADJACENCY ? __ldcs(P.root + u) : 0u;
        const uint8_t dom = P.method == METHOD_DIRECTIONAL ? P.dominated[u] : (uint8_t)0;
        // sharded jobs: a rank only decides the keys whose first record is its own
        if (!P.own_only || f < P.bitmap_n) {
            if (P.method == METHOD_DIRECTIONAL) {
                if (c >= 2) sel = !dom;
                else sel = !P.deadroot[r] && P.best[r] == u;
            } else if (P.method == METHOD_HIGHEST) {
                sel = P.best[r] == u;
            } else {
                sel = P.state[u] == 1;
            }
        }
        P.selected[u] = sel ? 1 : 0;
        if (sel && P.bitmap && f < P.bitmap_n) atomicOr(P.bitmap + (f >> 5), 1u << (f & 31));
    }
    block_add(sel ? 1u : 0u, &P.ctr->n_selected);   // one atomic per block: a per-warp atomic on one address serialises the kernel
}

// adjacency (__init__.py:105-122) == greedy maximal independent set in descending
// (count, key) order.  Round r: an undecided key whose undecided/selected higher
// neighbours are all gone becomes selected; a key with a selected higher neighbour is
// removed.  edges hold (higher, lower).
static __global__ void __launch_bounds__(256) adj_edge_kernel(const __grid_constant__ SelectParams P)
{
    const unsigned long long e = (unsigned long long)blockIdx.x * 256u + threadIdx.x;
    if (e >= P.n_edges) return;
    const uint2 ed = P.edges[e];
    const uint8_t sh = P.state[ed.x], sl = P.state[ed.y];
    if (sl != 0) return;
    if (sh == 1) P.state[ed.y] = 2;
    else if (sh == 0) P.stamp[ed.y] = P.round;
}
static __global__ void __launch_bounds__(256) adj_node_kernel(const __grid_constant__ SelectParams P)
{
    const uint32_t u = blockIdx.x * 256u + threadIdx.x;
    bool und = false;
    if (u < P.U && P.state[u] == 0) {
        if (P.stamp[u] != P.round) P.state[u] = 1;
        else und = true;
    }
    const uint32_t b = __ballot_sync(0xFFFFFFFFu, und);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(&P.ctr->undecided, (uint32_t)__popc(b));
}

// canonical cluster label for the result view: smallest `first` among the members
static __global__ void __launch_bounds__(256) label_min_kernel(const __grid_constant__ SelectParams P)
{
    const uint32_t u = blockIdx.x * 256u + threadIdx.x;
    if (u >= P.U) return;
    const uint32_t r = uf_find(P.parent_full, u);
    P.root[u] = r;
    atomicMin(P.minfirst + r, P.ufirst[u]);
}

// tile-sharded job: root of every own unique in the job-wide slot space (slot of unique u = id_add + u * id_mul) (the per-unique view of such a job labels a
// cluster by its root; the caller that holds all ranks' views turns roots into smallest-first labels)
static __global__ void __launch_bounds__(256) root_gid_kernel(uint32_t U, uint32_t id_mul, uint32_t id_add, uint32_t *parent,
                                                              uint32_t *root)
{
    const uint32_t u = blockIdx.x * 256u + threadIdx.x;
    if (u < U) root[u] = uf_find(parent, id_add + u * id_mul);
}

// ---- sharding across GPUs (DESIGN.md "Multi-GPU") ----------------------------------------------
//
// Records are split contiguously over the ranks.  Each rank dedupes its shard, sends every
// local unique {key, count, first} to the key's owner rank (hash % world), the owners merge
// them (sum of counts, min of first), the merged tables are all-gathered so every rank holds
// the whole unique set, and each rank then runs the pigeonhole passes for the buckets it
// owns.  Spanning-forest pairs and flags are exchanged at the end.

__device__ __forceinline__ uint32_t key_owner(uint64_t h, uint32_t world)
{
    return (uint32_t)((h >> 17) % world);   // bits disjoint from the slot index (mulhi of h)
}

// counts per owner (block-local histogram first; world <= 64)
template <int K, int PW>
static __global__ void __launch_bounds__(256) owner_count_kernel(uint32_t U, const uint32_t *__restrict__ ukey,
                                                                 uint32_t world, uint32_t *owner_cnt)
{
    __shared__ uint32_t h[64];
    if (threadIdx.x < 64) h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t u = blockIdx.x * 256u + threadIdx.x;
    if (u < U) {
        Key<K, PW> key;
        load_key<K, PW>(ukey, u, key);
        atomicAdd(&h[key_owner(hash_key(key), world)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < world && h[threadIdx.x]) atomicAdd(owner_cnt + threadIdx.x, h[threadIdx.x]);
}

// send buffer grouped by owner: records of slot_words(KW) words {key, count, first}
template <int K, int PW>
static __global__ void __launch_bounds__(256) owner_scatter_kernel(uint32_t U, const uint32_t *__restrict__ ukey,
                                                                   const uint32_t *__restrict__ ucount,
                                                                   const uint32_t *__restrict__ ufirst,
                                                                   uint32_t world, uint32_t *cursor /* starts at the owner offsets */,
                                                                   uint32_t *send)
{
    constexpr int KW = K * PW, RW = slot_words(KW);
    const uint32_t u = blockIdx.x * 256u + threadIdx.x;
    if (u >= U) return;
    Key<K, PW> key;
    load_key<K, PW>(ukey, u, key);
    const uint32_t o = key_owner(hash_key(key), world);
    // one atomic per (warp, owner)
    const uint32_t peers = __match_any_sync(__activemask(), o);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(cursor + o, (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    const uint32_t pos = base + __popc(peers & ((1u << lane) - 1u));
    uint32_t e[RW];
#pragma unroll
    for (int i = 0; i < RW; i++) e[i] = 0;
#pragma unroll
    for (int i = 0; i < KW; i++) e[i] = key.w[i];
    e[KW] = ucount[u];
    e[KW + 1] = ufirst[u];
    uint4 *dst = reinterpret_cast<uint4 *>(send + (size_t)pos * RW);
#pragma unroll
    for (int c = 0; c < RW / 4; c++) dst[c] = make_uint4(e[4 * c], e[4 * c + 1], e[4 * c + 2], e[4 * c + 3]);
}

// owner side: merge the received records (sum of counts, min of first)
template <int K, int PW>
static __global__ void __launch_bounds__(256) merge_insert_kernel(uint32_t n, const uint32_t *__restrict__ recs,
                                                                  const __grid_constant__ TableRef tab)
{
    constexpr int KW = K * PW, RW = slot_words(KW);
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    uint32_t claimed = NO_CLAIM;
    if (i < n) {
        const uint4 *src = reinterpret_cast<const uint4 *>(recs + (size_t)i * RW);
        uint32_t e[RW];
#pragma unroll
        for (int c = 0; c < RW / 4; c++) {
            const uint4 v = __ldg(src + c);
            e[4 * c] = v.x; e[4 * c + 1] = v.y; e[4 * c + 2] = v.z; e[4 * c + 3] = v.w;
        }
        Key<K, PW> key;
#pragma unroll
        for (int j = 0; j < KW; j++) key.w[j] = e[j];
        claimed = table_insert<K, PW>(tab, key, hash_key(key), e[KW + 1], e[KW]);
    }
    const uint32_t pos = block_reserve(claimed != NO_CLAIM, &tab.ctr->n_unique);
    if (claimed != NO_CLAIM) tab.uslot[pos] = claimed;
}

// gather that drops keys whose every record was filtered out (count 0)
template <int K, int PW>
static __global__ void __launch_bounds__(256) gather_nonzero_kernel(uint32_t U, const uint32_t *__restrict__ table,
                                                                    const uint32_t *__restrict__ uslot,
                                                                    uint32_t *__restrict__ ukey,
                                                                    uint32_t *__restrict__ ucount,
                                                                    uint32_t *__restrict__ ufirst,
                                                                    uint32_t *kept, int keep_zero = 0,
                                                                    uint32_t cap = 0xFFFFFFFFu, uint32_t *overflow = nullptr,
                                                                    const uint32_t *U_dev = nullptr)
{
    constexpr int KW = K * PW, RW = slot_words(KW);
    const uint32_t u = blockIdx.x * 256u + threadIdx.x;
    if (u >= U || (U_dev && u >= *U_dev)) return;
    const uint4 *rec = reinterpret_cast<const uint4 *>(table + (size_t)uslot[u] * RW);
    uint32_t w[RW];
#pragma unroll
    for (int c = 0; c < RW / 4; c++) {
        const uint4 v = __ldcs(rec + c);
        w[4 * c] = v.x; w[4 * c + 1] = v.y; w[4 * c + 2] = v.z; w[4 * c + 3] = v.w;
    }
    if (w[KW] == 0 && !keep_zero) return;
    const uint32_t pos = aggregated_inc(kept);
    if (pos >= cap) { *overflow = 1u; return; }
#pragma unroll
    for (int i = 0; i < KW; i++) ukey[(size_t)pos * KW + i] = w[i];
    ucount[pos] = w[KW];
    ufirst[pos] = w[KW + 1];
}

// The parent links (u, parent[u]) of one rank's forest: the same connectivity as its edges, read
// with one streaming pass (no find), to be replayed on every other rank.
static __global__ void __launch_bounds__(256) forest_links_kernel(uint32_t U, const uint32_t *__restrict__ parent,
                                                                  uint2 *pairs, uint32_t *n_pairs)
{
    const uint32_t u = blockIdx.x * 256u + threadIdx.x;
    const uint32_t p = u < U ? __ldcs(parent + u) : u;
    if (p != u) pairs[aggregated_inc(n_pairs)] = make_uint2(u, p);
}

static __global__ void __launch_bounds__(256) apply_pairs_kernel(uint32_t n, const uint2 *__restrict__ pairs,
                                                                 uint32_t *parent)
{
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i >= n) return;
    const uint2 p = pairs[i];
    uf_union(parent, p.x, p.y);
}

static __global__ void __launch_bounds__(256) max_u8_kernel(size_t n4, uint32_t *__restrict__ dst,
                                                            const uint32_t *__restrict__ src)
{
    const size_t i = (size_t)blockIdx.x * 256u + threadIdx.x;
    if (i < n4) dst[i] = __vmaxu4(dst[i], src[i]);
}

static __global__ void __launch_bounds__(256) count_roots_kernel(uint32_t U, const uint32_t *__restrict__ parent,
                                                                 uint32_t *n_roots)
{
    const uint32_t u = blockIdx.x * 256u + threadIdx.x;
    const bool root = u < U && parent[u] == u;
    const uint32_t b = __ballot_sync(0xFFFFFFFFu, root);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(n_roots, (uint32_t)__popc(b));
}

// ---- function-level kernels (batched _fastq / _distance entry points) -------------------

static __global__ void __launch_bounds__(128) error_rate_kernel(const uint8_t *phred, const uint64_t *off,
                                                         uint64_t n, uint32_t phred_offset,
                                                         double *out, DevCounters *ctr)
{
    __shared__ double lut_d[128];
    if (threadIdx.x < 128) lut_d[threadIdx.x] = __longlong_as_double((long long)FQD_PHRED_LUT_BITS[threadIdx.x]);
    __syncthreads();
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint8_t *q = phred + off[t];
    const uint64_t len = off[t + 1] - off[t];
    const uint32_t max_score = (126u - phred_offset) & 0xFFu;
    double total = 0.0;
    for (uint64_t i = 0; i < len; i++) {
        const uint32_t c = q[i];
        const uint32_t score = (c - phred_offset) & 0xFFu;
        if (score > max_score) {
            atomicMin(&ctr->phred_err, (unsigned long long)((t << 8) | c));
            return;
        }
        total = __dadd_rn(total, lut_d[score]);
    }
    out[t] = __ddiv_rn(total, (double)len);
}

// Byte-level predicates for arbitrary (latin-1) strings of any length: the reference's
// within_distance accepts any 1-byte-kind str (_distancemodule.c:64-73).  Hamming: byte
// compare with early exit (distances.h:8-31).  Levenshtein: banded DP, band |i-j| <= d,
// three rolling rows in local memory are avoided by keeping one row of 2d+1 cells.
constexpr int MAX_BAND_D = 31;
static __global__ void __launch_bounds__(128) within_distance_kernel(const uint8_t *a, const uint64_t *aoff,
                                                              const uint8_t *b, const uint64_t *boff,
                                                              uint64_t n, int d, int edit, uint8_t *out)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint8_t *x = a + aoff[t], *y = b + boff[t];
    const long long lx = (long long)(aoff[t + 1] - aoff[t]), ly = (long long)(boff[t + 1] - boff[t]);
    if (!edit) {
        if (lx != ly) { out[t] = 0; return; }
        int budget = d;
        for (long long i = 0; i < lx; i++) {
            if (x[i] != y[i]) { if (--budget < 0) { out[t] = 0; return; } }
        }
        out[t] = 1;
        return;
    }
    const long long diff = lx > ly ? lx - ly : ly - lx;
    if (d < 0 || diff > d) { out[t] = 0; return; }
    // row[k] holds D[i][i - d + k] for k in 0..2d
    const int INF = 1 << 28;
    int row[2 * MAX_BAND_D + 1];
    const int bw = 2 * d + 1;
    for (int k = 0; k < bw; k++) { const long long j = (long long)k - d; row[k] = (j >= 0 && j <= ly && j <= d) ? (int)j : INF; }
    for (long long i = 1; i <= lx; i++) {
        int prev_left = INF;   // D[i][j-1] of the new row
        for (int k = 0; k < bw; k++) {
            const long long j = i - d + k;
            int v = INF;
            if (j >= 0 && j <= ly) {
                if (j == 0) v = i <= d ? (int)i : INF;
                else {
                    const int sub = row[k] + (x[i - 1] != y[j - 1]);       // D[i-1][j-1]
                    const int del = (k + 1 < bw) ? row[k + 1] + 1 : INF;   // D[i-1][j]
                    const int ins = prev_left + 1;                          // D[i][j-1]
                    v = min(sub, min(del, ins));
                }
            }
            row[k] = v;
            prev_left = v;
        }
    }
    const long long kk = ly - lx + d;   // column of D[lx][ly] in the last row
    out[t] = (kk >= 0 && kk < bw && row[kk] <= d) ? 1 : 0;
}

}  // namespace fqd

#include "partitioned.cuh"
