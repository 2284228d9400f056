// pipeline.cu -- dispatch of a job to the (K, PW) instantiation that packs its keys (instances.h).
// The launch plans themselves are the host templates of pipeline_impl.cuh, compiled per instance group
// by pipeline_inst.cu.
#include <algorithm>
#include <vector>

#include "common.h"
#include "exchange.h"
#include "instances.h"

namespace fqd {

#define X(n) \
    int run_typed_group##n(int, int, fqd_context *, const DeviceJob &, const Codec &, fqd_cluster_stats *, uint32_t[8]); \
    int run_sharded_group##n(int, int, fqd_context **, const DeviceJob *, const uint32_t *, fqd_cluster_stats **, int, Exchange *, \
                             const ShardWorld &, const Codec &, uint32_t[8], bool);
X(0) X(1) X(2) X(3) X(4) X(5) X(6)
#undef X
static_assert(FQD_N_GROUPS == 7, "declare every instance group above");

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

static int pick_pw(int bits, uint32_t max_len)
{
    const uint32_t pw_needed = std::max(1u, (max_len + 31u) / 32u);
    int best_pw = 0;
#define X(K_, PW_) if (bits == K_ && (uint32_t)PW_ >= pw_needed && (best_pw == 0 || PW_ < best_pw)) best_pw = PW_;
    FQD_INSTANCES(X)
#undef X
    if (!best_pw)
        set_error("keys of %u symbols over a %d-bit alphabet exceed what this build packs "
                  "(max %u symbols)", max_len, bits, max_supported_length(bits));
    return best_pw;
}

int run_pipeline(fqd_context *ctx, const DeviceJob &job, const Codec &codec,
                 fqd_cluster_stats *stats, uint32_t unknown_out[8])
{
    const int bits = codec.bits;
    const int best_pw = pick_pw(bits, job.max_len);
    if (!best_pw) return FQD_ERR_UNSUPPORTED;
    int rc = RC_NOT_IN_GROUP;
#define X(n) if (rc == RC_NOT_IN_GROUP) rc = run_typed_group##n(bits, best_pw, ctx, job, codec, stats, unknown_out);
    X(0) X(1) X(2) X(3) X(4) X(5) X(6)
#undef X
    return rc == RC_NOT_IN_GROUP ? FQD_ERR_UNSUPPORTED : rc;
}

bool tile_plan_eligible(const DeviceJob &job, const Codec &codec, uint32_t max_len, int world)
{
    // 32-byte records: the packed key fits 6 words; Hamming only (the Levenshtein passes dereference unique ids)
    const uint32_t pw = std::max(1u, (max_len + 31u) / 32u);
    return !job.edit && world <= MAX_RANKS && slot_words((int)(codec.bits * pw)) == PART_RW;
}

// Arena bytes the tile-sharded plan may allocate past the inputs (an upper estimate: the plan checks that what
// other ranks read really lies inside the slab and otherwise hands the job to the replicated-set plan).
size_t tile_plan_bytes(uint64_t n_total, uint64_t n_max, int world, int d, int method)
{
    const uint64_t rec = 32, G = (uint64_t)std::max(world, 1);
    const uint64_t tiles_n = (uint64_t)tile_partitions(n_total) + G;             // regions of all tiles, on every rank
    const uint64_t tiles_h = (uint64_t)tile_partitions(std::max<uint64_t>(n_total / 2, 1u << 16)) + G;
    const uint64_t cap_u = n_total / G + n_total / (4 * G) + (1u << 16);         // uniques one owner can end up with
    uint64_t b = tiles_n * TILE_R * rec + tiles_n * 16;                          // records + fill counters
    b += (n_max / 4 + 4096) * rec;                                               // spill
    b += tiles_h * TILE_R * rec + tiles_h * 16;                                  // tiles of pass 1 written by the dedupe tiles
    if (d >= 2 || method == METHOD_ADJACENCY) b += 2 * (tiles_n * TILE_R * rec + tiles_n * 16);   // two pass buffers, worst case U = N
    b += cap_u * (24 + 4 + 4);                                                   // dense unique arrays
    b += (n_max + (1u << 16)) * 8;                                               // edge list
    if (method == METHOD_ADJACENCY) b += (2 * n_max + (1u << 16)) * 8 + (2 * n_total + (1u << 16)) * 8;
    b += cap_u * (rec + 4);                                                      // candidate list
    b += n_total / 8 + 4096;                                                     // keep bits over the records of the whole job
    const uint64_t ids = cap_u * G;                                              // job-wide id space
    b += ids * (4 + 4 + 4 + 4 + 5) + cap_u * 16;                                 // forests, best, flags, per-own-unique state
    b += (cap_u / 2) * rec * 3;                                                  // spill table of the oversize tiles (rarely large)
    return (size_t)(b + b / 16 + (64u << 20));
}

int run_sharded(fqd_context **ctxs, const DeviceJob *jobs, const uint32_t *index_base,
                fqd_cluster_stats **stats, int n_local, Exchange *ex, const ShardWorld &W, const Codec &codec,
                uint32_t max_len, uint32_t unknown_out[8], bool replicated_plan)
{
    if (!replicated_plan && !tile_plan_eligible(jobs[0], codec, max_len, W.world)) replicated_plan = true;
    const int bits = codec.bits;
    const int best_pw = pick_pw(bits, max_len);
    if (!best_pw) return FQD_ERR_UNSUPPORTED;
    int rc = RC_NOT_IN_GROUP;
#define X(n) \
    if (rc == RC_NOT_IN_GROUP) rc = run_sharded_group##n(bits, best_pw, ctxs, jobs, index_base, stats, n_local, ex, W, codec, unknown_out, replicated_plan);
    X(0) X(1) X(2) X(3) X(4) X(5) X(6)
#undef X
    return rc == RC_NOT_IN_GROUP ? FQD_ERR_UNSUPPORTED : rc;
}

}  // namespace fqd
