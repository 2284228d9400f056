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
                             int, const Codec &, uint32_t[8]);
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

int run_sharded(fqd_context **ctxs, const DeviceJob *jobs, const uint32_t *index_base,
                fqd_cluster_stats **stats, int n_local, Exchange *ex, int world, const Codec &codec,
                uint32_t max_len, uint32_t unknown_out[8])
{
    const int bits = codec.bits;
    const int best_pw = pick_pw(bits, max_len);
    if (!best_pw) return FQD_ERR_UNSUPPORTED;
    int rc = RC_NOT_IN_GROUP;
#define X(n) \
    if (rc == RC_NOT_IN_GROUP) rc = run_sharded_group##n(bits, best_pw, ctxs, jobs, index_base, stats, n_local, ex, world, codec, unknown_out);
    X(0) X(1) X(2) X(3) X(4) X(5) X(6)
#undef X
    return rc == RC_NOT_IN_GROUP ? FQD_ERR_UNSUPPORTED : rc;
}

}  // namespace fqd
