// pipeline_inst.cu -- one instance group of the launch plans (compiled with -DFQD_GROUP=n, see
// instances.h and build.py): exports run_typed_group<n> / run_sharded_group<n>, which pipeline.cu
// dispatches to by (bits, words per plane).
#include "pipeline_impl.cuh"
#include "instances.h"

#ifndef FQD_GROUP
#error "compile with -DFQD_GROUP=n"
#endif

#define FQD_CAT2(a, b) a##b
#define FQD_CAT(a, b) FQD_CAT2(a, b)
#define FQD_GROUP_INSTANCES FQD_CAT(FQD_INSTANCES_G, FQD_GROUP)

namespace fqd {

int FQD_CAT(run_typed_group, FQD_GROUP)(int bits, int pw, fqd_context *ctx, const DeviceJob &job, const Codec &codec,
                                        fqd_cluster_stats *stats, uint32_t unknown_out[8])
{
#define X(K_, PW_) if (bits == K_ && pw == PW_) return run_typed<K_, PW_>(ctx, job, codec, stats, unknown_out);
    FQD_GROUP_INSTANCES(X)
#undef X
    return RC_NOT_IN_GROUP;
}

int FQD_CAT(run_sharded_group, FQD_GROUP)(int bits, int pw, fqd_context **ctxs, const DeviceJob *jobs, const uint32_t *index_base,
                                          fqd_cluster_stats **stats, int n_local, Exchange *ex, const ShardWorld &W, const Codec &codec,
                                          uint32_t unknown_out[8], bool replicated_plan)
{
    std::vector<Shard> S(n_local);
    for (int i = 0; i < n_local; i++) {
        S[i].ctx = ctxs[i];
        S[i].job = jobs[i];
        S[i].index_base = index_base[i];
        S[i].st = stats[i];
    }
#define X(K_, PW_)                                                                                               \
    if (bits == K_ && pw == PW_)                                                                                 \
        return replicated_plan ? run_sharded_typed<K_, PW_>(S, ex, W.world, codec, unknown_out)                   \
                               : run_sharded_tiles<K_, PW_>(S, ex, W, codec, unknown_out);
    FQD_GROUP_INSTANCES(X)
#undef X
    return RC_NOT_IN_GROUP;
}

}  // namespace fqd
