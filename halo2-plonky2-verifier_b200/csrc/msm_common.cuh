// Small shared pieces of the MSM translation units (msm.cu: digits, sort, accumulate, combine; msm_reduce.cu: bucket reduction).
#pragma once
#include "collectives.cuh"

namespace b200zk {

DEV G1X g1x_load(const G1X* p) {
    G1X r;
    r.x = f_load(&p->x);
    r.y = f_load(&p->y);
    r.zz = f_load(&p->zz);
    r.zzz = f_load(&p->zzz);
    return r;
}
DEV void g1x_store(G1X* p, const G1X& v) {
    f_store(&p->x, v.x);
    f_store(&p->y, v.y);
    f_store(&p->zz, v.zz);
    f_store(&p->zzz, v.zzz);
}
DEV G1Affine g1a_load_ro(const G1Affine* p) {
    G1Affine r;
    r.x = f_load_ro(&p->x);
    r.y = f_load_ro(&p->y);
    return r;
}

constexpr uint32_t INVALID_KEY = 0xffffffffu;
constexpr int ACC_THREADS = 128;
// one level of the head combine (msm_reduce.cu): segmented sum of a key-sorted list of partial sums, 32 per warp
void msm_launch_combine(const G1X* pts, const uint32_t* keys, uint32_t n, G1X* bucket_sums, G1X* heads_out, uint32_t* keys_out, cudaStream_t s);
// Phase B of a commit batch: F(set) = Σ_b (b+1)·bucket[b] for G bucket sets of B buckets (msm_reduce.cu)
void msm_reduce_groups(Context& ctx, const G1X* bucket_sums, uint32_t G, uint32_t B, std::vector<G1X>& sums_host, int gather_ranks = 1);

}  // namespace b200zk
