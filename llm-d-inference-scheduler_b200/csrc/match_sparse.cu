// match_sparse.cu -- a2-a14 over the hash rows of a batch: the match / score / pick kernel of the throughput path (and of
// the endpoint-sharded mode): one warp per request, the per-request work is match_sparse.cuh.
#include <cstdlib>

#include "match_sparse.cuh"

namespace epp {

namespace {
constexpr int kWarps = 8;
}  // namespace

// MINCTA: resident CTAs per SM the register allocation is bounded for (8 -> 32 registers: 64 resident warps hide the
// probe latency; 6 -> 40 registers).  kTie / kStages: see match_sparse.cuh.
template <int MINCTA, bool kSharded, bool kTie, int kStages>
__global__ void __launch_bounds__(kWarps * 32, MINCTA) k_match_pick_sparse(PickParams p) {
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t gwarp = (int64_t)blockIdx.x * kWarps + warp;
    const int64_t nwarps = (int64_t)gridDim.x * kWarps;
    const bool counting = p.work_counters != nullptr;
    sparse::Work wk;
    for (int64_t r = gwarp; r < p.R; r += nwarps) sparse::match_request<false, kSharded, kTie, kStages>(p, r, lane, counting, wk);
    if (counting) {
        for (int o = 16; o; o >>= 1) wk.postings += __shfl_xor_sync(0xffffffffu, wk.postings, o);
        if (lane == 0) {
            atomicAdd(&p.work_counters[0], wk.probes);
            atomicAdd(&p.work_counters[1], wk.postings);
        }
    }
}

namespace {
template <int MINCTA, bool kSharded, bool kTie, int kStages>
cudaError_t launch_one(const PickParams &p, int sm_count, cudaStream_t s) {
    int dev = 0;
    cudaGetDevice(&dev);
    static int occ_dev[64] = {};                       // per device: engines of one process may sit on different GPUs
    int &occ = occ_dev[dev & 63];
    if (!occ) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_match_pick_sparse<MINCTA, kSharded, kTie, kStages>, kWarps * 32, 0);
        if (occ < 1) occ = 1;
    }
    if (sm_count <= 0) sm_count = 148;
    const int64_t need = (p.R + kWarps - 1) / kWarps;
    const int grid = (int)(need < (int64_t)sm_count * occ ? need : (int64_t)sm_count * occ);
    k_match_pick_sparse<MINCTA, kSharded, kTie, kStages><<<grid, kWarps * 32, 0, s>>>(p);
    return cudaGetLastError();
}
}  // namespace

cudaError_t launch_match_pick_sparse(const PickParams &p, int sm_count, cudaStream_t s, int *launches) {
    if (p.R <= 0) return cudaSuccess;
    static int ctas = 0;                               // EPP_MATCH_CTAS=6: the 40-register variant (A/B; 8 measured faster)
    if (!ctas) { const char *v = getenv("EPP_MATCH_CTAS"); ctas = (v && atoi(v) == 6) ? 6 : 8; }
    const bool tie = p.tie_seed != 0;
    const int stages = p.encode_on ? 3 : (p.n_profiles >= 2 ? 2 : 1);
    cudaError_t e;
    // the hot configuration (one profile, lowest-slot ties) gets the 32-register build; the variants that carry more
    // code (stages, tie rule, sharding) keep 40 registers
    if (p.global_masks) e = launch_one<6, true, false, 1>(p, sm_count, s);
    else if (stages == 1 && !tie) e = ctas == 8 ? launch_one<8, false, false, 1>(p, sm_count, s) : launch_one<6, false, false, 1>(p, sm_count, s);
    else if (stages == 1) e = launch_one<6, false, true, 1>(p, sm_count, s);
    else if (stages == 2) e = tie ? launch_one<6, false, true, 2>(p, sm_count, s) : launch_one<6, false, false, 2>(p, sm_count, s);
    else e = tie ? launch_one<6, false, true, 3>(p, sm_count, s) : launch_one<6, false, false, 3>(p, sm_count, s);
    if (launches) *launches += 1;
    return e;
}

}  // namespace epp
