// Bucket reduction of the Pippenger MSM (msm.cu has the front half: digits, counting sort, accumulation, head combine).
// Its own translation unit because its kernels want the OUT-OF-LINE multiplier bodies (B200ZK_NOINLINE_MUL): they chain full
// XYZZ additions (≈ 14 products each) with two accumulators live, the fully inlined code is large and register-bound, and
// calling the product / square as functions shortens it (measured on S20-bn: msm stage 93.2 -> 92.0 ms with every MSM kernel
// out of line although the accumulate kernel itself got 1.8 ms slower; profiles/ncu_summary_r02.md) — so the accumulate
// kernel keeps the inlined multiplier and this file does not.
#define B200ZK_NOINLINE_MUL 1
#include <algorithm>
#include <chrono>

#include "msm_common.cuh"

namespace b200zk {

// Segmented sum of a key-sorted list of partial sums (INVALID_KEY entries are holes), one entry per lane: a 5-step shuffle
// scan leaves the total of every run of equal keys in the run's first lane. A run that starts inside the warp is added into
// its bucket (one writer per bucket and launch); the part of a run that began in an earlier warp goes to the next level,
// which is 32x shorter. Latency per level: at most 5 dependent additions.
DEV G1X g1x_shfl_down(const G1X& v, uint32_t d) {
    G1X r;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        r.x.l[i] = __shfl_down_sync(0xffffffffu, v.x.l[i], d);
        r.y.l[i] = __shfl_down_sync(0xffffffffu, v.y.l[i], d);
        r.zz.l[i] = __shfl_down_sync(0xffffffffu, v.zz.l[i], d);
        r.zzz.l[i] = __shfl_down_sync(0xffffffffu, v.zzz.l[i], d);
    }
    return r;
}
__global__ void __launch_bounds__(ACC_THREADS) msm_combine_kernel(const G1X* pts, const uint32_t* keys, uint32_t n, G1X* bucket_sums, G1X* heads_out,
                                                                  uint32_t* keys_out) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31, warp = p >> 5;
    if ((p & ~31u) >= n) return;  // whole warps only
    const uint32_t key = p < n ? __ldg(keys + p) : INVALID_KEY;
    const bool live = key != INVALID_KEY;
    G1X acc = live ? g1x_load(pts + p) : g1x_identity();
    for (uint32_t d = 1; d < 32; d <<= 1) {
        const uint32_t okey = __shfl_down_sync(0xffffffffu, key, d);
        const G1X o = g1x_shfl_down(acc, d);
        if (live && lane + d < 32 && okey == key) acc = g1x_add(acc, o);
    }
    uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
    if (lane == 0) prev = p > 0 ? __ldg(keys + p - 1) : INVALID_KEY;
    const bool run_start = live && (lane == 0 || prev != key);
    const bool started_before = live && lane == 0 && prev == key;
    if (started_before) g1x_store(heads_out + warp, acc);
    else if (run_start) g1x_store(bucket_sums + key, g1x_add(g1x_load(bucket_sums + key), acc));
    if (lane == 0) keys_out[warp] = started_before ? key : INVALID_KEY;
}

void msm_launch_combine(const G1X* pts, const uint32_t* keys, uint32_t n, G1X* bucket_sums, G1X* heads_out, uint32_t* keys_out, cudaStream_t s) {
    msm_combine_kernel<<<(n + ACC_THREADS - 1) / ACC_THREADS, ACC_THREADS, 0, s>>>(pts, keys, n, bucket_sums, heads_out, keys_out);
    ++g_launch_count;
    CUDA_CHECK(cudaGetLastError());
}

// ---- bucket reduction: per window F(B) = sum_b (b+1)·B[b] ---------------------------------------------------
// Chunks of m consecutive entries give tot_q = Σ_r (r+1)·X[qm+r] and run_q = Σ_r X[qm+r] with 2 additions per
// entry; then F(X) = Σ_q tot_q + m·(F(run) − S), S = ΣX, so the same kernel recurses on the `run` list (÷m per
// level) and the per-level sums T_l = Σ_q tot_q are combined by a short Horner in m: A = S; A = T_l + m·(A − S).
#ifndef B200ZK_RED_LOG_M
#define B200ZK_RED_LOG_M 3
#endif
constexpr int RED_LOG_M = B200ZK_RED_LOG_M;
// lists are window-major: X[w·len + i]; outputs tot[w·(len/m) + q], run[w·(len/m) + q]
#ifndef B200ZK_REDUCE_MIN_CTAS
#define B200ZK_REDUCE_MIN_CTAS 1
#endif
__global__ void __launch_bounds__(128, B200ZK_REDUCE_MIN_CTAS) msm_reduce_chunks_kernel(const G1X* X, uint32_t total_chunks, uint32_t m, G1X* tot_out, G1X* run_out) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= total_chunks) return;
    const G1X* b = X + (size_t)j * m;
    G1X run = g1x_identity(), tot = g1x_identity();
    for (int i = (int)m - 1; i >= 0; --i) {
        run = g1x_add(run, g1x_load(b + i));
        tot = g1x_add(tot, run);
    }
    g1x_store(tot_out + j, tot);
    g1x_store(run_out + j, run);
}
// Σ over lists: block (w, level, slice) sums entries [slice·SUM_SLICE, (slice+1)·SUM_SLICE) of tot_level[w·len ..]
// into out[(level·W + w)·slices + slice]. Called twice: lists -> per-slice partials -> T[level·W + w].
constexpr uint32_t SUM_SLICE = 1024;
struct SumLevels {
    const G1X* tot[16];
    uint32_t len[16];
};
__global__ void __launch_bounds__(128) msm_reduce_sum_kernel(SumLevels lv, uint32_t W, uint32_t slices, G1X* out) {
    __shared__ G1X sh[128];
    const uint32_t w = blockIdx.x, level = blockIdx.y, slice = blockIdx.z;
    const uint32_t len = lv.len[level];
    const G1X* src = lv.tot[level] + (size_t)w * len;
    const uint32_t lo = slice * SUM_SLICE, hi = lo + SUM_SLICE < len ? lo + SUM_SLICE : len;
    G1X acc = g1x_identity();
    for (uint32_t q = lo + threadIdx.x; q < hi; q += blockDim.x) acc = g1x_add(acc, g1x_load(src + q));
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (uint32_t s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] = g1x_add(sh[threadIdx.x], sh[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) g1x_store(out + ((size_t)level * W + w) * slices + slice, sh[0]);
}
// Tail of the recursion: F = Σ_b (b+1)·X[b] and S = Σ_b X[b] of one list of len <= TAIL_MAX entries (a power of two) per
// block. Thread t takes E = len/256 consecutive entries (run_t, tot_t as above); with suff_t = Σ_{t'>=t} run_t' from a
// block-wide suffix scan, F = Σ_t tot_t + E·Σ_{t>=1} suff_t. About 26 dependent additions instead of four more launches of
// 16 each: the deep levels of the reduction are pure latency.
constexpr uint32_t TAIL_MAX = 1024, TAIL_THREADS = 256;
__global__ void __launch_bounds__(TAIL_THREADS) msm_reduce_tail_kernel(const G1X* X, uint32_t len, G1X* F_out, G1X* S_out) {
    extern __shared__ uint4 tail_smem[];
    G1X* cur = reinterpret_cast<G1X*>(tail_smem);
    G1X* nxt = cur + TAIL_THREADS;
    const uint32_t g = blockIdx.x, t = threadIdx.x;
    const uint32_t E = len > TAIL_THREADS ? len / TAIL_THREADS : 1, active = len / E;
    G1X run = g1x_identity(), tot = g1x_identity();
    if (t < active) {
        const G1X* b = X + (size_t)g * len + (size_t)t * E;
        for (int i = (int)E - 1; i >= 0; --i) {
            run = g1x_add(run, g1x_load(b + i));
            tot = g1x_add(tot, run);
        }
    }
    cur[t] = run;
    __syncthreads();
    for (uint32_t d = 1; d < TAIL_THREADS; d <<= 1) {  // inclusive suffix scan (Hillis–Steele)
        G1X v = cur[t];
        if (t + d < TAIL_THREADS) v = g1x_add(v, cur[t + d]);
        nxt[t] = v;
        G1X* tmp = cur;
        cur = nxt;
        nxt = tmp;
        __syncthreads();
    }
    G1X v = tot;
    if (t >= 1) {
        G1X e = cur[t];
        for (uint32_t k = 1; k < E; k <<= 1) e = g1x_dbl(e);
        v = g1x_add(v, e);
    }
    if (t == 0) g1x_store(S_out + g, cur[0]);
    nxt[t] = v;
    __syncthreads();
    for (uint32_t h = TAIL_THREADS / 2; h > 0; h >>= 1) {
        if (t < h) nxt[t] = g1x_add(nxt[t], nxt[t + h]);
        __syncthreads();
    }
    if (t == 0) g1x_store(F_out + g, nxt[0]);
}
// per-device kernel attributes; called by b200zk_create for its device
void msm_init_device() {
    CUDA_CHECK(cudaFuncSetAttribute(msm_reduce_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * TAIL_THREADS * sizeof(G1X))));
}
// one thread per window: Horner over the levels, starting from the tail's F and S
struct HornerLevels {
    uint32_t log_m[16];
    uint32_t levels;
};
__global__ void msm_reduce_horner_kernel(const G1X* T, const G1X* F_tail, const G1X* S, HornerLevels hl, uint32_t W, G1X* window_sums) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    const G1X neg_s = g1x_neg(g1x_load(S + w));
    G1X a = g1x_load(F_tail + w);
    for (int l = (int)hl.levels - 1; l >= 0; --l) {
        a = g1x_add(a, neg_s);
        for (uint32_t d = 0; d < hl.log_m[l]; ++d) a = g1x_dbl(a);
        a = g1x_add(a, g1x_load(T + (size_t)l * W + w));
    }
    g1x_store(window_sums + w, a);
}

// Phase B for `G` bucket sets of B buckets each (window-major): F(set) = Σ_b (b+1)·bucket[b] -> sums_host[G] (XYZZ).
// One reduction serves every column of a batch: the deep levels are latency bound (≈0.12 ms each whatever G is).
// With `gather_ranks` > 1 the G sums of every rank are all-gathered on the device (NCCL, straight out of the reduction's
// output buffer) before the one read-back: sums_host then holds gather_ranks·G entries, rank-major.
void msm_reduce_groups(Context& ctx, const G1X* bucket_sums, uint32_t G, uint32_t B, std::vector<G1X>& sums_host, int gather_ranks) {
    cudaStream_t s = ctx.stream;
    sums_host.assign((size_t)G * gather_ranks, g1x_identity());
    if (G == 0) return;
    DevBuf<G1X> wsums(G, s);
    {
        std::vector<DevBuf<G1X>> tots, runs;
        SumLevels sl{};
        HornerLevels hl{};
        const G1X* X = bucket_sums;
        uint32_t len = B, level = 0;
        while (len > TAIL_MAX) {
            const uint32_t lm = RED_LOG_M, m = 1u << lm, out_len = len >> lm, total_chunks = G * out_len;
            tots.emplace_back((size_t)total_chunks, s);
            runs.emplace_back((size_t)total_chunks, s);
            msm_reduce_chunks_kernel<<<(total_chunks + 127) / 128, 128, 0, s>>>(X, total_chunks, m, tots.back().get(), runs.back().get());
            ++g_launch_count;
            sl.tot[level] = tots.back().get();
            sl.len[level] = out_len;
            hl.log_m[level] = lm;
            X = runs.back().get();
            len = out_len;
            ++level;
        }
        hl.levels = level;
        // the remaining list (<= TAIL_MAX entries per set) in one launch
        const size_t tail_smem = 2 * TAIL_THREADS * sizeof(G1X);
        // the level sums T_l only need the `tot` lists: they run on an auxiliary stream beside the tail launch
        cudaStream_t side = ctx.aux_streams[0] && level ? ctx.aux_streams[0] : s;
        DevBuf<G1X> tail_f(G, s), tail_s(G, s), T((size_t)std::max<uint32_t>(level, 1) * G, s);
        uint32_t max_len = 0;
        for (uint32_t l = 0; l < level; ++l) max_len = std::max(max_len, sl.len[l]);
        const uint32_t slices = (max_len + SUM_SLICE - 1) / SUM_SLICE;
        DevBuf<G1X> part(slices > 1 ? (size_t)level * G * slices : 0, s);  // slices beyond a short list sum to the identity
        if (side != s) {
            CUDA_CHECK(cudaEventRecord(ctx.msm_fork, s));
            CUDA_CHECK(cudaStreamWaitEvent(side, ctx.msm_fork, 0));
        }
        if (level && slices <= 1) {
            msm_reduce_sum_kernel<<<dim3(G, level, 1), 128, 0, side>>>(sl, G, 1, T.get());
            ++g_launch_count;
        } else if (level) {
            msm_reduce_sum_kernel<<<dim3(G, level, slices), 128, 0, side>>>(sl, G, slices, part.get());
            SumLevels sl2{};
            for (uint32_t l = 0; l < level; ++l) {
                sl2.tot[l] = part.get() + (size_t)l * G * slices;
                sl2.len[l] = slices;
            }
            msm_reduce_sum_kernel<<<dim3(G, level, 1), 128, 0, side>>>(sl2, G, 1, T.get());
            g_launch_count += 2;
        }
        msm_reduce_tail_kernel<<<G, TAIL_THREADS, tail_smem, s>>>(X, len, tail_f.get(), tail_s.get());
        ++g_launch_count;
        if (side != s) {
            CUDA_CHECK(cudaEventRecord(ctx.msm_join[0], side));
            CUDA_CHECK(cudaStreamWaitEvent(s, ctx.msm_join[0], 0));
        }
        if (level == 0) {
            CUDA_CHECK(cudaMemcpyAsync(wsums.get(), tail_f.get(), G * sizeof(G1X), cudaMemcpyDeviceToDevice, s));
        } else {
            msm_reduce_horner_kernel<<<(G + 31) / 32, 32, 0, s>>>(T.get(), tail_f.get(), tail_s.get(), hl, G, wsums.get());
            ++g_launch_count;
        }
        CUDA_CHECK(cudaGetLastError());
    }
    if (gather_ranks > 1) {
        const auto t0 = std::chrono::steady_clock::now();
        Sharder(ctx).dev_to_host_allgather(wsums.get(), G * sizeof(G1X), sums_host.data());
        ctx.exchange_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return;
    }
    CUDA_CHECK(cudaMemcpyAsync(sums_host.data(), wsums.get(), G * sizeof(G1X), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
}

}  // namespace b200zk
