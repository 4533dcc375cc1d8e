// Fr NTT / iNTT and the fused coset variants (replaces
// halo2_proofs::arithmetic::best_fft and EvaluationDomain::{lagrange_to_coeff, coeff_to_extended,
// extended_to_coeff}, reached from the reference through verifier/src/stark/mod.rs:543,593).
//
// Semantics = best_fft: natural order in, natural order out, out[j] = sum_i a[i] w^(ij).
// Structure: decimation-in-time over the bit-reversed input, log_n stages grouped into passes of <= 8 stages.
// One CTA owns a tile of 2^r points × C adjacent columns in shared memory and runs the r stages of its pass
// there; HBM is touched once per pass (read + write, 16-byte vector accesses in runs of C·32 bytes).
//   * pass 1 gathers through the bit reversal (reads runs of C elements, writes runs of 2^r elements), so no
//     separate permutation pass exists;
//   * later passes work in place on index = hi·2^s1 + m·2^s0 + lo with C consecutive `lo` per CTA;
//   * twiddles come from one table of w^i, i < N/2 (inverse transforms read it mirrored and negated), loaded
//     through the read-only path: early stages hit a handful of L1-resident entries, the last stage streams
//     N/2 entries once;
//   * zero padding + coset scaling (coeff_to_extended) are folded into the first pass's loads and the
//     divisor + coset un-scaling + truncation (extended_to_coeff, lagrange_to_coeff) into the last pass's stores.
// The kernel is integer-pipe bound (≈1 Montgomery product per butterfly ≈ 140 IMAD.WIDE per 128 B of shared
// memory traffic); see DESIGN.md for the roofline.
#include "common.cuh"

namespace b200zk {

std::atomic<unsigned long long> g_launch_count{0};
bool g_prof_enabled = false;
std::vector<ProfSpan> g_prof_spans;
std::mutex g_prof_mu;

static std::mutex g_arena_mu;
static std::map<cudaStream_t, Arena*> g_arenas;
Arena* arena_for(cudaStream_t s) {
    std::lock_guard<std::mutex> lk(g_arena_mu);
    auto it = g_arenas.find(s);
    return it == g_arenas.end() ? nullptr : it->second;
}
void arena_register(cudaStream_t s, Arena* a) {
    std::lock_guard<std::mutex> lk(g_arena_mu);
    if (a) g_arenas[s] = a;
    else g_arenas.erase(s);
}

struct PassParams {
    const Fr* in;
    Fr* out;
    const Fr* table;
    const Fr* pre3;
    const Fr* post3;
    unsigned long long in_len, out_len, batch_stride_in, batch_stride_out, in_stride, in_offset;
    uint32_t L, s0, r, logC;
    uint32_t table_shift, half_table;
    uint32_t inverse, first, last;
    uint32_t skip2;  // first pass of a 4x zero-padded input: stages 1-2 only replicate each non-zero element 4 times
};

#ifndef B200ZK_NTT_THREADS
#define B200ZK_NTT_THREADS 256
#endif
constexpr int NTT_THREADS = B200ZK_NTT_THREADS;
constexpr int NTT_MAX_R = 8;
constexpr int NTT_MAX_LOGC = 3;
constexpr int SMEM_PAD = 4;  // uint4 units between the low-half and high-half planes (bank offset 16)

DEV uint32_t swz(uint32_t m, uint32_t c, uint32_t logC) { return (m << logC) + ((c + m) & ((1u << logC) - 1)); }

DEV Fr tile_get(const uint4* slo, const uint4* shi, uint32_t i) {
    uint4 a = slo[i], b = shi[i];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
DEV void tile_put(uint4* slo, uint4* shi, uint32_t i, const Fr& v) {
    slo[i] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    shi[i] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

__global__ void __launch_bounds__(NTT_THREADS) ntt_pass_kernel(PassParams P) {
    extern __shared__ uint4 smem[];
    const uint32_t T = 1u << (P.r + P.logC), C = 1u << P.logC, R = 1u << P.r;
    uint4* slo = smem;
    uint4* shi = smem + T + SMEM_PAD;
    const uint32_t tid = threadIdx.x, tile = blockIdx.x;
    const Fr* in = P.in + (size_t)blockIdx.y * P.batch_stride_in;
    Fr* out = P.out + (size_t)blockIdx.y * P.batch_stride_out;
    const uint32_t s1 = P.s0 + P.r;
    // tile coordinates
    uint32_t hi = 0, lo_base = 0;
    if (!P.first) {
        const uint32_t lo_tiles_log = P.s0 - P.logC;
        hi = tile >> lo_tiles_log;
        lo_base = (tile & ((1u << lo_tiles_log) - 1)) << P.logC;
    }
    // ---- load ----
    // zero-padded to 4x (skip2): after the bit reversal only every fourth row holds data and the first two stages just copy
    // it into the three rows above, so each distinct element is loaded and pre-scaled ONCE and stored four times
    const uint32_t load_rows_log = P.skip2 ? 2 : 0;
    for (uint32_t e = tid; e < (T >> load_rows_log); e += NTT_THREADS) {
        const uint32_t c = e & (C - 1), m = (e >> P.logC) << load_rows_log;
        size_t src;
        if (P.first) {
            const uint32_t hrev = tile * C + c;  // bit-reversed `hi`
            const uint32_t mrev = P.r ? (__brev(m) >> (32 - P.r)) : 0;
            src = ((size_t)mrev << (P.L - P.r)) + hrev;
        } else {
            src = ((size_t)hi << s1) + ((size_t)m << P.s0) + lo_base + c;
        }
        Fr v;
        if (P.in_len && src >= P.in_len) {
            v = f_zero<FrCfg>();
        } else {
            v = f_load(in + (P.first ? src * P.in_stride + P.in_offset : src));
            if (P.pre3) {
                const uint32_t k3 = (uint32_t)(src % 3);
                if (k3) v = f_mul(v, f_load_ro(P.pre3 + k3));
            }
        }
        tile_put(slo, shi, swz(m, c, P.logC), v);
        if (P.skip2) {
            tile_put(slo, shi, swz(m + 1, c, P.logC), v);
            tile_put(slo, shi, swz(m + 2, c, P.logC), v);
            tile_put(slo, shi, swz(m + 3, c, P.logC), v);
        }
    }
    __syncthreads();
    // ---- butterflies ----
    for (uint32_t t = P.skip2 ? 3 : 1; t <= P.r; ++t) {
        const uint32_t half = 1u << (t - 1);
        const uint32_t s = P.s0 + t;
        const bool trivial = P.first && t == 1;  // all twiddles are w^0
#if defined(B200ZK_NTT_ILP2)
        // two butterflies per iteration: both operand pairs and both twiddles are requested before either product starts
        auto bfly_addr = [&](uint32_t b, uint32_t& i0, uint32_t& i1, uint32_t& idx) {
            const uint32_t c = b & (C - 1), mm = b >> P.logC;
            const uint32_t j = mm & (half - 1);
            const uint32_t m0 = ((mm >> (t - 1)) << t) + j, m1 = m0 + half;
            i0 = swz(m0, c, P.logC);
            i1 = swz(m1, c, P.logC);
            const uint32_t lo = P.first ? 0 : lo_base + c;
            idx = (((j << P.s0) + lo) << (P.L - s)) << P.table_shift;
        };
        auto twiddle = [&](uint32_t idx) -> Fr {
            if (!P.inverse) return f_load_ro(P.table + idx);
            if (idx == 0) return f_one<FrCfg>();
            return f_neg(f_load_ro(P.table + (P.half_table - idx)));
        };
        for (uint32_t b = tid; b < T / 2; b += 2 * NTT_THREADS) {
            const uint32_t b2 = b + NTT_THREADS;
            const bool two = b2 < T / 2;
            uint32_t i0, i1, idx, k0 = 0, k1 = 0, kdx = 0;
            bfly_addr(b, i0, i1, idx);
            if (two) bfly_addr(b2, k0, k1, kdx);
            Fr w = trivial ? f_one<FrCfg>() : twiddle(idx), w2 = (trivial || !two) ? f_one<FrCfg>() : twiddle(kdx);
            Fr u = tile_get(slo, shi, i0), v = tile_get(slo, shi, i1);
            Fr u2 = u, v2 = v;
            if (two) {
                u2 = tile_get(slo, shi, k0);
                v2 = tile_get(slo, shi, k1);
            }
            if (!trivial) {
                v = f_mul(v, w);
                if (two) v2 = f_mul(v2, w2);
            }
            tile_put(slo, shi, i0, f_add(u, v));
            tile_put(slo, shi, i1, f_sub(u, v));
            if (two) {
                tile_put(slo, shi, k0, f_add(u2, v2));
                tile_put(slo, shi, k1, f_sub(u2, v2));
            }
        }
#else
        for (uint32_t b = tid; b < T / 2; b += NTT_THREADS) {
            const uint32_t c = b & (C - 1), mm = b >> P.logC;
            const uint32_t j = mm & (half - 1);
            const uint32_t m0 = ((mm >> (t - 1)) << t) + j, m1 = m0 + half;
            const uint32_t i0 = swz(m0, c, P.logC), i1 = swz(m1, c, P.logC);
            Fr u = tile_get(slo, shi, i0), v = tile_get(slo, shi, i1);
            if (!trivial) {
                const uint32_t lo = P.first ? 0 : lo_base + c;
                const uint32_t ex = ((j << P.s0) + lo) << (P.L - s);  // exponent w.r.t. the 2^L-th root
                const uint32_t idx = ex << P.table_shift;
                Fr w;
                if (!P.inverse) {
                    w = f_load_ro(P.table + idx);
                } else if (idx == 0) {
                    w = f_one<FrCfg>();
                } else {
                    w = f_neg(f_load_ro(P.table + (P.half_table - idx)));
                }
                v = f_mul(v, w);
            }
            tile_put(slo, shi, i0, f_add(u, v));
            tile_put(slo, shi, i1, f_sub(u, v));
        }
#endif
        __syncthreads();
    }
    // ---- store ----
    for (uint32_t e = tid; e < T; e += NTT_THREADS) {
        uint32_t c, m;
        size_t dst;
        if (P.first) {
            m = e & (R - 1);
            c = e >> P.r;
            const uint32_t hrev = tile * C + c;
            const uint32_t hbits = P.L - P.r;
            const uint32_t h = hbits ? (__brev(hrev) >> (32 - hbits)) : 0;
            dst = ((size_t)h << P.r) + m;
        } else {
            c = e & (C - 1);
            m = e >> P.logC;
            dst = ((size_t)hi << s1) + ((size_t)m << P.s0) + lo_base + c;
        }
        if (P.last && P.out_len && dst >= P.out_len) continue;
        Fr v = tile_get(slo, shi, swz(m, c, P.logC));
        if (P.last && P.post3) v = f_mul(v, f_load_ro(P.post3 + (uint32_t)(dst % 3)));
        f_store(out + dst, v);
    }
}

// per-device kernel attributes (dynamic shared memory above 48 KiB); called by b200zk_create for its device
void ntt_init_device() {
    const size_t max_smem = ((size_t)2 << (NTT_MAX_R + NTT_MAX_LOGC)) * 16 + SMEM_PAD * 16;
    CUDA_CHECK(cudaFuncSetAttribute(ntt_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
}

int ntt_num_passes(uint32_t log_n) { return log_n == 0 ? 0 : (int)((log_n + NTT_MAX_R - 1) / NTT_MAX_R); }

void ntt_run_batch(const NttPlan& plan, const Fr* in, Fr* out, Fr* scratch, uint32_t batch, size_t stride_in, size_t stride_out,
                   size_t stride_scratch, cudaStream_t stream) {
    const uint32_t L = plan.log_n;
    if (L == 0) {
        if (in != out) CUDA_CHECK(cudaMemcpyAsync(out, in, sizeof(Fr), cudaMemcpyDeviceToDevice, stream));
        return;
    }
    if (plan.table_log < L) throw std::runtime_error("ntt: twiddle table too small");
    const int np = ntt_num_passes(L);
    uint32_t s0 = 0;
    for (int p = 0; p < np; ++p) {
        const uint32_t r = L / np + ((uint32_t)p < L % np ? 1 : 0);
        PassParams P{};
        P.first = p == 0;
        P.last = p == np - 1;
        P.in = P.first ? in : scratch;
        P.out = P.last ? out : scratch;
        P.batch_stride_in = P.first ? stride_in : stride_scratch;
        P.batch_stride_out = P.last ? stride_out : stride_scratch;
        P.table = plan.table;
        P.table_shift = plan.table_log - L;
        P.half_table = 1u << (plan.table_log - 1);
        P.inverse = plan.inverse;
        P.L = L;
        P.s0 = s0;
        P.r = r;
        const uint32_t avail = P.first ? L - r : s0;  // log2 of the number of columns that exist
        P.logC = avail < (uint32_t)NTT_MAX_LOGC ? avail : NTT_MAX_LOGC;
        P.in_stride = plan.in_stride;
        P.in_offset = plan.in_offset;
        P.pre3 = P.first ? plan.pre_scale3 : nullptr;
        P.in_len = P.first ? plan.in_len : 0;
        // zero-padded to 4x (coeff_to_extended): positions with (m & 3) != 0 hold zeros after the bit reversal, so the first two
        // stages turn [a,0,0,0] into [a,a,a,a] — load that directly and start at stage 3
        P.skip2 = P.first && r >= 2 && plan.in_len != 0 && plan.in_len * 4 == ((size_t)1 << L) ? 1u : 0u;
        P.post3 = P.last ? plan.post_scale3 : nullptr;
        P.out_len = P.last ? plan.out_len : 0;
        const uint32_t T = 1u << (r + P.logC);
        const size_t smem = (size_t)2 * T * 16 + SMEM_PAD * 16;
        dim3 grid(1u << (L - r - P.logC), batch);
        const int prof_h = prof_begin(PROF_NTT_PASS, stream, (double)batch * (double)((size_t)1 << (L - 1)) * (r - (P.skip2 ? 2 : 0)));
        ntt_pass_kernel<<<grid, NTT_THREADS, smem, stream>>>(P);
        prof_end(prof_h, stream);
        ++g_launch_count;
        CUDA_CHECK(cudaGetLastError());
        s0 += r;
    }
}

// ---- twiddle table: T[i] = w^i, i < 2^(log_n-1), from two small power tables -----------------------------
__global__ void twiddle_small_kernel(Fr* lo, Fr* hi, Fr omega, uint32_t lo_bits, uint32_t count_hi) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nlo = 1u << lo_bits;
    if (i < nlo) f_store(lo + i, f_pow_u64(omega, i));
    if (i < count_hi) f_store(hi + i, f_pow_u64(omega, (uint64_t)i << lo_bits));
}
__global__ void twiddle_expand_kernel(Fr* table, const Fr* lo, const Fr* hi, uint32_t lo_bits, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t a = i & ((1u << lo_bits) - 1), b = i >> lo_bits;
    Fr v = f_load_ro(lo + a);
    if (b) v = f_mul(v, f_load_ro(hi + b));
    f_store(table + i, v);
}
void build_twiddle_table(Fr* table, const Fr& omega, uint32_t log_n, cudaStream_t stream) {
    const uint32_t n = log_n == 0 ? 1 : 1u << (log_n - 1);
    const uint32_t lo_bits = log_n > 11 ? 10 : (log_n > 1 ? (log_n - 1) : 0);
    const uint32_t nlo = 1u << lo_bits, nhi = (n + nlo - 1) >> lo_bits;
    DevBuf<Fr> lo(nlo, stream), hi(nhi, stream);
    const uint32_t m = nlo > nhi ? nlo : nhi;
    twiddle_small_kernel<<<(m + 127) / 128, 128, 0, stream>>>(lo.get(), hi.get(), omega, lo_bits, nhi);
    twiddle_expand_kernel<<<(n + 255) / 256, 256, 0, stream>>>(table, lo.get(), hi.get(), lo_bits, n);
    g_launch_count += 2;
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace b200zk
