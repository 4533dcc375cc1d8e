// Column / polynomial kernels between the big primitives of create_proof (SURVEY.md §8a rows E, F, G, I;
// upstream: halo2_proofs::{arithmetic::{eval_polynomial, kate_division}, plonk::permutation::prover,
// plonk::lookup::prover, poly::Polynomial ops}, halo2curves batch_invert). All are HBM-streaming kernels with
// 1–5 Montgomery products per element; sequential recurrences (running products, synthetic division) are cut
// into per-thread chunks and stitched by a block-level + recursive grid-level scan of the chunk summaries.
#include "poly.cuh"

namespace b200zk {

void exclusive_scan_u32(const uint32_t* in, uint32_t* out, size_t n, cudaStream_t stream);

static inline unsigned nblocks(size_t n, unsigned t) { return (unsigned)((n + t - 1) / t); }
#define LAUNCHED(k) do { g_launch_count += (k); CUDA_CHECK(cudaGetLastError()); } while (0)

// ---- elementwise -------------------------------------------------------------------------------------------------
__global__ void fill_kernel(Fr* a, Fr v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) f_store(a + i, v);
}
void fr_fill(Fr* a, const Fr& v, size_t n, cudaStream_t s) {
    if (!n) return;
    fill_kernel<<<nblocks(n, 256), 256, 0, s>>>(a, v, n);
    LAUNCHED(1);
}
__global__ void scale_kernel(Fr* a, Fr c, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) f_store(a + i, f_mul(f_load(a + i), c));
}
void fr_scale(Fr* a, const Fr& c, size_t n, cudaStream_t s) {
    if (!n) return;
    scale_kernel<<<nblocks(n, 256), 256, 0, s>>>(a, c, n);
    LAUNCHED(1);
}
constexpr int LC_MAX = 48;
struct LincombArgs {
    const Fr* polys[LC_MAX];
    Fr coeffs[LC_MAX];
    uint32_t m;
    uint32_t accumulate;
};
__global__ void __launch_bounds__(256) lincomb_kernel(Fr* out, LincombArgs A, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr acc = A.accumulate ? f_load(out + i) : f_zero<FrCfg>();
    for (uint32_t j = 0; j < A.m; ++j) acc = f_add(acc, f_mul(f_load(A.polys[j] + i), A.coeffs[j]));
    f_store(out + i, acc);
}
void fr_lincomb(Fr* out, const std::vector<const Fr*>& polys, const std::vector<Fr>& coeffs, size_t n, bool accumulate, cudaStream_t s) {
    if (!n) return;
    if (polys.empty() && !accumulate) {
        CUDA_CHECK(cudaMemsetAsync(out, 0, n * sizeof(Fr), s));
        return;
    }
    for (size_t off = 0; off < polys.size(); off += LC_MAX) {
        LincombArgs A;
        A.m = (uint32_t)std::min<size_t>(LC_MAX, polys.size() - off);
        A.accumulate = accumulate || off > 0;
        for (uint32_t j = 0; j < A.m; ++j) {
            A.polys[j] = polys[off + j];
            A.coeffs[j] = coeffs[off + j];
        }
        lincomb_kernel<<<nblocks(n, 256), 256, 0, s>>>(out, A, n);
        LAUNCHED(1);
    }
}
struct SmallArgs {
    Fr v[8];
    uint32_t m;
};
__global__ void sub_low_kernel(Fr* a, SmallArgs S) {
    uint32_t i = threadIdx.x;
    if (i < S.m) f_store(a + i, f_sub(f_load(a + i), S.v[i]));
}
void fr_sub_low(Fr* a, const Fr* small_host, uint32_t m, cudaStream_t s) {
    if (!m) return;
    if (m > 8) throw std::invalid_argument("fr_sub_low: m > 8");
    SmallArgs S;
    S.m = m;
    for (uint32_t i = 0; i < m; ++i) S.v[i] = small_host[i];
    sub_low_kernel<<<1, 32, 0, s>>>(a, S);
    LAUNCHED(1);
}

// ---- batch inversion -----------------------------------------------------------------------------------------------
// Montgomery's trick over strided groups: thread g owns elements g, g+G, g+2G, ... (coalesced); prefix products go
// through a scratch column. Zeros are skipped (treated as one) and stay zero, like halo2's batch_invert.
__global__ void __launch_bounds__(128) batch_invert_kernel(Fr* a, Fr* scratch, size_t n, size_t G) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    Fr acc = f_one<FrCfg>();
    for (size_t i = g; i < n; i += G) {
        f_store(scratch + i, acc);
        const Fr v = f_load(a + i);
        if (!f_is_zero(v)) acc = f_mul(acc, v);
    }
    acc = f_inv(acc);
    size_t cnt = (n - g + G - 1) / G;
    for (size_t j = cnt; j-- > 0;) {
        const size_t i = g + j * G;
        const Fr v = f_load(a + i);
        if (f_is_zero(v)) continue;
        f_store(a + i, f_mul(acc, f_load(scratch + i)));
        acc = f_mul(acc, v);
    }
}
void fr_batch_invert(Fr* a, size_t n, cudaStream_t s) {
    if (!n) return;
    // elements per inversion: longer groups amortise the ~380-product inversion, shorter ones keep enough threads in flight
    size_t per = n >= ((size_t)1 << 22) ? 128 : (n >= ((size_t)1 << 18) ? 64 : (n >= 4096 ? 16 : 4));
    size_t G = (n + per - 1) / per;
    DevBuf<Fr> scratch(n, s);
    batch_invert_kernel<<<nblocks(G, 128), 128, 0, s>>>(a, scratch.get(), n, G);
    LAUNCHED(1);
}

// ---- exclusive running product --------------------------------------------------------------------------------------
constexpr int SCAN_E = 8;      // elements per thread
constexpr int SCAN_T = 256;    // threads per block
constexpr int SCAN_TILE = SCAN_E * SCAN_T;
// Phase 1: per-thread totals, block-level exclusive scan of the totals -> thread_prefix[t], block_total[b]
__global__ void __launch_bounds__(SCAN_T) prodscan_phase1(const Fr* m, size_t n, Fr* thread_prefix, Fr* block_total) {
    __shared__ Fr sh[2][SCAN_T];
    const size_t t = (size_t)blockIdx.x * SCAN_T + threadIdx.x;
    const size_t base = t * SCAN_E;
    Fr tot = f_one<FrCfg>();
    for (int j = 0; j < SCAN_E; ++j)
        if (base + j < n) tot = f_mul(tot, f_load(m + base + j));
    int cur = 0;
    sh[0][threadIdx.x] = tot;
    __syncthreads();
    for (int d = 1; d < SCAN_T; d <<= 1) {  // Hillis–Steele inclusive scan
        Fr v = sh[cur][threadIdx.x];
        if ((int)threadIdx.x >= d) v = f_mul(sh[cur][threadIdx.x - d], v);
        sh[cur ^ 1][threadIdx.x] = v;
        cur ^= 1;
        __syncthreads();
    }
    const Fr excl = threadIdx.x == 0 ? f_one<FrCfg>() : sh[cur][threadIdx.x - 1];
    if (base < n) f_store(thread_prefix + t, excl);
    if (threadIdx.x == SCAN_T - 1) f_store(block_total + blockIdx.x, sh[cur][SCAN_T - 1]);
}
// Phase 3: z[i] = block_prefix[b] * thread_prefix[t] * prod_{local j < i} m[j]
__global__ void __launch_bounds__(SCAN_T) prodscan_phase3(const Fr* m, size_t n, const Fr* thread_prefix, const Fr* block_prefix, Fr* z) {
    const size_t t = (size_t)blockIdx.x * SCAN_T + threadIdx.x;
    const size_t base = t * SCAN_E;
    if (base >= n) return;
    Fr acc = f_mul(f_load(block_prefix + blockIdx.x), f_load(thread_prefix + t));
    for (int j = 0; j < SCAN_E; ++j) {
        if (base + j >= n) break;
        f_store(z + base + j, acc);
        acc = f_mul(acc, f_load(m + base + j));
    }
}
void fr_prefix_product(Fr* z, const Fr* m, const Fr& first, size_t n, cudaStream_t s) {
    if (!n) return;
    const size_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    const size_t nt = (n + SCAN_E - 1) / SCAN_E;
    DevBuf<Fr> thread_prefix(nt, s), block_total(nb, s), block_prefix(nb, s);
    prodscan_phase1<<<(unsigned)nb, SCAN_T, 0, s>>>(m, n, thread_prefix.get(), block_total.get());
    LAUNCHED(1);
    if (nb == 1) {
        fr_fill(block_prefix.get(), first, 1, s);
    } else {
        fr_prefix_product(block_prefix.get(), block_total.get(), first, nb, s);
    }
    prodscan_phase3<<<(unsigned)nb, SCAN_T, 0, s>>>(m, n, thread_prefix.get(), block_prefix.get(), z);
    LAUNCHED(1);
}

// ---- evaluation ------------------------------------------------------------------------------------------------------
constexpr int EV_E = 16, EV_T = 256;
struct EvalArgs {
    const Fr* polys[LC_MAX];
};
// partial[p][block] = sum over the block's tile of c_i x^i ; xe_table[t] = (x^EV_E)^t
__global__ void __launch_bounds__(EV_T) eval_partial_kernel(EvalArgs A, size_t n, Fr x, const Fr* xe_table, Fr* partial) {
    __shared__ Fr sh[EV_T];
    const Fr* c = A.polys[blockIdx.y];
    const size_t t = (size_t)blockIdx.x * EV_T + threadIdx.x;
    const size_t base = t * EV_E;
    Fr acc = f_zero<FrCfg>();
    if (base < n) {
        const int cnt = (int)(n - base < (size_t)EV_E ? n - base : EV_E);
        for (int j = cnt - 1; j >= 0; --j) acc = f_add(f_mul(acc, x), f_load(c + base + j));
        acc = f_mul(acc, f_load_ro(xe_table + t));
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int d = EV_T / 2; d > 0; d >>= 1) {
        if ((int)threadIdx.x < d) sh[threadIdx.x] = f_add(sh[threadIdx.x], sh[threadIdx.x + d]);
        __syncthreads();
    }
    if (threadIdx.x == 0) f_store(partial + (size_t)blockIdx.y * gridDim.x + blockIdx.x, sh[0]);
}
__global__ void __launch_bounds__(EV_T) eval_final_kernel(const Fr* partial, uint32_t nparts, Fr* out) {
    __shared__ Fr sh[EV_T];
    Fr acc = f_zero<FrCfg>();
    for (uint32_t i = threadIdx.x; i < nparts; i += EV_T) acc = f_add(acc, f_load(partial + (size_t)blockIdx.x * nparts + i));
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int d = EV_T / 2; d > 0; d >>= 1) {
        if ((int)threadIdx.x < d) sh[threadIdx.x] = f_add(sh[threadIdx.x], sh[threadIdx.x + d]);
        __syncthreads();
    }
    if (threadIdx.x == 0) f_store(out + blockIdx.x, sh[0]);
}
void fr_eval_many(Context& ctx, const std::vector<const Fr*>& polys, size_t n, const Fr& point, Fr* out_host) {
    if (polys.empty()) return;
    cudaStream_t s = ctx.stream;
    const size_t nt = (n + EV_E - 1) / EV_E;
    uint32_t tl = 1;
    while (((size_t)1 << (tl - 1)) < nt) ++tl;  // table of 2^(tl-1) >= nt powers
    DevBuf<Fr> xe(((size_t)1 << (tl - 1)), s);
    build_twiddle_table(xe.get(), f_pow_u64(point, EV_E), tl, s);
    const unsigned nbx = nblocks(nt, EV_T);
    DevBuf<Fr> partial((size_t)nbx * LC_MAX, s), out(LC_MAX, s);
    for (size_t off = 0; off < polys.size(); off += LC_MAX) {
        EvalArgs A;
        const uint32_t m = (uint32_t)std::min<size_t>(LC_MAX, polys.size() - off);
        for (uint32_t j = 0; j < m; ++j) A.polys[j] = polys[off + j];
        eval_partial_kernel<<<dim3(nbx, m), EV_T, 0, s>>>(A, n, point, xe.get(), partial.get());
        eval_final_kernel<<<m, EV_T, 0, s>>>(partial.get(), nbx, out.get());
        LAUNCHED(2);
        CUDA_CHECK(cudaMemcpyAsync(out_host + off, out.get(), m * sizeof(Fr), cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaStreamSynchronize(s));
    }
}

// ---- synthetic division by (X - b) -------------------------------------------------------------------------------------
// s_j = a[j+1] + b·s_{j+1}, s_{n-1} = 0, q[j] = s_j. Threads own chunks of KD_E consecutive j; chunk summaries
// A_t (value at the chunk's low end for zero inflow) obey S_t = A_t + b^E·S_{t+1}: a suffix Horner scan.
constexpr int KD_E = 8, KD_T = 256;
// generic suffix Horner scan: S_t = A_t + B·S_{t+1} over T values, S_T = 0. In place over A.
__global__ void __launch_bounds__(KD_T) suffix_horner_phase1(Fr* A, size_t T, Fr B, Fr* block_head) {
    __shared__ Fr sh[2][KD_T];
    const size_t t = (size_t)blockIdx.x * KD_T + threadIdx.x;
    int cur = 0;
    sh[0][threadIdx.x] = t < T ? f_load(A + t) : f_zero<FrCfg>();
    __syncthreads();
    Fr bp = B;  // B^(d)
    for (int d = 1; d < KD_T; d <<= 1) {
        Fr v = sh[cur][threadIdx.x];
        if ((int)threadIdx.x + d < KD_T) v = f_add(v, f_mul(bp, sh[cur][threadIdx.x + d]));
        sh[cur ^ 1][threadIdx.x] = v;
        cur ^= 1;
        bp = f_sqr(bp);
        __syncthreads();
    }
    if (t < T) f_store(A + t, sh[cur][threadIdx.x]);
    if (threadIdx.x == 0) f_store(block_head + blockIdx.x, sh[cur][0]);
}
// powers[j] = B^(j+1), j < KD_T
__global__ void __launch_bounds__(KD_T) power_table_kernel(Fr* powers, Fr B) {
    f_store(powers + threadIdx.x, f_pow_u64(B, (uint64_t)threadIdx.x + 1));
}
// S_t += B^(KD_T - threadIdx)·S_next_block_start
__global__ void __launch_bounds__(KD_T) suffix_horner_phase3(Fr* A, size_t T, const Fr* powers, const Fr* block_S, size_t nblocks_) {
    const size_t t = (size_t)blockIdx.x * KD_T + threadIdx.x;
    if (t >= T || blockIdx.x + 1 >= nblocks_) return;
    const Fr inflow = f_load(block_S + blockIdx.x + 1);
    const Fr p = f_load_ro(powers + (KD_T - 1 - threadIdx.x));  // B^(KD_T - threadIdx)
    f_store(A + t, f_add(f_load(A + t), f_mul(p, inflow)));
}
static void suffix_horner_scan(Fr* A, size_t T, const Fr& B, cudaStream_t s) {
    if (!T) return;
    const size_t nb = (T + KD_T - 1) / KD_T;
    DevBuf<Fr> heads(nb, s);
    suffix_horner_phase1<<<(unsigned)nb, KD_T, 0, s>>>(A, T, B, heads.get());
    LAUNCHED(1);
    if (nb > 1) {
        DevBuf<Fr> powers(KD_T, s);
        power_table_kernel<<<1, KD_T, 0, s>>>(powers.get(), B);
        suffix_horner_scan(heads.get(), nb, f_pow_u64(B, KD_T), s);
        suffix_horner_phase3<<<(unsigned)nb, KD_T, 0, s>>>(A, T, powers.get(), heads.get(), nb);
        LAUNCHED(2);
    }
}
__global__ void __launch_bounds__(256) kate_local_kernel(const Fr* a, size_t n, Fr b, Fr* chunk_A) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t lo = t * KD_E;
    if (lo >= n) return;
    const size_t hi = lo + KD_E < n ? lo + KD_E : n;  // j in [lo, hi)
    Fr s = f_zero<FrCfg>();
    for (size_t j = hi; j-- > lo;) {
        const Fr aj1 = j + 1 < n ? f_load(a + j + 1) : f_zero<FrCfg>();
        s = f_add(aj1, f_mul(b, s));
    }
    f_store(chunk_A + t, s);
}
__global__ void __launch_bounds__(256) kate_final_kernel(const Fr* a, size_t n, Fr b, const Fr* chunk_S, size_t T, Fr* q) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t lo = t * KD_E;
    if (lo >= n) return;
    const size_t hi = lo + KD_E < n ? lo + KD_E : n;
    Fr s = t + 1 < T ? f_load(chunk_S + t + 1) : f_zero<FrCfg>();
    for (size_t j = hi; j-- > lo;) {
        const Fr aj1 = j + 1 < n ? f_load(a + j + 1) : f_zero<FrCfg>();
        s = f_add(aj1, f_mul(b, s));
        f_store(q + j, s);
    }
}
void fr_kate_division(Context& ctx, const Fr* a, Fr* q, size_t n, const Fr& b) {
    if (!n) return;
    if (a == q) throw std::invalid_argument("kate_division: output must not alias input");
    cudaStream_t s = ctx.stream;
    const size_t T = (n + KD_E - 1) / KD_E;
    DevBuf<Fr> chunk(T, s);
    kate_local_kernel<<<nblocks(T, 256), 256, 0, s>>>(a, n, b, chunk.get());
    LAUNCHED(1);
    suffix_horner_scan(chunk.get(), T, f_pow_u64(b, KD_E), s);
    kate_final_kernel<<<nblocks(T, 256), 256, 0, s>>>(a, n, b, chunk.get(), T, q);
    LAUNCHED(1);
}

// ---- permutation / lookup products -----------------------------------------------------------------------------------
__global__ void perm_den_kernel(Fr* m, const Fr* v, const Fr* sigma, Fr beta, Fr gamma, size_t n, int first) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr t = f_add(f_add(f_mul(beta, f_load(sigma + i)), gamma), f_load(v + i));
    if (!first) t = f_mul(f_load(m + i), t);
    f_store(m + i, t);
}
void perm_denominator(Fr* m, const Fr* v, const Fr* sigma, const Fr& beta, const Fr& gamma, size_t n, bool first, cudaStream_t s) {
    perm_den_kernel<<<nblocks(n, 256), 256, 0, s>>>(m, v, sigma, beta, gamma, n, first);
    LAUNCHED(1);
}
__global__ void perm_num_kernel(Fr* m, const Fr* v, Fr dpb, Fr gamma, const Fr* table, uint32_t table_log, uint32_t k) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ((size_t)1 << k)) return;
    const Fr w = omega_pow_from_table(table, table_log, k, (uint32_t)i);
    Fr t = f_add(f_add(f_mul(dpb, w), gamma), f_load(v + i));
    f_store(m + i, f_mul(f_load(m + i), t));
}
void perm_numerator(Fr* m, const Fr* v, const Fr& delta_pow_beta, const Fr& gamma, const Fr* table, uint32_t table_log, uint32_t k, cudaStream_t s) {
    perm_num_kernel<<<nblocks((size_t)1 << k, 256), 256, 0, s>>>(m, v, delta_pow_beta, gamma, table, table_log, k);
    LAUNCHED(1);
}
__global__ void lookup_den_kernel(Fr* p, const Fr* a, const Fr* sp, Fr beta, Fr gamma, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    f_store(p + i, f_mul(f_add(beta, f_load(a + i)), f_add(gamma, f_load(sp + i))));
}
void lookup_denominator(Fr* p, const Fr* a, const Fr* sp, const Fr& beta, const Fr& gamma, size_t n, cudaStream_t s) {
    lookup_den_kernel<<<nblocks(n, 256), 256, 0, s>>>(p, a, sp, beta, gamma, n);
    LAUNCHED(1);
}
__global__ void lookup_num_kernel(Fr* p, const Fr* in, const Fr* tab, Fr beta, Fr gamma, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr t = f_mul(f_load(p + i), f_add(f_load(in + i), beta));
    f_store(p + i, f_mul(t, f_add(f_load(tab + i), gamma)));
}
void lookup_numerator(Fr* p, const Fr* in, const Fr* tab, const Fr& beta, const Fr& gamma, size_t n, cudaStream_t s) {
    lookup_num_kernel<<<nblocks(n, 256), 256, 0, s>>>(p, in, tab, beta, gamma, n);
    LAUNCHED(1);
}

// ---- lookup permutation (permute_expression_pair) ----------------------------------------------------------------------
// Values are table entries < n, so the sort is a counting sort over the value domain [0, n):
//   hist_in / hist_tab -> start_in (scan), distinct-inclusive count D, leftover counts (scan) ->
//   a'[row] = value whose run contains row; s'[row] = a'[row] on the first row of a run, else the leftover table value
//   number R-1-idx in ascending order, idx = rank of the row among repeated rows (upstream pops repeated rows from
//   the end while walking leftovers in ascending order).
__global__ void lp_hist_kernel(const Fr* col, size_t usable, uint32_t n, uint32_t* hist, uint32_t* err, uint32_t err_bit) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= usable) return;
    const Fr v = f_from_mont(f_load(col + i));
    uint32_t hi = 0;
#pragma unroll
    for (int t = 1; t < 8; ++t) hi |= v.l[t];
    if (hi != 0 || v.l[0] >= n) {
        atomicOr(err, err_bit);  // outside the value domain [0, n) this counting sort covers
        return;
    }
    atomicAdd(hist + v.l[0], 1u);
}
// per value: distinct flag and leftover count; flags an input value missing from the table
__global__ void lp_value_kernel(const uint32_t* hist_in, const uint32_t* hist_tab, uint32_t n, uint32_t* distinct, uint32_t* left, uint32_t* err) {
    uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v > n) return;
    if (v == n) {
        distinct[v] = 0;
        left[v] = 0;
        return;
    }
    const uint32_t ci = hist_in[v], ct = hist_tab[v];
    const uint32_t d = ci > 0 ? 1u : 0u;
    if (d && ct == 0) atomicOr(err, 2u);  // input value missing from the table
    distinct[v] = d;
    left[v] = ct >= d ? ct - d : 0;
}
DEV uint32_t upper_idx(const uint32_t* starts, uint32_t n, uint32_t p) {  // largest v in [0,n) with starts[v] <= p
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(starts + mid) <= p) lo = mid;
        else hi = mid;
    }
    return lo;
}
__global__ void lp_write_kernel(const uint32_t* start_in, const uint32_t* dist_excl, const uint32_t* left_start, uint32_t n, size_t usable,
                                uint32_t R, Fr* a_out, Fr* s_out, int fill_from_end) {
    size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= usable) return;
    const uint32_t v = upper_idx(start_in, n, (uint32_t)row);  // start_in[v] <= row < start_in[v+1] (runs are non-empty)
    Fr fv = f_zero<FrCfg>();
    fv.l[0] = v;
    fv = f_to_mont(fv);
    f_store(a_out + row, fv);
    if (start_in[v] == row) {
        f_store(s_out + row, fv);
    } else {
        const uint32_t d_incl = dist_excl[v] + 1;                 // distinct values <= v
        const uint32_t idx = (uint32_t)row - d_incl;              // rank among repeated rows (ascending)
        const uint32_t j = fill_from_end ? R - 1 - idx : idx;     // leftover number j (ascending) [UNVERIFIED-2]
        const uint32_t t = upper_idx(left_start, n, j);
        Fr ft = f_zero<FrCfg>();
        ft.l[0] = t;
        f_store(s_out + row, f_to_mont(ft));
    }
}
int lookup_permute(Context& ctx, const Fr* input, const Fr* table, Fr* a_out, Fr* s_out, size_t n, size_t usable, bool fill_from_end) {
    cudaStream_t s = ctx.stream;
    const uint32_t N = (uint32_t)n;
    DevBuf<uint32_t> hist_in(N + 1, s), hist_tab(N + 1, s), distinct(N + 1, s), left(N + 1, s), err(1, s);
    CUDA_CHECK(cudaMemsetAsync(hist_in.get(), 0, (N + 1) * 4, s));
    CUDA_CHECK(cudaMemsetAsync(hist_tab.get(), 0, (N + 1) * 4, s));
    CUDA_CHECK(cudaMemsetAsync(err.get(), 0, 4, s));
    // a TABLE value >= n is outside what this sort supports; an INPUT value >= n is then simply missing from the table
    lp_hist_kernel<<<nblocks(usable, 256), 256, 0, s>>>(input, usable, N, hist_in.get(), err.get(), (uint32_t)LOOKUP_NOT_IN_TABLE);
    lp_hist_kernel<<<nblocks(usable, 256), 256, 0, s>>>(table, usable, N, hist_tab.get(), err.get(), (uint32_t)LOOKUP_UNSUPPORTED);
    lp_value_kernel<<<nblocks(N + 1, 256), 256, 0, s>>>(hist_in.get(), hist_tab.get(), N, distinct.get(), left.get(), err.get());
    LAUNCHED(3);
    exclusive_scan_u32(hist_in.get(), hist_in.get(), N + 1, s);    // start_in
    exclusive_scan_u32(distinct.get(), distinct.get(), N + 1, s);  // distinct values < v
    exclusive_scan_u32(left.get(), left.get(), N + 1, s);          // leftover start
    uint32_t h[2] = {0, 0};
    CUDA_CHECK(cudaMemcpyAsync(&h[0], err.get(), 4, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaMemcpyAsync(&h[1], left.get() + N, 4, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    if (h[0] != 0) return (int)h[0];
    // start_in has empty runs (equal consecutive starts); upper_idx picks the last value whose start <= row, which is
    // the non-empty run containing the row. left_start likewise.
    lp_write_kernel<<<nblocks(usable, 256), 256, 0, s>>>(hist_in.get(), distinct.get(), left.get(), N, usable, h[1], a_out, s_out, fill_from_end ? 1 : 0);
    LAUNCHED(1);
    return 0;
}

// ---- radix-4 combine of four size-n inverse transforms into extended_to_coeff's output --------------------------------------
// x (4n values on the extended domain) = four stride-4 subsequences x_j[m] = x[4m + j]; with Y_j = the plain inverse
// transform of x_j (root ω⁻¹ = W⁴, W = ω_ext⁻¹, no 1/n), the size-4n inverse transform is
//   X[k + q·n] = Σ_j W^{j(k + qn)} Y_j[k] = Σ_j (I⁴_q)^j · (W^{jk} Y_j[k]),  I = W^n (a primitive fourth root of unity),
// so with z_j = W^{jk}·Y_j[k]:  X[k] = z0+z1+z2+z3,  X[k+n] = (z0−z2) + I·(z1−z3),  X[k+2n] = (z0+z2) − (z1+z3)
// (X[k+3n] is dropped by extended_to_coeff's truncation to 3n). The post factors (1/4n)·ζ^(−i mod 3) are applied here.
__global__ void e2c_combine_kernel(const Fr* Y, Fr* out, size_t n, size_t k_lo, size_t k_hi, const Fr* table, uint32_t table_log, uint32_t ext_k, Fr I,
                                   const Fr* post3) {
    const size_t k = k_lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= k_hi) return;
    const uint32_t en = 1u << ext_k;
    auto w_inv = [&](uint32_t e) -> Fr {  // ω_ext^(−e), e < 4n
        if (e == 0) return f_one<FrCfg>();
        return omega_pow_from_table(table, table_log, ext_k, en - e);
    };
    const Fr z0 = f_load(Y + k);
    const Fr z1 = f_mul(f_load(Y + n + k), w_inv((uint32_t)k));
    const Fr z2 = f_mul(f_load(Y + 2 * n + k), w_inv((uint32_t)(2 * k)));
    const Fr z3 = f_mul(f_load(Y + 3 * n + k), w_inv((uint32_t)(3 * k)));
    const Fr s02 = f_add(z0, z2), d02 = f_sub(z0, z2), s13 = f_add(z1, z3), d13 = f_mul(f_sub(z1, z3), I);
    const Fr x0 = f_add(s02, s13), x1 = f_add(d02, d13), x2 = f_sub(s02, s13);
    const size_t i0 = k, i1 = k + n, i2 = k + 2 * n;
    f_store(out + i0, f_mul(x0, f_load_ro(post3 + (uint32_t)(i0 % 3))));
    f_store(out + i1, f_mul(x1, f_load_ro(post3 + (uint32_t)(i1 % 3))));
    f_store(out + i2, f_mul(x2, f_load_ro(post3 + (uint32_t)(i2 % 3))));
}
void fr_e2c_combine(const Fr* Y, Fr* out, size_t n, size_t k_lo, size_t k_hi, const Fr* table, uint32_t table_log, uint32_t ext_k, const Fr& I, const Fr* post3,
                    cudaStream_t s) {
    if (k_hi <= k_lo) return;
    e2c_combine_kernel<<<nblocks(k_hi - k_lo, 256), 256, 0, s>>>(Y, out, n, k_lo, k_hi, table, table_log, ext_k, I, post3);
    LAUNCHED(1);
}

// ---- Fr::random stream (ChaCha) ------------------------------------------------------------------------------------------
DEV uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
struct ChaChaKey {
    uint32_t k[8];
};
__global__ void random_stream_kernel(Fr* out, size_t n, ChaChaKey key, unsigned long long counter0, int rounds) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long ctr = counter0 + i;
    uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key.k[0], key.k[1], key.k[2], key.k[3],
                       key.k[4], key.k[5], key.k[6], key.k[7], (uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
    uint32_t x[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] = in[j];
#define QR(a, b, c, d)                            \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 16); \
    x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 12); \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 8);  \
    x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 7);
    for (int r = 0; r < rounds; r += 2) {
        QR(0, 4, 8, 12) QR(1, 5, 9, 13) QR(2, 6, 10, 14) QR(3, 7, 11, 15)
        QR(0, 5, 10, 15) QR(1, 6, 11, 12) QR(2, 7, 8, 13) QR(3, 4, 9, 14)
    }
#undef QR
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] += in[j];
    f_store(out + i, f_from_u512<FrCfg>(x));
}
// out[i] = from_u512(words[16 i .. 16 i + 16)) for host-supplied random words (an external RngCore); in place is fine
// (`words` may alias the first half of a 64·n-byte buffer whose prefix is `out`? no: separate buffers)
__global__ void from_u512_kernel(Fr* out, const uint32_t* words, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t w[16];
    const uint4* q = reinterpret_cast<const uint4*>(words + 16 * i);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint4 v = q[j];
        w[4 * j] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w;
    }
    f_store(out + i, f_from_u512<FrCfg>(w));
}
void fr_from_u512(Fr* out, const uint32_t* words_dev, size_t n, cudaStream_t s) {
    if (!n) return;
    from_u512_kernel<<<nblocks(n, 128), 128, 0, s>>>(out, words_dev, n);
    LAUNCHED(1);
}
void fr_random_stream(Fr* out, size_t n, const uint32_t key[8], uint64_t counter0, int rounds, cudaStream_t s) {
    if (!n) return;
    ChaChaKey k;
    memcpy(k.k, key, 32);
    random_stream_kernel<<<nblocks(n, 128), 128, 0, s>>>(out, n, k, counter0, rounds);
    LAUNCHED(1);
}

// ---- sigma columns -----------------------------------------------------------------------------------------------------
__global__ void sigma_kernel(Fr* sigma, const uint32_t* map_col, const uint32_t* map_row, const Fr* delta_pows, const Fr* table,
                             uint32_t table_log, uint32_t k) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ((size_t)1 << k)) return;
    const Fr w = omega_pow_from_table(table, table_log, k, map_row[i]);
    f_store(sigma + i, f_mul(f_load_ro(delta_pows + map_col[i]), w));
}
void sigma_from_mapping(Fr* sigma, const uint32_t* map_col, const uint32_t* map_row, const Fr* delta_pows_dev, const Fr* table,
                        uint32_t table_log, uint32_t k, cudaStream_t s) {
    sigma_kernel<<<nblocks((size_t)1 << k, 256), 256, 0, s>>>(sigma, map_col, map_row, delta_pows_dev, table, table_log, k);
    LAUNCHED(1);
}

}  // namespace b200zk
