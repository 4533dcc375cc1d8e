// G1 Pippenger MSM for sm_100a (replaces halo2curves::msm::best_multiexp as called by
// ParamsKZG::{commit, commit_lagrange} from create_proof — reference entry verifier/src/stark/mod.rs:543,593).
//
// Only the group element Σ sᵢ·Pᵢ matters (canonical affine at the boundary), so the decomposition is free:
//   0. tables   : for the SRS bases T[w][i] = 2^(c·w)·P_i is built once (c = k−3 up to k = 20: 15 windows at k = 20;
//                 msm_table_window_bits), so every window of a column lands in ONE set of 2^(c-1) buckets and no doubling
//                 chain is left; caller-supplied bases keep one bucket set per window and a short host-side fold.
//   1. digits   : scalars leave Montgomery form; each is cut into W signed c-bit digits dₗ ∈ [-2^(c-1), 2^(c-1)];
//                 zero digits create no work (advice-like scalars are mostly < 2^84, lookup columns < 2^(k-1)).
//   2. sort     : counting sort by bucket: histogram with warp-aggregated atomics, exclusive scan, scatter. Order inside
//                 a bucket is irrelevant to the group sum, so the result stays deterministic.
//   3. accumulate: the sorted entry list is cut into equal chunks of T entries, one thread each (perfect balance
//                 whatever the bucket histogram: hot buckets simply span many chunks). A thread walks its chunk
//                 with one XYZZ accumulator and mixed additions (6 products + 2 squares + 1 dual product); runs that begin
//                 inside the chunk are stored to their bucket, the run that began earlier goes to a "head" list, which is
//                 segment-summed one partial per lane with a shuffle scan (fan-in 32 per level; msm_reduce.cu) until one
//                 warp covers it. Optional batched-affine pre-reduction rounds exist and are off (measured slower).
//   4. reduce   : (msm_reduce.cu) Σ b·B_b via recursive chunked running sums (2 additions per bucket, Horner in the chunk
//                 size) down to 1024 entries per set, then one tail launch (block-wide suffix scan); done ONCE for all
//                 columns of a commit batch — its deep levels are latency bound.
// Up to four columns of a batch are in flight on separate streams so that the latency-bound phases of one column
// (atomics, scans, the entry-count read-back, short combine levels) hide under another column's accumulate.
// A batch may mix the two SRS bases (a base pointer per column) and may be issued while its columns are still arriving
// (Context::column_gate). Multi-GPU: q·world columns are dealt by column, the remainder is split by point range and the
// partial sums are all-gathered over NCCL straight from the reduction's output (see msm_batch_distribute).
// Bound: the FMA-heavy (IMAD) pipe — 73 % busy in msm_accumulate_kernel, whose multiply-adds are mostly half-rate carry
// forms (profiles/ncu_summary_r02.md) — not HBM.
#include <algorithm>
#include <chrono>

#include "msm_common.cuh"

namespace b200zk {


struct MsmConfig {
    uint32_t c, W, B;   // window bits, windows, buckets per window (2^(c-1))
    uint32_t merged;    // 1: bases are the precomputed table T[w][i] = 2^(c·w)·P_i — one bucket set for all windows
    uint32_t groups;    // bucket sets that get reduced: W (generic) or 1 (merged)
    unsigned long long table_n;  // row length of the precomputed table
};
MsmConfig msm_config(size_t n) {
    uint32_t lg = 0;
    while (((size_t)1 << (lg + 1)) <= n) ++lg;
    uint32_t c = lg > 4 ? lg - 4 : 1;
    if (c < 4) c = 4;
    if (c > 16) c = 16;
    if (lg >= 24) c = 18;
    if (lg >= 26) c = 20;
    MsmConfig m;
    m.c = c;
    m.W = (255 + c - 1) / c;
    m.B = 1u << (c - 1);
    m.merged = 0;
    m.groups = m.W;
    m.table_n = 0;
    return m;
}
MsmConfig msm_config_merged(uint32_t c, size_t table_n) {
    MsmConfig m;
    m.c = c;
    m.W = (255 + c - 1) / c;
    m.B = 1u << (c - 1);
    m.merged = 1;
    m.groups = 1;
    m.table_n = table_n;
    return m;
}
// Window bits of the precomputed tables. The column's accumulation costs n·ceil(255/c) mixed additions, its bucket
// reduction ≈ 2.3·2^(c-1) full additions whatever n is, so among the c with the same window count the smallest wins and the
// best c follows k. Measured on one GPU at S20-bn (profiles/bench_r02_table_bits.json): c=17 (15 windows, 2^16 buckets)
// 140.0 ms, c=18 140.4, c=20 (13 windows) 141.8, c=19 143.9, c=16 144.5, c=21 152.5 — k−3 up to k=20; at k=22 (S22-bn/-gl,
// four times the points per bucket set) the cap of 20 stays. On 4 and more GPUs a column's accumulation shrinks with the rank
// count but its reduction does not, which favoured small windows even more (S20-bn on 8 GPUs: c=20 47.4 ms, c=18 43.7 ms,
// c=16 44.1 ms; profiles/bench_r02_8gpu_c*.json).
uint32_t msm_table_window_bits(uint32_t k, int world) {
    uint32_t c = k < 8 ? 8 : (k > 20 ? 20 : k);
    if (c >= 12 && k <= 20) c = k - 3;
    else if (world >= 4 && c >= 12) c = std::min<uint32_t>(c, k - 2);
    if (const char* e = getenv("B200ZK_TABLE_BITS")) {  // experiments: window bits of the precomputed tables
        const int v = atoi(e);
        if (v >= 8 && v <= 22) c = (uint32_t)v;
    }
    return c;
}

constexpr int MAX_W = 64;

// Signed-digit recoding done on the fly: window w holds bits [w·c, w·c+c) plus the carry of the window below;
// values above 2^(c-1) become negative digits with a carry into the next window.
// mode 0: histogram; mode 1: scatter (counters = running offsets)
__global__ void msm_digits_kernel(const Fr* scalars, size_t n, MsmConfig cfg, uint32_t* counters, uint32_t* entries, int mode) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < n;  // no early exit: the warp votes below need every lane
    const uint32_t lane = threadIdx.x & 31, lanes_below = (1u << lane) - 1;
    Fr s = f_zero<FrCfg>();
    if (valid) s = f_from_mont(f_load(scalars + i));
    const uint32_t c = cfg.c, mask = (1u << c) - 1, halfv = 1u << (c - 1);
    uint32_t carry = 0;
    for (uint32_t w = 0; w < cfg.W; ++w) {
        const uint32_t o = w * c, limb = o >> 5, sh = o & 31;
        uint32_t v = 0;
#pragma unroll
        for (int t = 0; t < 8; ++t) {  // register-resident limb select
            if ((uint32_t)t == limb) v = s.l[t] >> sh;
        }
        if (sh + c > 32) {
#pragma unroll
            for (int t = 1; t < 8; ++t) {
                if ((uint32_t)t == limb + 1) v |= s.l[t] << (32 - sh);
            }
        }
        v = (v & mask) + carry;
        const bool neg = v > halfv;
        const uint32_t mag = neg ? (1u << c) - v : v;
        carry = neg ? 1u : 0u;
        const bool act = valid && mag != 0;
        if (!__any_sync(0xffffffffu, act)) continue;
        const uint32_t bucket = cfg.merged ? mag - 1 : w * cfg.B + mag - 1;
        // Warp-aggregated atomics: lanes that hit the same bucket elect a leader that adds the group size once. Range-check
        // witnesses put most carries of the low window into buckets 0 and 1 — without this those two counters serialise
        // ~10^6 atomics per column (0.37 ms per pass, profiles/launches_msm_batch4_r01.csv.gz).
        const uint32_t key = act ? bucket : 0xffffffe0u + lane;  // idle lanes: singleton groups (buckets stay below 2^31)
        const uint32_t peers = __match_any_sync(0xffffffffu, key);
        const uint32_t leader = __ffs(peers) - 1, cnt = __popc(peers), rank = __popc(peers & lanes_below);
        uint32_t base = 0;
        if (act && lane == leader) base = atomicAdd(counters + bucket, cnt);
        if (mode != 0) {
            base = __shfl_sync(0xffffffffu, base, leader);
            if (act) {
                const uint32_t base_index = cfg.merged ? (uint32_t)(w * cfg.table_n + i) : (uint32_t)i;
                entries[base + rank] = base_index | (neg ? 0x80000000u : 0u);
            }
        }
    }
}

// ---- u32 exclusive scan (3 phases, 4096 items per block) ----------------------------------------------------
constexpr int SCAN_THREADS = 1024, SCAN_ITEMS = 4;
__global__ void scan_block_kernel(const uint32_t* in, uint32_t* out, uint32_t* block_sums, size_t n) {
    __shared__ uint32_t warp_sums[32];
    const size_t base = ((size_t)blockIdx.x * SCAN_THREADS + threadIdx.x) * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], sum = 0;
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = base + k < n ? in[base + k] : 0;
        sum += v[k];
    }
    // inclusive warp scan of thread sums
    uint32_t x = sum;
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= (uint32_t)o) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = warp_sums[lane], ws = w;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, ws, o);
            if (lane >= (uint32_t)o) ws += y;
        }
        warp_sums[lane] = ws - w;  // exclusive
        if (lane == 31 && block_sums) block_sums[blockIdx.x] = ws;
    }
    __syncthreads();
    uint32_t run = warp_sums[wid] + x - sum;
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
}
__global__ void scan_add_kernel(uint32_t* out, const uint32_t* block_offsets, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += block_offsets[i / (SCAN_THREADS * SCAN_ITEMS)];
}
// out[i] = sum_{j<i} in[j]; out may alias in. Returns nothing; total is out[n-1] + in[n-1] (callers append a 0).
void exclusive_scan_u32(const uint32_t* in, uint32_t* out, size_t n, cudaStream_t stream) {
    if (n == 0) return;
    const size_t per = SCAN_THREADS * SCAN_ITEMS;
    const size_t nb = (n + per - 1) / per;
    if (nb == 1) {
        scan_block_kernel<<<1, SCAN_THREADS, 0, stream>>>(in, out, nullptr, n);
        ++g_launch_count;
        return;
    }
    DevBuf<uint32_t> sums(nb, stream);
    scan_block_kernel<<<(unsigned)nb, SCAN_THREADS, 0, stream>>>(in, out, sums.get(), n);
    exclusive_scan_u32(sums.get(), sums.get(), nb, stream);
    scan_add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(out, sums.get(), n);
    g_launch_count += 2;
    CUDA_CHECK(cudaGetLastError());
}

// ---- accumulate ---------------------------------------------------------------------------------------------
DEV uint32_t find_bucket(const uint32_t* offsets, uint32_t nb, uint32_t p) {
    uint32_t lo = 0, hi = nb;  // offsets[lo] <= p < offsets[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(offsets + mid) <= p) lo = mid;
        else hi = mid;
    }
    return lo;
}

#ifndef B200ZK_ACC_T_MAX
#define B200ZK_ACC_T_MAX 128
#endif
constexpr int ACC_T_MIN = 16, ACC_T_MAX = B200ZK_ACC_T_MAX;  // entries per thread, level 1 (chosen per launch below)
constexpr int COMB_T = 32;       // fan-in of the head combine levels (one warp per COMB_T partial sums)

// `entries` == nullptr: the list is `bases` itself (entry p = point p, no sign) — the output of the batched-affine rounds;
// `total_dev` != nullptr: the list length is read from the device (the host only knows an upper bound, which sizes the grid)
__global__ void __launch_bounds__(ACC_THREADS) msm_accumulate_kernel(const G1Affine* bases, const uint32_t* entries, const uint32_t* offsets,
                                                                     uint32_t nb, uint32_t total, uint32_t chunk, G1X* bucket_sums, G1X* heads,
                                                                     uint32_t* head_keys, const uint32_t* total_dev) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t p0l = (uint64_t)t * chunk;
    if (total_dev) {
        const uint32_t actual = __ldg(total_dev);
        if (p0l >= total) return;   // beyond the grid's nominal range: no head slot either
        if (p0l >= actual) {        // inside the nominal range but past the real end: an empty head slot
            head_keys[t] = INVALID_KEY;
            return;
        }
        total = actual;
    }
    if (p0l >= total) return;
    const uint32_t p0 = (uint32_t)p0l;
    const uint32_t p1 = total - p0 > chunk ? p0 + chunk : total;
    uint32_t cur = find_bucket(offsets, nb, p0);
    uint32_t end = __ldg(offsets + cur + 1);
    bool started_before = __ldg(offsets + cur) < p0;
    head_keys[t] = started_before ? cur : INVALID_KEY;
    G1X acc = g1x_identity();
    // software pipeline: the entry and the 64-byte base of the NEXT addition are requested before the current one starts —
    // the base is a random gather from the window table (mostly DRAM) and would otherwise sit in front of every addition
    uint32_t e_next = entries ? __ldg(entries + p0) : p0;
    G1Affine b_next = g1a_load_ro(bases + (e_next & 0x7fffffffu));
    for (uint32_t p = p0; p < p1; ++p) {
        const uint32_t e = e_next;
        G1Affine b = b_next;
        if (p + 1 < p1) {
            e_next = entries ? __ldg(entries + p + 1) : p + 1;
            b_next = g1a_load_ro(bases + (e_next & 0x7fffffffu));
        }
        if (p == end) {
            g1x_store(started_before ? heads + t : bucket_sums + cur, acc);
            acc = g1x_identity();
            started_before = false;
            // next non-empty bucket: binary search, not a linear walk — small-scalar columns leave tens of thousands of
            // consecutive empty buckets (a serial skip cost 19 ms per lookup-column MSM, profiles/launches_r01_before_fix.csv)
            cur = find_bucket(offsets, nb, p);
            end = __ldg(offsets + cur + 1);
        }
        if (e >> 31) b.y = f_neg(b.y);
        acc = g1x_add_affine(acc, b);
    }
    g1x_store(started_before ? heads + t : bucket_sums + cur, acc);
}

// ---- batched-affine pre-reduction ----------------------------------------------------------------------------------------
// Before the XYZZ accumulation, R rounds of pairwise AFFINE additions shrink every bucket's run of the sorted list:
// round r adds elements 2i and 2i+1 of each bucket (an odd last element is copied), so a bucket of c entries has
// ceil(c / 2^R) afterwards and (1 − 2^-R) of all additions are done at 5 products + 1 square + a share of ONE inversion
// instead of the ≈ 9.2 product-equivalents of an XYZZ mixed addition. The inversions of a round are batched over the whole
// column (Montgomery's trick, two levels): kernel A walks K outputs per thread, multiplies up the denominators x2 − x1
// (prefix products to scratch, the thread's product to `roots`), the roots are inverted by strided groups with one
// Fermat inversion per group (all lanes busy), kernel C walks the same outputs backwards, peels off each denominator's
// inverse and writes the sums. Bucket runs stay contiguous: off_dst = scan(ceil(cnt_src / 2)).
// Affine addition has no identity and no doubling: an identity operand or equal x-coordinates (P + P, P − P: impossible
// for distinct SRS points, possible for caller-supplied bases) raise `flag` and the column is redone on the XYZZ path.
// MEASURED on B200 (S20-bn, profiles/bench_r02_affine_rounds.json): bit-exact, but every round ADDS ≈ 8 ms per proof
// (accumulate phase 53.9 ms with 0 rounds, 68.1 / 76.1 / 84.2 ms with 2 / 3 / 4): a round streams the column's points
// through HBM twice more (the x-coordinates for the denominators, the points for the sums — in round 1 both are random
// 64-byte gathers from the window table) plus 96 bytes of prefix / output per pair, pays a serial 380-product inversion
// per group of roots, and crosses a bucket boundary every few outputs in the later rounds; together that costs more than
// the ≈ 3 products per addition it saves on this part. The rounds therefore stay OFF by default
// (b200zk_set_msm_affine_rounds / B200ZK_MSM_AFFINE switch them on; tests/test_gpu_msm.py keeps the path bit-exact).
static int msm_affine_rounds_env() {
    static const int rounds = [] {
        const char* e = getenv("B200ZK_MSM_AFFINE");
        const int v = e ? atoi(e) : -1;
        return v > 6 ? 6 : v;
    }();
    return rounds;
}
constexpr int BA_K = 16;        // outputs per thread
constexpr int BA_THREADS = 128;
__global__ void ba_counts_kernel(const uint32_t* off_src, uint32_t nb, uint32_t* cnt_dst) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nb) return;
    cnt_dst[b] = b == nb ? 0u : (off_src[b + 1] - off_src[b] + 1u) >> 1;
}
struct BaSrc {
    const G1Affine* pts;       // round >= 2: the previous round's points; round 1: the (window-table) bases
    const uint32_t* entries;   // round 1: sorted entries (index | sign << 31); nullptr afterwards
};
DEV G1Affine ba_fetch(const BaSrc& S, uint32_t s) {
    if (!S.entries) return g1a_load_ro(S.pts + s);
    const uint32_t e = __ldg(S.entries + s);
    G1Affine p = g1a_load_ro(S.pts + (e & 0x7fffffffu));
    if (e >> 31) p.y = f_neg(p.y);
    return p;
}
DEV Fq ba_fetch_x(const BaSrc& S, uint32_t s) {
    if (!S.entries) return f_load_ro(&S.pts[s].x);
    return f_load_ro(&S.pts[__ldg(S.entries + s) & 0x7fffffffu].x);
}
// kernel A: prefix products of the denominators of this thread's outputs
__global__ void __launch_bounds__(BA_THREADS) ba_prefix_kernel(BaSrc S, const uint32_t* off_src, const uint32_t* off_dst, uint32_t nb, Fq* pref, Fq* roots,
                                                               uint32_t* flag) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t L = __ldg(off_dst + nb);
    const uint64_t j0l = (uint64_t)t * BA_K;
    if (j0l >= L) return;
    const uint32_t j0 = (uint32_t)j0l, j1 = L - j0 > (uint32_t)BA_K ? j0 + BA_K : L;
    uint32_t b = find_bucket(off_dst, nb, j0), end = __ldg(off_dst + b + 1), dbase = __ldg(off_dst + b), sbase = __ldg(off_src + b),
             scnt = __ldg(off_src + b + 1) - sbase;
    Fq run = f_one<FqCfg>();
    uint32_t bad = 0;
    for (uint32_t j = j0; j < j1; ++j) {
        if (j == end) {
            b = find_bucket(off_dst, nb, j);
            end = __ldg(off_dst + b + 1);
            dbase = __ldg(off_dst + b);
            sbase = __ldg(off_src + b);
            scnt = __ldg(off_src + b + 1) - sbase;
        }
        const uint32_t i2 = 2 * (j - dbase);
        f_store(pref + j, run);
        if (i2 + 1 < scnt) {
            const Fq x1 = ba_fetch_x(S, sbase + i2), x2 = ba_fetch_x(S, sbase + i2 + 1);
            const Fq dx = f_sub(x2, x1);
            if (f_is_zero(dx)) bad = 1;          // P + P or P − P (or two identities)
            else run = f_mul(run, dx);
        }
    }
    f_store(roots + t, run);
    if (bad) atomicOr(flag, 1u);
}
// strided batch inversion of the roots (nonzero by construction): thread g owns roots g, g + G, ...
__global__ void __launch_bounds__(128) ba_invert_roots_kernel(Fq* roots, Fq* scratch, const uint32_t* off_dst, uint32_t nb, uint32_t G) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t L = __ldg(off_dst + nb), n = (L + BA_K - 1) / BA_K;
    if (g >= G || g >= n) return;
    Fq acc = f_one<FqCfg>();
    for (uint32_t i = g; i < n; i += G) {
        f_store(scratch + i, acc);
        acc = f_mul(acc, f_load(roots + i));
    }
    acc = f_inv(acc);
    const uint32_t cnt = (n - g + G - 1) / G;
    for (uint32_t q = cnt; q-- > 0;) {
        const uint32_t i = g + q * G;
        const Fq v = f_load(roots + i);
        f_store(roots + i, f_mul(acc, f_load(scratch + i)));
        acc = f_mul(acc, v);
    }
}
// kernel C: the sums (outputs walked backwards: inverse of denominator j = inv_run · pref[j], inv_run ·= dx_j)
__global__ void __launch_bounds__(BA_THREADS) ba_add_kernel(BaSrc S, const uint32_t* off_src, const uint32_t* off_dst, uint32_t nb, const Fq* pref,
                                                            const Fq* roots_inv, G1Affine* dst, uint32_t* flag) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t L = __ldg(off_dst + nb);
    const uint64_t j0l = (uint64_t)t * BA_K;
    if (j0l >= L) return;
    const uint32_t j0 = (uint32_t)j0l, j1 = L - j0 > (uint32_t)BA_K ? j0 + BA_K : L;
    uint32_t b = find_bucket(off_dst, nb, j1 - 1), dbase = __ldg(off_dst + b), sbase = __ldg(off_src + b), scnt = __ldg(off_src + b + 1) - sbase;
    Fq inv_run = f_load(roots_inv + t);
    uint32_t bad = 0;
    for (uint32_t j = j1; j-- > j0;) {
        if (j < dbase) {
            b = find_bucket(off_dst, nb, j);
            dbase = __ldg(off_dst + b);
            sbase = __ldg(off_src + b);
            scnt = __ldg(off_src + b + 1) - sbase;
        }
        const uint32_t i2 = 2 * (j - dbase);
        const G1Affine p1 = ba_fetch(S, sbase + i2);
        if (g1_is_identity(p1)) bad = 1;
        G1Affine o = p1;
        if (i2 + 1 < scnt) {
            const G1Affine p2 = ba_fetch(S, sbase + i2 + 1);
            if (g1_is_identity(p2)) bad = 1;
            const Fq dx = f_sub(p2.x, p1.x);
            if (!f_is_zero(dx)) {
                const Fq inv = f_mul(inv_run, f_load(pref + j));
                inv_run = f_mul(inv_run, dx);
                const Fq lam = f_mul(f_sub(p2.y, p1.y), inv);
                o.x = f_sub(f_sub(f_sqr(lam), p1.x), p2.x);
                o.y = f_sub(f_mul(lam, f_sub(p1.x, o.x)), p1.y);
            }
        }
        f_store(&dst[j].x, o.x);
        f_store(&dst[j].y, o.y);
    }
    if (bad) atomicOr(flag, 1u);
}

// Σ 2^(c·w)·S_w on the host (64-bit limb path of the shared field code)
G1X msm_fold_windows(const G1X* window_sums, uint32_t W, uint32_t c) {
    G1X acc = g1x_identity();
    for (uint32_t w = W; w-- > 0;) {
        for (uint32_t i = 0; i < c; ++i) acc = g1x_dbl(acc);
        acc = g1x_add(acc, window_sums[w]);
    }
    return acc;
}

// Phase A of one MSM column (digits -> counting sort -> chunked accumulation -> head combine), cut in two halves around
// the one host read-back (the entry count) so that TWO columns can be in flight on two streams: while one column's
// accumulate saturates the multiplier pipe, the other's latency-bound pieces (histogram atomics, scans, the read-back,
// the short combine levels) fill the gaps. All scratch is preallocated at its worst-case size per slot.
struct MsmSlot {
    cudaStream_t st = nullptr;
    cudaEvent_t ev = nullptr;
    uint32_t* total_host = nullptr;  // pinned
    bool empty = false;              // the column in flight has no points on this rank
    DevBuf<uint32_t> counters, offsets, entries, keys_a, keys_b;
    DevBuf<G1X> heads_a, heads_b;
    // batched-affine rounds (optional): ping-pong offsets and points, prefix products, per-thread roots, exception flag
    int affine_rounds = 0;
    DevBuf<uint32_t> ba_off[2];
    DevBuf<G1Affine> ba_pts[2];
    DevBuf<Fq> ba_pref, ba_roots, ba_rscratch;
    void alloc(Context& ctx, uint32_t nb, size_t max_entries, int rounds) {
        affine_rounds = rounds;
        if (rounds > 0) {
            const size_t l2 = (max_entries + nb) / 2 + 2, l3 = (l2 + nb) / 2 + 2;
            for (auto& o : ba_off) o.alloc(nb + 1, ctx.stream);
            ba_pts[0].alloc(l2, ctx.stream);
            if (rounds > 1) ba_pts[1].alloc(l3, ctx.stream);
            ba_pref.alloc(l2, ctx.stream);
            ba_roots.alloc(l2 / BA_K + 2, ctx.stream);
            ba_rscratch.alloc(l2 / BA_K + 2, ctx.stream);
        }
        const size_t t1 = (max_entries + ACC_T_MIN - 1) / ACC_T_MIN + 1, t2 = (t1 + COMB_T - 1) / COMB_T + 1;
        counters.alloc(nb + 1, ctx.stream);
        offsets.alloc(nb + 1, ctx.stream);
        entries.alloc(max_entries + 1, ctx.stream);
        heads_a.alloc(t1, ctx.stream);
        keys_a.alloc(t1, ctx.stream);
        heads_b.alloc(t2, ctx.stream);
        keys_b.alloc(t2, ctx.stream);
    }
};
static void msm_issue_count(MsmSlot& sl, const Fr* scalars, size_t n, const MsmConfig& cfg) {
    cudaStream_t s = sl.st;
    sl.empty = n == 0;
    if (sl.empty) return;
    const uint32_t nb = cfg.groups * cfg.B;
    CUDA_CHECK(cudaMemsetAsync(sl.counters.get(), 0, (nb + 1) * 4, s));
    msm_digits_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(scalars, n, cfg, sl.counters.get(), nullptr, 0);
    ++g_launch_count;
    exclusive_scan_u32(sl.counters.get(), sl.offsets.get(), nb + 1, s);
    CUDA_CHECK(cudaMemcpyAsync(sl.total_host, sl.offsets.get() + nb, 4, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaMemcpyAsync(sl.counters.get(), sl.offsets.get(), (nb + 1) * 4, cudaMemcpyDeviceToDevice, s));
    CUDA_CHECK(cudaEventRecord(sl.ev, s));
}
// `flag_dev`: where the affine rounds report an exceptional pair for this column (nullptr: XYZZ path only)
static void msm_issue_accumulate(MsmSlot& sl, const G1Affine* bases, const Fr* scalars, size_t n, const MsmConfig& cfg, G1X* bucket_sums,
                                 uint32_t* flag_dev) {
    cudaStream_t s = sl.st;
    const uint32_t nb = cfg.groups * cfg.B;
    if (sl.empty) return;
    CUDA_CHECK(cudaEventSynchronize(sl.ev));
    const uint32_t total = *sl.total_host;
    if (total == 0) return;
    msm_digits_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(scalars, n, cfg, sl.counters.get(), sl.entries.get(), 1);
    ++g_launch_count;
    // dense columns (>= 8 entries per bucket on average) first go through the batched-affine rounds
    const G1Affine* acc_bases = bases;
    const uint32_t* acc_entries = sl.entries.get();
    const uint32_t* acc_offsets = sl.offsets.get();
    const uint32_t* acc_total_dev = nullptr;
    uint32_t acc_total = total;
    const bool affine = flag_dev != nullptr && sl.affine_rounds > 0 && (uint64_t)total >= (uint64_t)8 * nb;
    const int prof_h = prof_begin(PROF_MSM_ACCUMULATE, s, (double)total);
    if (affine) {
        BaSrc src{bases, sl.entries.get()};
        const uint32_t* off_src = sl.offsets.get();
        uint64_t lub = total;
        for (int r = 0; r < sl.affine_rounds; ++r) {
            uint32_t* off_dst = sl.ba_off[r & 1].get();
            G1Affine* dst = sl.ba_pts[r & 1].get();
            lub = (lub + nb) / 2 + 1;  // upper bound of the round's output length (the exact one stays on the device)
            ba_counts_kernel<<<(nb + 1 + 255) / 256, 256, 0, s>>>(off_src, nb, sl.counters.get());
            ++g_launch_count;
            exclusive_scan_u32(sl.counters.get(), off_dst, nb + 1, s);
            const uint32_t nt = (uint32_t)((lub + BA_K - 1) / BA_K), G = std::max<uint32_t>(1, nt / 64);
            ba_prefix_kernel<<<(nt + BA_THREADS - 1) / BA_THREADS, BA_THREADS, 0, s>>>(src, off_src, off_dst, nb, sl.ba_pref.get(), sl.ba_roots.get(), flag_dev);
            ba_invert_roots_kernel<<<(G + 127) / 128, 128, 0, s>>>(sl.ba_roots.get(), sl.ba_rscratch.get(), off_dst, nb, G);
            ba_add_kernel<<<(nt + BA_THREADS - 1) / BA_THREADS, BA_THREADS, 0, s>>>(src, off_src, off_dst, nb, sl.ba_pref.get(), sl.ba_roots.get(), dst, flag_dev);
            g_launch_count += 3;
            CUDA_CHECK(cudaGetLastError());
            src = BaSrc{dst, nullptr};
            off_src = off_dst;
        }
        acc_bases = src.pts;
        acc_entries = nullptr;
        acc_offsets = off_src;
        acc_total_dev = off_src + nb;
        acc_total = (uint32_t)lub;
    }
    // chunk length: long chunks halve the head list for big columns, short ones keep small columns (1–2 M entries) wide
    // enough to fill 148 SMs
    uint32_t chunk = ACC_T_MAX;
    while (chunk > (uint32_t)ACC_T_MIN && (uint64_t)acc_total / chunk < (uint64_t)148 * 16 * 32) chunk >>= 1;
    const uint32_t nthreads = (uint32_t)(((uint64_t)acc_total + chunk - 1) / chunk);
    msm_accumulate_kernel<<<(nthreads + ACC_THREADS - 1) / ACC_THREADS, ACC_THREADS, 0, s>>>(acc_bases, acc_entries, acc_offsets, nb, acc_total, chunk, bucket_sums,
                                                                                              sl.heads_a.get(), sl.keys_a.get(), acc_total_dev);
    prof_end(prof_h, s);
    ++g_launch_count;
    CUDA_CHECK(cudaGetLastError());
    // levels >= 2: segment-sum the head list until a single warp covers it
    uint32_t len = nthreads;
    bool flip = false;
    while (true) {
        const uint32_t nt = (len + COMB_T - 1) / COMB_T;
        G1X* in_p = flip ? sl.heads_b.get() : sl.heads_a.get();
        uint32_t* in_k = flip ? sl.keys_b.get() : sl.keys_a.get();
        G1X* out_p = flip ? sl.heads_a.get() : sl.heads_b.get();
        uint32_t* out_k = flip ? sl.keys_a.get() : sl.keys_b.get();
        msm_launch_combine(in_p, in_k, len, bucket_sums, out_p, out_k, s);
        if (nt == 1) break;  // a single warp has no predecessor: nothing can be left in its head slot
        len = nt;
        flip = !flip;
    }
}

// optional cross-rank combine of partial window sums (set through b200zk_set_allgather)
// point range of this rank for an n-point MSM (contiguous shards; the last rank takes the remainder)
static void shard_range(const Context& ctx, size_t n, size_t& lo, size_t& len) {
    if (!ctx.sharded()) {
        lo = 0;
        len = n;
        return;
    }
    const size_t per = n / ctx.world;
    lo = per * ctx.rank;
    len = ctx.rank == ctx.world - 1 ? n - lo : per;
}

static void check_cfg(const MsmConfig& cfg, size_t n) {
    if (cfg.W > (uint32_t)MAX_W) throw std::runtime_error("msm: too many windows");
    if (n >= ((size_t)1 << 31) || (cfg.merged && cfg.table_n * cfg.W >= ((size_t)1 << 31))) throw std::invalid_argument("msm: index space must be < 2^31");
}
// `ncols` MSMs with one configuration: per-column bucket accumulation, ONE bucket reduction for all columns. Column j reads
// the bases at col_bases[j] (the two SRS bases can be mixed in a batch); with col_partial[j] set, only this rank's point
// range of column j is accumulated. sums_out gets cfg.groups entries per column: the (partial) window sums, XYZZ.
// `gather_ranks` > 1 (one reduction round only): sums_out holds every rank's sums, rank-major (see msm_reduce_groups).
static void msm_batch_core(Context& ctx, const G1Affine* const* col_bases, const Fr* const* cols, const uint8_t* col_partial, size_t ncols, size_t n,
                           const MsmConfig& cfg, std::vector<G1X>& sums_out, int gather_ranks = 1) {
    check_cfg(cfg, n);
    cudaStream_t s = ctx.stream;
    size_t shard_lo = 0, shard_len = n;
    shard_range(ctx, n, shard_lo, shard_len);
    const uint32_t nb = cfg.groups * cfg.B;
    sums_out.clear();
    // bound the scratch: at most 32 columns (≈2 GiB of bucket sums at c = 20) per reduction round
    const size_t round = 32;
    if (gather_ranks > 1 && ncols > round) throw std::logic_error("msm: device-side gather needs a single reduction round");
    for (size_t c0 = 0; c0 < ncols; c0 += round) {
        const size_t nc = std::min(round, ncols - c0);
        DevBuf<G1X> bucket_sums((size_t)nc * nb, s);
        CUDA_CHECK(cudaMemsetAsync(bucket_sums.get(), 0, (size_t)nc * nb * sizeof(G1X), s));
        auto lo_of = [&](size_t j) { return col_partial && col_partial[j] ? shard_lo : (size_t)0; };
        auto len_of = [&](size_t j) { return col_partial && col_partial[j] ? shard_len : n; };
        size_t max_len = 0;
        for (size_t j = 0; j < nc; ++j) max_len = std::max(max_len, len_of(c0 + j));
        if (max_len > 0) {
            // several columns in flight: slot 0 on the context stream, the others on auxiliary streams
            // (under the event profiler one column at a time: concurrent streams would inflate each other's bracketed durations)
            const int nslots = ctx.aux_streams[0] && !g_prof_enabled ? (int)std::min<size_t>(nc, MSM_SLOTS) : 1;
            MsmSlot slots[MSM_SLOTS];
            // declared after the slots, so it runs before their buffers go back to the arena: if anything below throws while
            // auxiliary streams still have kernels queued on those buffers, drain every stream first
            struct DrainOnUnwind {
                Context& c;
                int nslots;
                bool armed = true;
                ~DrainOnUnwind() {
                    if (!armed) return;
                    for (int q = 1; q < nslots; ++q) cudaStreamSynchronize(c.aux_streams[q - 1]);
                    cudaStreamSynchronize(c.stream);
                }
            } drain{ctx, nslots};
            // batched-affine rounds for dense columns (merged-table configurations; B200ZK_MSM_AFFINE=0 switches them off)
            const int affine_rounds = cfg.merged ? (msm_affine_rounds_env() >= 0 ? msm_affine_rounds_env() : ctx.msm_affine_rounds) : 0;
            DevBuf<uint32_t> col_flags(affine_rounds ? nc : 0, s);
            if (affine_rounds) CUDA_CHECK(cudaMemsetAsync(col_flags.get(), 0, nc * 4, s));
            for (int q = 0; q < nslots; ++q) {
                slots[q].st = q == 0 ? s : ctx.aux_streams[q - 1];
                slots[q].ev = ctx.msm_events[q];
                slots[q].total_host = ctx.pinned_u32 + q;
                slots[q].alloc(ctx, nb, max_len * cfg.W, affine_rounds);
            }
            if (nslots > 1) {  // the aux streams may touch the buffers only after everything queued so far on the main stream
                CUDA_CHECK(cudaEventRecord(ctx.msm_fork, s));
                for (int q = 1; q < nslots; ++q) CUDA_CHECK(cudaStreamWaitEvent(ctx.aux_streams[q - 1], ctx.msm_fork, 0));
            }
            auto count = [&](MsmSlot& sl, size_t j) {
                if (ctx.column_gate) ctx.column_gate(j, sl.st);  // the column may still be on its way to the device
                msm_issue_count(sl, cols[j] + lo_of(j), len_of(j), cfg);
            };
            for (size_t j = 0; j < (size_t)nslots && j < nc; ++j) count(slots[j], c0 + j);
            for (size_t j = 0; j < nc; ++j) {
                MsmSlot& sl = slots[j % nslots];
                const size_t col = c0 + j;
                msm_issue_accumulate(sl, col_bases[col] + lo_of(col), cols[col] + lo_of(col), len_of(col), cfg, bucket_sums.get() + j * nb,
                                     affine_rounds ? col_flags.get() + j : nullptr);
                if (j + nslots < nc) count(sl, col + nslots);
            }
            for (int q = 1; q < nslots; ++q) {  // main stream (reduce, frees) continues after the aux streams have drained
                CUDA_CHECK(cudaEventRecord(ctx.msm_join[q - 1], ctx.aux_streams[q - 1]));
                CUDA_CHECK(cudaStreamWaitEvent(s, ctx.msm_join[q - 1], 0));
            }
            CUDA_CHECK(cudaStreamSynchronize(s));  // slot buffers are released below: nothing may still be using them
            if (affine_rounds) {
                // a column whose affine rounds met an identity operand or equal x-coordinates is redone on the XYZZ path
                std::vector<uint32_t> flags(nc);
                CUDA_CHECK(cudaMemcpy(flags.data(), col_flags.get(), nc * 4, cudaMemcpyDeviceToHost));
                for (size_t j = 0; j < nc; ++j) {
                    if (!flags[j]) continue;
                    const size_t col = c0 + j;
                    CUDA_CHECK(cudaMemsetAsync(bucket_sums.get() + j * nb, 0, (size_t)nb * sizeof(G1X), s));
                    count(slots[0], col);
                    msm_issue_accumulate(slots[0], col_bases[col] + lo_of(col), cols[col] + lo_of(col), len_of(col), cfg, bucket_sums.get() + j * nb, nullptr);
                    CUDA_CHECK(cudaStreamSynchronize(s));
                }
            }
            drain.armed = false;
        }
        std::vector<G1X> ws;
        msm_reduce_groups(ctx, bucket_sums.get(), (uint32_t)(nc * cfg.groups), cfg.B, ws, gather_ranks);
        sums_out.insert(sums_out.end(), ws.begin(), ws.end());
    }
}
// window sums of one column -> canonical affine result
static G1Affine msm_finish(const MsmConfig& cfg, const G1X* sums) {
    return cfg.merged ? g1x_to_affine(sums[0]) : g1x_to_affine(msm_fold_windows(sums, cfg.W, cfg.c));
}

// Multi-GPU split of a batch (both levels of the north star). With `world` ranks and ncols = q·world + rem:
// the first q·world columns are dealt out by COLUMN — rank r commits columns r, r+world, ... over the full point range —
// and the remaining rem (< world) columns are split by POINT RANGE, every rank accumulating its contiguous shard. All of
// a rank's work goes through ONE accumulate batch and one bucket reduction; one all-gather (ncclAllGather out of the
// reduction's output buffer once the library's communicator is up, the host callback otherwise) then carries
// the q finished sums and the rem partial window sums of every rank, and the partial ones are added on the host.
static void msm_batch_distribute(Context& ctx, const G1Affine* const* col_bases, const Fr* const* cols, size_t ncols, size_t n, const MsmConfig& cfg,
                                 G1Affine* out) {
    const bool dist = ctx.sharded();
    const size_t G = cfg.groups;
    std::vector<G1X> sums;
    if (!dist) {
        msm_batch_core(ctx, col_bases, cols, nullptr, ncols, n, cfg, sums);
        for (size_t j = 0; j < ncols; ++j) out[j] = msm_finish(cfg, sums.data() + j * G);
        return;
    }
    const size_t world = ctx.world, q = ncols / world, rem = ncols % world, dealt = q * world;
    std::vector<const G1Affine*> my_bases;
    std::vector<const Fr*> my_cols;
    std::vector<uint8_t> partial;
    for (size_t j = ctx.rank; j < dealt; j += world) my_bases.push_back(col_bases[j]), my_cols.push_back(cols[j]), partial.push_back(0);
    for (size_t j = dealt; j < ncols; ++j) my_bases.push_back(col_bases[j]), my_cols.push_back(cols[j]), partial.push_back(1);
    const size_t per = (q + rem) * G;  // entries per rank, the same on every rank
    std::vector<G1X> all;
    if (ctx.nccl && my_cols.size() <= 32) {
        // the library's own communicator is up (create_proof brings it up): the partial sums go from the reduction's output
        // buffer through one ncclAllGather and come back to the host in one copy
        msm_batch_core(ctx, my_bases.data(), my_cols.data(), partial.data(), my_cols.size(), n, cfg, all, (int)world);
    } else {
        msm_batch_core(ctx, my_bases.data(), my_cols.data(), partial.data(), my_cols.size(), n, cfg, sums);
        all.resize(per * world);
        const auto t0 = std::chrono::steady_clock::now();
        if (ctx.nccl) {
            Sharder(ctx).host_allgather(sums.data(), per * sizeof(G1X), all.data());
        } else {  // no communicator: the host's all-gather callback carries the sums (hosts without NCCL, single-GPU replays)
            if (!ctx.allgather || ctx.allgather(ctx.allgather_user, sums.data(), per * sizeof(G1X), all.data()) != 0)
                throw std::runtime_error("msm: all-gather callback failed");
        }
        ctx.exchange_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    for (size_t j = 0; j < dealt; ++j) out[j] = msm_finish(cfg, all.data() + (j % world) * per + (j / world) * G);
    std::vector<G1X> acc(G);
    for (size_t j = dealt; j < ncols; ++j) {
        for (size_t g = 0; g < G; ++g) {
            acc[g] = g1x_identity();
            for (size_t r = 0; r < world; ++r) acc[g] = g1x_add(acc[g], all[r * per + (q + (j - dealt)) * G + g]);
        }
        out[j] = msm_finish(cfg, acc.data());
    }
}

// arbitrary bases (best_multiexp): windows kept separate, folded on the host
G1Affine msm_run(Context& ctx, const G1Affine* bases, const Fr* scalars, size_t n) {
    G1Affine r;
    msm_batch_distribute(ctx, &bases, &scalars, 1, n, msm_config(n), &r);
    return r;
}

// SRS bases: the precomputed window tables when they exist (one bucket set, no fold). `basis[j]` selects g (0) or
// g_lagrange (1) for column j.
void msm_batch_srs_mixed(Context& ctx, const int* basis, const Fr* const* cols, size_t ncols, size_t n, G1Affine* out) {
    if (ncols == 0) return;
    const Srs& srs = *ctx.srs;
    bool uniform = true;
    for (size_t j = 1; j < ncols; ++j) uniform = uniform && basis[j] == basis[0];
    auto table = [&](int b) -> const DevBuf<G1Affine>& { return b == 0 ? srs.g_tab : srs.gl_tab; };
    if (ctx.sharded() && ctx.nccl && srs.tab_agreed[0] < 0) {  // one exchange per SRS: use a table only if every rank has it
        uint32_t mine[2] = {table(0).size() != 0, table(1).size() != 0};
        std::vector<uint32_t> all(2 * (size_t)ctx.world);
        Sharder(ctx).host_allgather(mine, sizeof(mine), all.data());
        for (int b = 0; b < 2; ++b) {
            srs.tab_agreed[b] = 1;
            for (int r = 0; r < ctx.world; ++r) srs.tab_agreed[b] &= (int)all[2 * r + b];
        }
    }
    auto has_table = [&](int b) { return table(b).size() != 0 && srs.tab_agreed[b] != 0 && n * 8 >= srs.n; };
    if (!uniform && !(has_table(0) && has_table(1))) {  // the two bases would use different configurations: one call per run
        for (size_t j0 = 0; j0 < ncols;) {
            size_t j1 = j0 + 1;
            while (j1 < ncols && basis[j1] == basis[j0]) ++j1;
            msm_batch_srs_mixed(ctx, basis + j0, cols + j0, j1 - j0, n, out + j0);
            j0 = j1;
        }
        return;
    }
    const bool tabled = has_table(basis[0]);
    std::vector<const G1Affine*> col_bases(ncols);
    for (size_t j = 0; j < ncols; ++j)
        col_bases[j] = tabled ? table(basis[j]).get() : (basis[j] == 0 ? srs.g.get() : srs.g_lagrange.get());
    msm_batch_distribute(ctx, col_bases.data(), cols, ncols, n, tabled ? msm_config_merged(srs.tab_c, srs.n) : msm_config(n), out);
}
void msm_batch_srs(Context& ctx, int basis, const Fr* const* cols, size_t ncols, size_t n, G1Affine* out) {
    std::vector<int> b(ncols, basis);
    msm_batch_srs_mixed(ctx, b.data(), cols, ncols, n, out);
}
G1Affine msm_run_srs(Context& ctx, int basis, const Fr* scalars, size_t n) {
    G1Affine r;
    msm_batch_srs(ctx, basis, &scalars, 1, n, &r);
    return r;
}

// T[w][i] = 2^(c·w)·P_i as affine points, w < W: built once per SRS
// Each thread takes TABLE_PTS points (a stride of n / TABLE_PTS apart, so a warp's accesses stay coalesced) and shares ONE
// field inversion among them (Montgomery's trick): c doublings ≈ 160 products per point, against 380 for a Fermat inversion
// per point — the table build (13 rows × 2 bases at k=20) halves.
constexpr int TABLE_PTS = 4;
__global__ void __launch_bounds__(128) msm_table_step_kernel(const G1Affine* prev, G1Affine* next, size_t n, uint32_t c) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x, per = (n + TABLE_PTS - 1) / TABLE_PTS;
    if (t >= per) return;
    G1X a[TABLE_PTS];
    Fq d[TABLE_PTS], pref[TABLE_PTS];
    Fq run = f_one<FqCfg>();
#pragma unroll
    for (int j = 0; j < TABLE_PTS; ++j) {
        const size_t i = t + (size_t)j * per;
        a[j] = g1x_identity();
        if (i < n) {
            G1Affine p;
            p.x = f_load(&prev[i].x);
            p.y = f_load(&prev[i].y);
            a[j] = g1x_from_affine(p);
            for (uint32_t q = 0; q < c; ++q) a[j] = g1x_dbl(a[j]);
        }
        d[j] = g1_is_identity(a[j]) ? f_one<FqCfg>() : f_mul(a[j].zz, a[j].zzz);  // identity: a harmless factor of one
        pref[j] = run;
        run = f_mul(run, d[j]);
    }
    Fq inv = f_inv(run);
#pragma unroll
    for (int j = TABLE_PTS - 1; j >= 0; --j) {
        const size_t i = t + (size_t)j * per;
        const Fq dinv = f_mul(inv, pref[j]);  // 1 / (zz·zzz) of point j
        inv = f_mul(inv, d[j]);
        if (i >= n) continue;
        G1Affine r;
        if (g1_is_identity(a[j])) {
            r.x = f_zero<FqCfg>();
            r.y = f_zero<FqCfg>();
        } else {
            r.x = f_mul(a[j].x, f_mul(dinv, a[j].zzz));  // X / ZZ
            r.y = f_mul(a[j].y, f_mul(dinv, a[j].zz));   // Y / ZZZ
        }
        f_store(&next[i].x, r.x);
        f_store(&next[i].y, r.y);
    }
}
void msm_build_table(Context& ctx, const G1Affine* bases, size_t n, uint32_t c, DevBuf<G1Affine>& table) {
    const uint32_t W = (255 + c - 1) / c;
    table.alloc_persistent((size_t)W * n, ctx.stream);
    CUDA_CHECK(cudaMemcpyAsync(table.get(), bases, n * sizeof(G1Affine), cudaMemcpyDeviceToDevice, ctx.stream));
    for (uint32_t w = 1; w < W; ++w) {
        const size_t per = (n + TABLE_PTS - 1) / TABLE_PTS;
        msm_table_step_kernel<<<(unsigned)((per + 127) / 128), 128, 0, ctx.stream>>>(table.get() + (size_t)(w - 1) * n, table.get() + (size_t)w * n, n, c);
        ++g_launch_count;
    }
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace b200zk
