// Device-resident mirror of halo2_proofs::plonk::{keygen_vk, keygen_pk, create_proof} with the KZG / SHPLONK
// backend for the ConstraintSystem of halo2-base's BaseConfig (SURVEY.md §3.2, Appendix B, D.3–D.12). This is the
// path the reference drives through `base_test().k(k).bench_builder(..)` at verifier/src/stark/mod.rs:543 and :593.
//
// Every column stays in HBM between steps; only commitments (64 B), evaluations (32 B) and challenges cross the
// PCIe bus after the witness upload. The transcript and the Fr::random stream are host-side (row L), exactly as in
// the reference; everything else is a launch of the kernels in ntt.cu / msm.cu / poly.cu / quotient.cu.
#include "prover.cuh"
#include "upload.cuh"

#include "collectives.cuh"

#include <algorithm>
#include <condition_variable>
#include <chrono>

namespace b200zk {

G1Affine msm_run_srs(Context& ctx, int basis, const Fr* scalars, size_t n);
void msm_batch_srs(Context& ctx, int basis, const Fr* const* cols, size_t ncols, size_t n, G1Affine* out);
void msm_batch_srs_mixed(Context& ctx, const int* basis, const Fr* const* cols, size_t ncols, size_t n, G1Affine* out);

static const Srs& need_srs(Context& ctx, uint32_t k) {
    if (!ctx.srs) throw std::runtime_error("no SRS loaded");
    if (ctx.srs->k != k) throw std::runtime_error("SRS size does not match the circuit (k)");
    return *ctx.srs;
}
static G1Affine commit_coeff(Context& ctx, const Fr* coeffs, size_t n) { return msm_run_srs(ctx, 0, coeffs, n); }
// commitments of `count` columns `stride` apart: bucket accumulation per column, one bucket reduction for the batch
static std::vector<G1Affine> commit_batch(Context& ctx, int basis, const Fr* first, size_t stride, size_t count, size_t n) {
    std::vector<const Fr*> cols(count);
    for (size_t i = 0; i < count; ++i) cols[i] = first + i * stride;
    std::vector<G1Affine> out(count);
    if (count) msm_batch_srs(ctx, basis, cols.data(), count, n, out.data());
    return out;
}

// ---- permutation::keygen::Assembly (host, serial — as upstream) ------------------------------------------------------
struct Assembly {
    size_t n;
    uint32_t p;
    std::vector<uint32_t> map_col, map_row, aux_col, aux_row, sizes;  // [p][n] flattened
    Assembly(size_t n_, uint32_t p_) : n(n_), p(p_), map_col(n_ * p_), map_row(n_ * p_), aux_col(n_ * p_), aux_row(n_ * p_), sizes(n_ * p_, 1) {
        for (uint32_t c = 0; c < p; ++c)
            for (size_t r = 0; r < n; ++r) {
                map_col[c * n + r] = aux_col[c * n + r] = c;
                map_row[c * n + r] = aux_row[c * n + r] = (uint32_t)r;
            }
    }
    void copy(uint32_t lc, uint32_t lr, uint32_t rc, uint32_t rr) {
        if (lc >= p || rc >= p || lr >= n || rr >= n) throw std::invalid_argument("copy constraint out of range");
        size_t li = lc * n + lr, ri = rc * n + rr;
        uint32_t lcc = aux_col[li], lcr = aux_row[li], rcc = aux_col[ri], rcr = aux_row[ri];
        if (lcc == rcc && lcr == rcr) return;
        if (sizes[lcc * n + lcr] < sizes[rcc * n + rcr]) {
            std::swap(lcc, rcc);
            std::swap(lcr, rcr);
        }
        sizes[lcc * n + lcr] += sizes[rcc * n + rcr];
        uint32_t ic = rcc, ir = rcr;
        for (;;) {
            size_t ii = ic * n + ir;
            aux_col[ii] = lcc;
            aux_row[ii] = lcr;
            uint32_t nc = map_col[ii], nr = map_row[ii];
            ic = nc;
            ir = nr;
            if (ic == rcc && ir == rcr) break;
        }
        std::swap(map_col[li], map_col[ri]);
        std::swap(map_row[li], map_row[ri]);
    }
};

// [UNVERIFIED-5 of SURVEY §8c] stand-in for upstream's Debug-string hash; same bytes as the oracle's rendering.
static Fr default_transcript_repr(const ProvingKeyDev& pk) {
    host::Blake2b512 h("Halo2-Verify-Key");
    const Shape& sh = pk.shape;
    std::string s = "b200zk-vk k=" + std::to_string(sh.k) + " A=" + std::to_string(sh.A) + " L=" + std::to_string(sh.L) + " F=" + std::to_string(sh.F);
    uint64_t len = s.size() + 64 * (pk.fixed_commitments.size() + pk.perm_commitments.size());
    h.absorb(&len, 8);
    h.absorb(s.data(), s.size());
    auto absorb = [&](const G1Affine& p) {
        uint8_t b[64];
        host::field_to_bytes(p.x, b);
        host::field_to_bytes(p.y, b + 32);
        h.absorb(b, 64);
    };
    for (auto& p : pk.fixed_commitments) absorb(p);
    for (auto& p : pk.perm_commitments) absorb(p);
    uint8_t d[64];
    h.peek(d);
    uint32_t w[16];
    memcpy(w, d, 64);
    return f_from_u512<FrCfg>(w);
}

std::unique_ptr<ProvingKeyDev> keygen(Context& ctx, const Shape& sh, const Fr* fixed_host, const uint32_t* copies, size_t ncopies) {
    if (sh.k < 4 || sh.k + 2 > FrConsts::S || sh.A == 0) throw std::invalid_argument("keygen: unsupported shape");
    if (sh.num_advice() > Q_MAX_ADVICE || sh.num_fixed() > Q_MAX_FIXED || sh.num_perm() > Q_MAX_PERM || sh.num_sets() > Q_MAX_SETS)
        throw std::invalid_argument("keygen: too many columns for the h(X) kernel argument block");
    need_srs(ctx, sh.k);
    cudaStream_t s = ctx.stream;
    const size_t n = sh.n(), en = 4 * n;
    const Domain& dom = ctx.domain(sh.k);
    const TwiddleTable& tw = ctx.std_table(sh.k + 2);
    auto pk = std::make_unique<ProvingKeyDev>();
    pk->shape = sh;
    const uint32_t NF = sh.num_fixed(), P = sh.num_perm();
    // fixed columns
    pk->fixed_values.alloc_persistent((size_t)NF * n, s);
    CUDA_CHECK(cudaMemcpyAsync(pk->fixed_values.get(), fixed_host, (size_t)NF * n * sizeof(Fr), cudaMemcpyHostToDevice, s));
    // sigma columns from the copy-constraint cycles
    pk->sigma_values.alloc_persistent((size_t)P * n, s);
    {
        Assembly as(n, P);
        for (size_t i = 0; i < ncopies; ++i) as.copy(copies[4 * i], copies[4 * i + 1], copies[4 * i + 2], copies[4 * i + 3]);
        DevBuf<uint32_t> mc((size_t)P * n, s), mr((size_t)P * n, s);
        CUDA_CHECK(cudaMemcpyAsync(mc.get(), as.map_col.data(), (size_t)P * n * 4, cudaMemcpyHostToDevice, s));
        CUDA_CHECK(cudaMemcpyAsync(mr.get(), as.map_row.data(), (size_t)P * n * 4, cudaMemcpyHostToDevice, s));
        std::vector<Fr> dp(P);
        dp[0] = f_one<FrCfg>();
        for (uint32_t j = 1; j < P; ++j) dp[j] = f_mul(dp[j - 1], FrConsts::delta());
        DevBuf<Fr> dpd(P, s);
        CUDA_CHECK(cudaMemcpyAsync(dpd.get(), dp.data(), P * sizeof(Fr), cudaMemcpyHostToDevice, s));
        for (uint32_t j = 0; j < P; ++j)
            sigma_from_mapping(pk->sigma_values.get() + (size_t)j * n, mc.get() + (size_t)j * n, mr.get() + (size_t)j * n, dpd.get(), tw.t.get(),
                               tw.log_n, sh.k, s);
        CUDA_CHECK(cudaStreamSynchronize(s));  // host staging vectors go out of scope
    }
    // keygen_vk: commitments (Lagrange basis, no blinding)
    pk->fixed_commitments = commit_batch(ctx, 1, pk->fixed_values.get(), n, NF, n);
    pk->perm_commitments = commit_batch(ctx, 1, pk->sigma_values.get(), n, P, n);
    pk->transcript_repr = default_transcript_repr(*pk);
    // keygen_pk: coefficient forms and extended cosets
    pk->fixed_polys.alloc_persistent((size_t)NF * n, s);
    pk->sigma_polys.alloc_persistent((size_t)P * n, s);
    CUDA_CHECK(cudaMemcpyAsync(pk->fixed_polys.get(), pk->fixed_values.get(), (size_t)NF * n * sizeof(Fr), cudaMemcpyDeviceToDevice, s));
    CUDA_CHECK(cudaMemcpyAsync(pk->sigma_polys.get(), pk->sigma_values.get(), (size_t)P * n * sizeof(Fr), cudaMemcpyDeviceToDevice, s));
    for (uint32_t i = 0; i < NF; ++i) dev_lagrange_to_coeff(ctx, sh.k, pk->fixed_polys.get() + (size_t)i * n);
    for (uint32_t j = 0; j < P; ++j) dev_lagrange_to_coeff(ctx, sh.k, pk->sigma_polys.get() + (size_t)j * n);
    pk->fixed_cosets.alloc_persistent((size_t)NF * en, s);
    pk->sigma_cosets.alloc_persistent((size_t)P * en, s);
    for (uint32_t i = 0; i < NF; ++i) dev_coeff_to_extended(ctx, sh.k, pk->fixed_polys.get() + (size_t)i * n, pk->fixed_cosets.get() + (size_t)i * en);
    for (uint32_t j = 0; j < P; ++j) dev_coeff_to_extended(ctx, sh.k, pk->sigma_polys.get() + (size_t)j * n, pk->sigma_cosets.get() + (size_t)j * en);
    // l0, l_last, l_active_row = 1 - l_last - l_blind on the extended domain
    {
        pk->l_polys.alloc_persistent(3 * en, s);
        DevBuf<Fr> tmp(3 * n, s);
        CUDA_CHECK(cudaMemsetAsync(tmp.get(), 0, 3 * n * sizeof(Fr), s));
        const Fr one = f_one<FrCfg>();
        const uint32_t bf = Shape::blinding_factors;
        fr_fill(tmp.get(), one, 1, s);                              // l0: row 0
        fr_fill(tmp.get() + n + (n - bf - 1), one, 1, s);           // l_last: row n-bf-1
        fr_fill(tmp.get() + 2 * n + (n - bf), one, bf, s);          // l_blind: rows n-bf..n-1
        for (int j = 0; j < 3; ++j) {
            dev_lagrange_to_coeff(ctx, sh.k, tmp.get() + (size_t)j * n);
            dev_coeff_to_extended(ctx, sh.k, tmp.get() + (size_t)j * n, pk->l_polys.get() + (size_t)j * en);
        }
        // l_active = 1 - l_last - l_blind, written over the l_blind slot
        std::vector<const Fr*> ps = {pk->l_polys.get() + en, pk->l_polys.get() + 2 * en};
        const Fr minus_one = f_neg(one);
        DevBuf<Fr> act(en, s);
        fr_fill(act.get(), one, en, s);
        fr_lincomb(act.get(), ps, {minus_one, minus_one}, en, true, s);
        CUDA_CHECK(cudaMemcpyAsync(pk->l_polys.get() + 2 * en, act.get(), en * sizeof(Fr), cudaMemcpyDeviceToDevice, s));
    }
    CUDA_CHECK(cudaStreamSynchronize(s));
    (void)dom;
    return pk;
}

void pk_set_gates(Context& ctx, ProvingKeyDev& pk, const GateCalc* calcs, size_t ncalcs, const Fr* constants, size_t nconstants,
                  const uint32_t* results, size_t nresults) {
    const Shape& sh = pk.shape;
    cudaStream_t s = ctx.stream;
    if (ncalcs == 0) {  // back to the specialised kernel
        CUDA_CHECK(cudaStreamSynchronize(s));
        pk.gate_calcs.release();
        pk.gate_constants.release();
        pk.gate_results.release();
        return;
    }
    if (!calcs || !results || (nconstants && !constants)) throw std::invalid_argument("set_gates: null argument");
    if (ncalcs > GATE_MAX_CALCS || nconstants > GATE_MAX_CONSTANTS || nresults == 0 || nresults > GATE_MAX_RESULTS)
        throw std::invalid_argument("set_gates: program too large (128 calculations, 64 constants, 64 gates)");
    std::vector<uint32_t> degree(ncalcs, 0);
    auto src_degree = [&](const GateSrc& v, size_t j) -> uint32_t {
        switch (v.kind) {
            case GATE_SRC_CONSTANT:
                if (v.index >= nconstants) throw std::invalid_argument("set_gates: constant index out of range");
                return 0;
            case GATE_SRC_INTERMEDIATE:
                if (v.index >= j) throw std::invalid_argument("set_gates: an intermediate may only refer to an earlier calculation");
                return degree[v.index];
            case GATE_SRC_FIXED:
                if (v.index >= sh.num_fixed() || v.rotation != 0) throw std::invalid_argument("set_gates: fixed columns are queried at rotation 0 only");
                return 1;
            case GATE_SRC_ADVICE:
                // the proof opens gate columns at rotations 0..3 and lookup columns at 0: a program must stay inside that query set
                if (v.index >= sh.num_advice() || v.rotation < 0 || v.rotation > (v.index < sh.A ? 3 : 0))
                    throw std::invalid_argument("set_gates: advice rotation outside the proof's query set (gate columns 0..3, lookup columns 0)");
                return 1;
            default: throw std::invalid_argument("set_gates: unknown value source");
        }
    };
    for (size_t j = 0; j < ncalcs; ++j) {
        const GateCalc& c = calcs[j];
        const uint32_t da = src_degree(c.a, j);
        switch (c.op) {
            case GATE_ADD: case GATE_SUB: degree[j] = std::max(da, src_degree(c.b, j)); break;
            case GATE_MUL: degree[j] = da + src_degree(c.b, j); break;
            case GATE_SQUARE: degree[j] = 2 * da; break;
            case GATE_DOUBLE: case GATE_NEGATE: case GATE_STORE: degree[j] = da; break;
            default: throw std::invalid_argument("set_gates: unknown calculation");
        }
    }
    for (size_t g = 0; g < nresults; ++g) {
        if (results[g] >= ncalcs) throw std::invalid_argument("set_gates: result index out of range");
        if (degree[results[g]] > Shape::degree) throw std::invalid_argument("set_gates: gate degree exceeds the constraint system's degree (4)");
    }
    CUDA_CHECK(cudaStreamSynchronize(s));
    pk.gate_calcs.alloc_persistent(ncalcs, s);
    pk.gate_constants.alloc_persistent(std::max<size_t>(nconstants, 1), s);
    pk.gate_results.alloc_persistent(nresults, s);
    CUDA_CHECK(cudaMemcpyAsync(pk.gate_calcs.get(), calcs, ncalcs * sizeof(GateCalc), cudaMemcpyHostToDevice, s));
    if (nconstants) CUDA_CHECK(cudaMemcpyAsync(pk.gate_constants.get(), constants, nconstants * sizeof(Fr), cudaMemcpyHostToDevice, s));
    CUDA_CHECK(cudaMemcpyAsync(pk.gate_results.get(), results, nresults * 4, cudaMemcpyHostToDevice, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
}

// ---- SHPLONK bookkeeping (host) -----------------------------------------------------------------------------------------
struct Query {
    size_t poly;
    Fr point, eval;
};
static int fr_cmp(const Fr& a, const Fr& b) {  // numeric order of the canonical integers
    const Fr x = f_from_mont(a), y = f_from_mont(b);
    for (int i = 7; i >= 0; --i)
        if (x.l[i] != y.l[i]) return x.l[i] < y.l[i] ? -1 : 1;
    return 0;
}
struct RotationSet {
    std::vector<Fr> points;
    std::vector<size_t> polys;
    std::vector<std::vector<Fr>> evals;
};
static void sorted_insert(std::vector<Fr>& v, const Fr& x) {
    size_t i = 0;
    for (; i < v.size(); ++i) {
        int c = fr_cmp(x, v[i]);
        if (c == 0) return;
        if (c < 0) break;
    }
    v.insert(v.begin() + i, x);
}
static bool same_points(const std::vector<Fr>& a, const std::vector<Fr>& b) {
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); ++i)
        if (!f_eq(a[i], b[i])) return false;
    return true;
}
static void construct_intermediate_sets(const std::vector<Query>& queries, std::vector<RotationSet>& sets, std::vector<Fr>& super) {
    std::vector<std::pair<size_t, std::vector<Fr>>> by_poly;
    for (auto& q : queries) {
        sorted_insert(super, q.point);
        bool found = false;
        for (auto& e : by_poly)
            if (e.first == q.poly) {
                sorted_insert(e.second, q.point);
                found = true;
                break;
            }
        if (!found) by_poly.push_back({q.poly, {q.point}});
    }
    for (auto& e : by_poly) {
        bool found = false;
        for (auto& rs : sets)
            if (same_points(rs.points, e.second)) {
                rs.polys.push_back(e.first);
                found = true;
                break;
            }
        if (!found) {
            RotationSet rs;
            rs.points = e.second;
            rs.polys.push_back(e.first);
            sets.push_back(rs);
        }
    }
    for (auto& rs : sets)
        for (size_t id : rs.polys) {
            std::vector<Fr> ev;
            for (auto& pt : rs.points) {
                bool ok = false;
                for (auto& q : queries)
                    if (q.poly == id && f_eq(q.point, pt)) {
                        ev.push_back(q.eval);
                        ok = true;
                        break;
                    }
                if (!ok) throw std::runtime_error("shplonk: missing evaluation");
            }
            rs.evals.push_back(ev);
        }
}
static std::vector<Fr> lagrange_interpolate(const std::vector<Fr>& pts, const std::vector<Fr>& evals) {
    const size_t m = pts.size();
    std::vector<Fr> res(m, f_zero<FrCfg>());
    for (size_t j = 0; j < m; ++j) {
        std::vector<Fr> num(1, f_one<FrCfg>());
        Fr denom = f_one<FrCfg>();
        for (size_t i = 0; i < m; ++i) {
            if (i == j) continue;
            std::vector<Fr> nx(num.size() + 1, f_zero<FrCfg>());
            for (size_t t = 0; t < num.size(); ++t) {
                nx[t + 1] = f_add(nx[t + 1], num[t]);
                nx[t] = f_sub(nx[t], f_mul(num[t], pts[i]));
            }
            num = nx;
            denom = f_mul(denom, f_sub(pts[j], pts[i]));
        }
        const Fr sc = f_mul(evals[j], f_inv(denom));
        for (size_t t = 0; t < num.size(); ++t) res[t] = f_add(res[t], f_mul(num[t], sc));
    }
    return res;
}
static Fr eval_small(const std::vector<Fr>& p, const Fr& x) {
    Fr acc = f_zero<FrCfg>();
    for (size_t i = p.size(); i-- > 0;) acc = f_add(f_mul(acc, x), p[i]);
    return acc;
}
static Fr vanishing_eval(const std::vector<Fr>& roots, const Fr& z) {
    Fr r = f_one<FrCfg>();
    for (auto& x : roots) r = f_mul(r, f_sub(z, x));
    return r;
}

// evaluations of many polynomials at one point; when sharded, rank r evaluates polynomials r, r+world, ... (every rank
// holds all coefficient forms) and the 32-byte results are all-gathered over NCCL (Sharder::host_allgather)
static void eval_many_dist(Context& ctx, Sharder& shard, const std::vector<const Fr*>& polys, size_t n, const Fr& point, Fr* out) {
    const size_t m = polys.size(), world = ctx.world;
    if (!shard.on() || m < 2 * world) {
        fr_eval_many(ctx, polys, n, point, out);
        return;
    }
    const size_t per = (m + world - 1) / world;
    std::vector<const Fr*> mine;
    for (size_t i = ctx.rank; i < m; i += world) mine.push_back(polys[i]);
    std::vector<Fr> send(per, f_zero<FrCfg>()), recv(per * world);
    fr_eval_many(ctx, mine, n, point, send.data());
    shard.host_allgather(send.data(), per * sizeof(Fr), recv.data());
    for (size_t i = 0; i < m; ++i) out[i] = recv[(i % world) * per + i / world];
}

// ---- create_proof ---------------------------------------------------------------------------------------------------------
// rotation reach of the h(X) kernels in extended rows: -4·(blinding_factors+1) = -28 (z of the previous set) ... +12 (gate rotation 3)
static constexpr size_t HALO_BEFORE = 32, HALO_AFTER = 16;
// ownership offsets of the column families that are only transformed (not built) by their owner
static constexpr size_t OFF_ADVICE_NTT = 1, OFF_LOOKUP_COSETS = 2, OFF_PERM_IN = 0, OFF_PERM_TAB = 3, OFF_LOOKUP_Z = 6;
// lagrange_to_coeff of `count` columns that every rank holds: by column across the ranks, then one all-gather
static void lagrange_to_coeff_dist(Context& ctx, Sharder& shard, uint32_t k, Fr* base, uint32_t count, size_t n, size_t off) {
    if (count == 0) return;
    if (!shard.on()) {
        dev_lagrange_to_coeff(ctx, k, base, count, n);
        return;
    }
    for (uint32_t c = 0; c < count; ++c)
        if (shard.mine(c, off)) dev_lagrange_to_coeff(ctx, k, base + (size_t)c * n);
    shard.allgather_columns(base, count, n, off);
}
// Row H: evaluation::Evaluator::evaluate_h followed by divide_by_vanishing_poly (the t_inv scaling is fused into the last
// kernel), on the extended domain. Inputs: the advice and permutation-product cosets (NA resp. NS columns of 4n), and per
// lookup the coefficient forms of Z, a', s' (their cosets are made here, three at a time, so they never all coexist).
// Sharded: every rank evaluates a contiguous slice of the extended rows and h is all-gathered at the end.
template <class Lap>
static void evaluate_h_dev(Context& ctx, Sharder& shard, const ProvingKeyDev& pk, const Fr* advice_cosets, const Fr* z_cosets, const Fr* lk_z_poly,
                           const Fr* perm_in_poly, const Fr* perm_tab_poly, const Fr& y, const Fr& beta, const Fr& gamma, Fr* h, ProofTimings* tm,
                           Lap&& lap) {
    const Shape& sh = pk.shape;
    cudaStream_t s = ctx.stream;
    const size_t n = sh.n(), en = 4 * n;
    const uint32_t bf = Shape::blinding_factors, NA = sh.num_advice(), A = sh.A, L = sh.L, F = sh.F, P = sh.num_perm(), NS = sh.num_sets();
    const Domain& dom = ctx.domain(sh.k);
    const TwiddleTable& tw = ctx.std_table(sh.k + 2);
    if (shard.on() && en % ctx.world) throw std::runtime_error("sharded prover: world size must divide the extended domain");
    QuotientArgs Q{};
    Q.k = sh.k; Q.A = A; Q.L = L; Q.F = F; Q.P = P; Q.num_sets = NS; Q.blinding_factors = bf;
    Q.table = tw.t.get();
    Q.table_log = tw.log_n;
    Q.t_inv = dom.t_inv_dev();
    for (uint32_t c = 0; c < NA; ++c) Q.advice[c] = advice_cosets + (size_t)c * en;
    for (uint32_t i = 0; i < sh.num_fixed(); ++i) Q.fixed[i] = pk.fixed_cosets.get() + (size_t)i * en;
    for (uint32_t j = 0; j < P; ++j) {
        Q.perm_cols[j] = sh.perm_is_fixed(j) ? Q.fixed[sh.perm_col_index(j)] : Q.advice[sh.perm_col_index(j)];
        Q.sigma[j] = pk.sigma_cosets.get() + (size_t)j * en;
    }
    for (uint32_t set = 0; set < NS; ++set) Q.z[set] = z_cosets + (size_t)set * en;
    Q.l0 = pk.l_polys.get();
    Q.l_last = pk.l_polys.get() + en;
    Q.l_active = pk.l_polys.get() + 2 * en;
    Q.y = y; Q.beta = beta; Q.gamma = gamma; Q.delta = FrConsts::delta();
    Q.beta_zeta = f_mul(beta, FrConsts::zeta());
    const size_t rows_per_rank = shard.on() ? en / ctx.world : en;
    Q.row_begin = shard.on() ? rows_per_rank * ctx.rank : 0;
    Q.row_end = Q.row_begin + rows_per_rank;
    if (pk.gate_calcs.size()) h_gates_program(Q, pk.gate_program(), h, s);
    else h_gates(Q, h, s);
    h_permutation(Q, h, L == 0, s);
    lap(tm ? &tm->quotient : nullptr);
    if (L) {
        // the cosets of Z, a', s' are made for two lookups at a time (six columns: enough to occupy six ranks at once) and
        // never all coexist
        const uint32_t GL = L >= 2 ? 2 : 1;
        DevBuf<Fr> lc((size_t)3 * GL * en, s);
        for (uint32_t l0 = 0; l0 < L; l0 += GL) {
            const uint32_t g = std::min(GL, L - l0);
            Fr* cs[6];
            for (uint32_t q = 0; q < 3 * g; ++q) {
                const uint32_t l = l0 + q / 3;
                const Fr* src = (q % 3 == 0 ? lk_z_poly : q % 3 == 1 ? perm_in_poly : perm_tab_poly) + (size_t)l * n;
                cs[q] = lc.get() + (size_t)q * en;
                if (shard.mine(3 * l0 + q, OFF_LOOKUP_COSETS)) dev_coeff_to_extended(ctx, sh.k, src, cs[q]);
            }
            shard.exchange_row_slices(cs, 3 * g, [&](size_t q) { return shard.owner(3 * l0 + q, OFF_LOOKUP_COSETS); }, en, HALO_BEFORE, HALO_AFTER);
            lap(tm ? &tm->ntt : nullptr);
            for (uint32_t i = 0; i < g; ++i) {
                const uint32_t l = l0 + i;
                LookupCosets Lk{cs[3 * i], cs[3 * i + 1], cs[3 * i + 2], Q.advice[A + l], Q.fixed[sh.table_col()]};
                h_lookup(Q, Lk, h, l + 1 == L, s);
            }
            lap(tm ? &tm->quotient : nullptr);
        }
    }
    if (shard.on()) {
        shard.all_gather_inplace(h, en / ctx.world);
        lap(tm ? &tm->quotient : nullptr);
    }
}

// b200zk_evaluate_h: row H on its own, for a host that keeps halo2's create_proof and replaces only evaluate_h. All
// polynomials arrive in coefficient form (as halo2 holds them at that point); returns h on the extended domain, already
// divided by the vanishing polynomial. Runs unsharded on the calling rank's GPU.
void evaluate_h(Context& ctx, const ProvingKeyDev& pk, const Fr* advice_coeff, const Fr* z_coeff, const Fr* lookup_coeff, const Fr& y, const Fr& beta,
                const Fr& gamma, Fr* h_out) {
    const Shape& sh = pk.shape;
    cudaStream_t s = ctx.stream;
    const size_t n = sh.n(), en = 4 * n;
    const uint32_t NA = sh.num_advice(), L = sh.L, NS = sh.num_sets();
    Sharder shard(ctx);
    shard.enabled = false;
    DevBuf<Fr> polys((size_t)std::max<uint32_t>(NA, std::max<uint32_t>(NS, 3 * L)) * n, s);
    DevBuf<Fr> advice_cosets((size_t)NA * en, s), z_cosets((size_t)NS * en, s), h(en, s);
    auto to_cosets = [&](const Fr* host, uint32_t count, Fr* cosets) {
        CUDA_CHECK(cudaMemcpyAsync(polys.get(), host, (size_t)count * n * sizeof(Fr), cudaMemcpyHostToDevice, s));
        for (uint32_t c = 0; c < count; ++c) dev_coeff_to_extended(ctx, sh.k, polys.get() + (size_t)c * n, cosets + (size_t)c * en);
    };
    to_cosets(advice_coeff, NA, advice_cosets.get());
    to_cosets(z_coeff, NS, z_cosets.get());
    // lookup_coeff: per lookup l the three polynomials Z_l, a'_l, s'_l, each n coefficients
    DevBuf<Fr> lk((size_t)3 * L * n, s);
    for (uint32_t l = 0; l < L; ++l)
        for (uint32_t j = 0; j < 3; ++j)
            CUDA_CHECK(cudaMemcpyAsync(lk.get() + ((size_t)j * L + l) * n, lookup_coeff + ((size_t)3 * l + j) * n, n * sizeof(Fr), cudaMemcpyHostToDevice, s));
    evaluate_h_dev(ctx, shard, pk, advice_cosets.get(), z_cosets.get(), lk.get(), lk.get() + (size_t)L * n, lk.get() + (size_t)2 * L * n, y, beta, gamma,
                   h.get(), nullptr, [](double*) {});
    CUDA_CHECK(cudaMemcpyAsync(h_out, h.get(), en * sizeof(Fr), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
}

static std::vector<uint8_t> create_proof_body(Context& ctx, const ProvingKeyDev& pk, const Fr* advice_in, bool advice_on_device,
                                              host::FrRandomStream& rng, ProofTimings* tm);
std::vector<uint8_t> create_proof(Context& ctx, const ProvingKeyDev& pk, const Fr* advice_in, bool advice_on_device, host::FrRandomStream& rng,
                                  ProofTimings* tm) {
    try {
        return create_proof_body(ctx, pk, advice_in, advice_on_device, rng, tm);
    } catch (...) {
        // the transient columns have gone back to the arena (bookkeeping only): an exchange still running on the comm stream
        // must drain before the next call can hand that memory out again
        if (ctx.comm_stream) cudaStreamSynchronize(ctx.comm_stream);
        if (ctx.ntt_stream) cudaStreamSynchronize(ctx.ntt_stream);
        ctx.comm_pending = false;
        throw;
    }
}
static std::vector<uint8_t> create_proof_body(Context& ctx, const ProvingKeyDev& pk, const Fr* advice_in, bool advice_on_device,
                                              host::FrRandomStream& rng, ProofTimings* tm) {
    const Shape& sh = pk.shape;
    need_srs(ctx, sh.k);
    cudaStream_t s = ctx.stream;
    const size_t n = sh.n(), en = 4 * n, u = sh.usable_rows();
    const uint32_t bf = Shape::blinding_factors, NA = sh.num_advice(), A = sh.A, L = sh.L, F = sh.F, P = sh.num_perm(), NS = sh.num_sets();
    const Domain& dom = ctx.domain(sh.k);
    const TwiddleTable& tw = ctx.std_table(sh.k + 2);
    host::Transcript tr(ctx.compat.point_sign_bit);
    // [UNVERIFIED-1] Blind(..) scalars: drawn upstream (the stream advances) but unused by KZG commitments
    auto skip_unused_blinds = [&](uint64_t count) {
        if (ctx.compat.draw_unused_blinds) rng.skip(count);
    };
    const double exchange0 = ctx.exchange_seconds;
    struct TraceGuard {  // collectives are traced only while a timed call is running
        Context& c;
        bool was;
        ~TraceGuard() { c.comm_trace = was; }
    } trace_guard{ctx, ctx.comm_trace};
    ctx.comm_trace = tm != nullptr;
    const double comm0 = ctx.comm_seconds;
    Sharder shard(ctx);
    if (shard.on()) shard.nccl();  // communicator up before the first timed exchange
    auto clock_now = [&]() {
        CUDA_CHECK(cudaStreamSynchronize(s));
        return std::chrono::steady_clock::now();
    };
    auto t_start = clock_now();
    auto lap = [&](double* slot) {
        if (!tm) return;
        auto t = clock_now();
        *slot += std::chrono::duration<double>(t - t_start).count();
        t_start = t;
    };
    const Fr one = f_one<FrCfg>();

    // step 0
    tr.common_scalar(pk.transcript_repr);
    // step 1: upload, blind, commit advice (D.3)
    DevBuf<Fr> advice((size_t)NA * n, s);
    // The advice columns' coefficient forms and extended cosets do not depend on the transcript: once the blinded columns
    // are on the device they are computed on the side stream, beside the commit batches (which leave part of the
    // multiplier pipe idle in their latency-bound phases); the main stream joins right before h(X) needs them.
    DevBuf<Fr> advice_polys((size_t)NA * n, s), advice_cosets((size_t)NA * en, s);
    const bool side_ntt = ctx.ntt_stream != nullptr && !g_prof_enabled;
    auto advice_transforms = [&](bool side) {  // sharded: column c is transformed by rank (c + OFF_ADVICE_NTT) mod world
        cudaStream_t st = side ? ctx.ntt_stream : s;
        if (!shard.on()) {
            CUDA_CHECK(cudaMemcpyAsync(advice_polys.get(), advice.get(), (size_t)NA * n * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
            dev_lagrange_to_coeff(ctx, sh.k, advice_polys.get(), NA, n, side);
            for (uint32_t c = 0; c < NA; ++c)
                dev_coeff_to_extended(ctx, sh.k, advice_polys.get() + (size_t)c * n, advice_cosets.get() + (size_t)c * en, 1, 0, 0, side);
            return;
        }
        for (uint32_t c = 0; c < NA; ++c)
            if (shard.mine(c, OFF_ADVICE_NTT)) {
                CUDA_CHECK(cudaMemcpyAsync(advice_polys.get() + (size_t)c * n, advice.get() + (size_t)c * n, n * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
                dev_lagrange_to_coeff(ctx, sh.k, advice_polys.get() + (size_t)c * n, 1, 0, side);
                dev_coeff_to_extended(ctx, sh.k, advice_polys.get() + (size_t)c * n, advice_cosets.get() + (size_t)c * en, 1, 0, 0, side);
            }
    };
    // starts the side-stream transforms once everything queued so far on the main stream (and, if given, `also`) is done
    auto fork_advice_transforms = [&](cudaEvent_t also) {
        if (!side_ntt) return;
        CUDA_CHECK(cudaEventRecord(ctx.ntt_fork, s));
        CUDA_CHECK(cudaStreamWaitEvent(ctx.ntt_stream, ctx.ntt_fork, 0));
        if (also) CUDA_CHECK(cudaStreamWaitEvent(ctx.ntt_stream, also, 0));
        advice_transforms(true);
        CUDA_CHECK(cudaEventRecord(ctx.ntt_done, ctx.ntt_stream));
    };
    std::vector<Fr> blind((size_t)NA * (bf + 1));
    for (auto& b : blind) b = rng.next();
    skip_unused_blinds(NA);  // Blind(..) per advice column
    // the copy stream must be idle before `advice` can go back to the arena, whichever way this function is left
    struct CopyJoin {
        cudaStream_t st;
        ~CopyJoin() {
            if (st) cudaStreamSynchronize(st);
        }
    } copy_join{nullptr};
    // a pageable witness (a Rust Vec) goes through pinned staging chunks filled by worker threads (upload.cuh)
    const bool pageable = !advice_on_device && host_pointer_is_pageable(advice_in);
    // (declared before the stager: its worker threads report into it and are joined by the stager's destructor)
    struct ColumnsReady {  // which columns have been queued (with their event recorded): set by the uploaders, read by the gate
        std::mutex mu;
        std::condition_variable cv;
        std::vector<uint8_t> ready;
        bool failed = false;
    } cols_ready;
    StagedUpload stager(ctx);
    // columns [c0, c1) from the caller's buffer, queued on `st`; with `wait` the call returns once everything is queued
    auto upload = [&](uint32_t c0, uint32_t c1, cudaStream_t st, bool wait) {
        if (c0 >= c1) return;
        Fr* dst = advice.get() + (size_t)c0 * n;
        const Fr* src = advice_in + (size_t)c0 * n;
        const size_t bytes = (size_t)(c1 - c0) * n * sizeof(Fr);
        if (pageable) {
            stager.start(dst, src, bytes, st);
            if (wait) stager.join();
        } else {
            CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, advice_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
        }
    };
    // blinding rows of columns [c0, c1): must be queued behind the column data
    auto blind_rows = [&](uint32_t c0, uint32_t c1, cudaStream_t st) {
        for (uint32_t c = c0; c < c1; ++c)
            CUDA_CHECK(cudaMemcpyAsync(advice.get() + (size_t)c * n + u, blind.data() + (size_t)c * (bf + 1), (bf + 1) * sizeof(Fr),
                                       cudaMemcpyHostToDevice, st));
    };
    // host witness on one GPU: the columns go up on the copy stream one after the other, each followed by its blinding rows
    // and an event; ONE commit batch is issued at once and each column's first kernel waits for that column only
    // (Context::column_gate), so the commitments start when the first column has landed and the upload hides behind them
    const bool streamed = !advice_on_device && !shard.on() && ctx.copy_stream != nullptr;
    struct GateReset {  // the gate must not outlive this call
        Context& c;
        ~GateReset() { c.column_gate = nullptr; }
    } gate_reset{ctx};
    if (streamed) {
        cols_ready.ready.assign(NA, 0);
        for (uint32_t c = 0; c < NA; ++c) ctx.column_event(c);  // create the events before any worker thread needs one
        CUDA_CHECK(cudaEventRecord(ctx.copy_fork, s));
        CUDA_CHECK(cudaStreamWaitEvent(ctx.copy_stream, ctx.copy_fork, 0));
        copy_join.st = ctx.copy_stream;
        const int device = ctx.device;
        // queued behind column c's data on the copy stream: its blinding rows, then the event the commit batch waits for
        auto finish_column = [&, device](size_t c) {
            bool ok = cudaSetDevice(device) == cudaSuccess;
            ok = ok && cudaMemcpyAsync(advice.get() + c * n + u, blind.data() + c * (bf + 1), (bf + 1) * sizeof(Fr), cudaMemcpyHostToDevice,
                                       ctx.copy_stream) == cudaSuccess;
            ok = ok && cudaEventRecord(ctx.column_events[c], ctx.copy_stream) == cudaSuccess;
            std::lock_guard<std::mutex> lk(cols_ready.mu);
            cols_ready.ready[c] = 1;
            cols_ready.failed = cols_ready.failed || !ok;
            cols_ready.cv.notify_all();
        };
        if (pageable) {
            stager.start(advice.get(), advice_in, (size_t)NA * n * sizeof(Fr), ctx.copy_stream, n * sizeof(Fr), finish_column);
        } else {
            for (uint32_t c = 0; c < NA; ++c) {
                CUDA_CHECK(cudaMemcpyAsync(advice.get() + (size_t)c * n, advice_in + (size_t)c * n, n * sizeof(Fr), cudaMemcpyHostToDevice, ctx.copy_stream));
                finish_column(c);
            }
        }
        ctx.column_gate = [&](size_t c, cudaStream_t st) {
            std::unique_lock<std::mutex> lk(cols_ready.mu);
            cols_ready.cv.wait(lk, [&] { return cols_ready.ready[c] != 0 || cols_ready.failed; });
            if (cols_ready.failed) throw CudaError("witness upload failed");
            lk.unlock();
            CUDA_CHECK(cudaStreamWaitEvent(st, ctx.column_events[c], 0));
        };
        lap(tm ? &tm->upload : nullptr);
        for (const G1Affine& cm : commit_batch(ctx, 1, advice.get(), n, NA, n)) tr.write_point(cm);
        ctx.column_gate = nullptr;
        stager.join();
        CUDA_CHECK(cudaEventRecord(ctx.copy_done, ctx.copy_stream));
        fork_advice_transforms(ctx.copy_done);
        CUDA_CHECK(cudaStreamWaitEvent(s, ctx.copy_done, 0));
    }
    const uint32_t ahead = NA;
    if (streamed) {
        // done above
    } else if (shard.on() && !advice_on_device) {
        // every rank holds the host witness: each uploads only its share over PCIe — ONE contiguous block of columns, so a
        // pageable buffer goes through a single staged upload — and the rest arrives over NVLink (in-place broadcasts of the
        // blocks, one NCCL group)
        const uint32_t c_lo = (uint32_t)((uint64_t)NA * ctx.rank / ctx.world), c_hi = (uint32_t)((uint64_t)NA * (ctx.rank + 1) / ctx.world);
        upload(c_lo, c_hi, s, true);
        shard.broadcast_blocks(advice.get(), NA, n);
        blind_rows(0, NA, s);
        fork_advice_transforms(nullptr);
    } else {
        upload(0, NA, s, true);
        blind_rows(0, NA, s);
        fork_advice_transforms(nullptr);
    }
    if (!streamed) {
        CUDA_CHECK(cudaStreamSynchronize(s));
        lap(tm ? &tm->upload : nullptr);
        for (const G1Affine& cm : commit_batch(ctx, 1, advice.get(), n, ahead, n)) tr.write_point(cm);
    }
    lap(tm ? &tm->msm : nullptr);
    const Fr theta = tr.squeeze_challenge();
    (void)theta;  // single-expression lookups: theta-compression is the identity
    // step 3: lookups, permuted columns (D.4)
    const Fr* table_values = pk.fixed_values.get() + (size_t)sh.table_col() * n;
    DevBuf<Fr> perm_in((size_t)L * n, s), perm_tab((size_t)L * n, s), perm_in_poly((size_t)L * n, s), perm_tab_poly((size_t)L * n, s);
    // perm_cols holds a'_0, s'_0, a'_1, s'_1, ... so that one batch commits them in transcript order
    DevBuf<Fr> perm_cols((size_t)2 * L * n, s);
    uint32_t lookup_failed = 0;
    for (uint32_t l = 0; l < L; ++l) {  // sharded: lookup l is permuted by rank l mod world
        Fr* a_out = perm_cols.get() + (size_t)(2 * l) * n;
        Fr* s_out = perm_cols.get() + (size_t)(2 * l + 1) * n;
        std::vector<Fr> blind(2 * (bf + 1));
        for (auto& b : blind) b = rng.next();
        skip_unused_blinds(2);  // the two Blind(..) draws of commit_values
        if (!shard.mine(l)) continue;
        if (const int st = lookup_permute(ctx, advice.get() + (size_t)(A + l) * n, table_values, a_out, s_out, n, u, ctx.compat.lookup_fill_from_end)) {
            lookup_failed |= (uint32_t)st;
            continue;
        }
        CUDA_CHECK(cudaMemcpyAsync(a_out + u, blind.data(), (bf + 1) * sizeof(Fr), cudaMemcpyHostToDevice, s));
        CUDA_CHECK(cudaMemcpyAsync(s_out + u, blind.data() + bf + 1, (bf + 1) * sizeof(Fr), cudaMemcpyHostToDevice, s));
        CUDA_CHECK(cudaStreamSynchronize(s));
    }
    if (shard.on() && L) {  // a failure seen by one rank must stop every rank before the next collective
        std::vector<uint32_t> all(ctx.world);
        shard.host_allgather(&lookup_failed, sizeof(uint32_t), all.data());
        for (uint32_t f : all) lookup_failed |= f;
    }
    if (lookup_failed & LOOKUP_UNSUPPORTED) throw std::invalid_argument("create_proof: lookup table values must be < 2^k (range-style tables only)");
    if (lookup_failed) throw SynthesisError("ConstraintSystemFailure: lookup input not in table");
    shard.allgather_columns(perm_cols.get(), L, 2 * n);  // column l = a'_l followed by s'_l
    lap(tm ? &tm->lookup : nullptr);
    {
        const std::vector<G1Affine> cms = commit_batch(ctx, 1, perm_cols.get(), n, 2 * L, n);
        lap(tm ? &tm->msm : nullptr);
        for (uint32_t l = 0; l < L; ++l) {
            CUDA_CHECK(cudaMemcpyAsync(perm_in.get() + (size_t)l * n, perm_cols.get() + (size_t)(2 * l) * n, n * sizeof(Fr), cudaMemcpyDeviceToDevice, s));
            CUDA_CHECK(cudaMemcpyAsync(perm_tab.get() + (size_t)l * n, perm_cols.get() + (size_t)(2 * l + 1) * n, n * sizeof(Fr), cudaMemcpyDeviceToDevice, s));
        }
        CUDA_CHECK(cudaMemcpyAsync(perm_in_poly.get(), perm_in.get(), (size_t)L * n * sizeof(Fr), cudaMemcpyDeviceToDevice, s));
        CUDA_CHECK(cudaMemcpyAsync(perm_tab_poly.get(), perm_tab.get(), (size_t)L * n * sizeof(Fr), cudaMemcpyDeviceToDevice, s));
        if (L) {
            lagrange_to_coeff_dist(ctx, shard, sh.k, perm_in_poly.get(), L, n, OFF_PERM_IN);
            lagrange_to_coeff_dist(ctx, shard, sh.k, perm_tab_poly.get(), L, n, OFF_PERM_TAB);
        }
        lap(tm ? &tm->ntt : nullptr);
        for (const G1Affine& cm : cms) tr.write_point(cm);
    }
    const Fr beta = tr.squeeze_challenge();
    const Fr gamma = tr.squeeze_challenge();
    // step 5: permutation grand products (D.5)
    auto perm_values = [&](uint32_t j) -> const Fr* {
        return sh.perm_is_fixed(j) ? pk.fixed_values.get() + (size_t)sh.perm_col_index(j) * n : advice.get() + (size_t)sh.perm_col_index(j) * n;
    };
    DevBuf<Fr> z_polys((size_t)NS * n, s), z_cosets((size_t)NS * en, s);
    {
        std::vector<Fr> dpow(P);  // delta^j·beta for permutation column j
        dpow[0] = beta;
        for (uint32_t j = 1; j < P; ++j) dpow[j] = f_mul(dpow[j - 1], FrConsts::delta());
        // The chain z_s[0] = z_{s-1}[n-bf-1] only couples the sets through one scalar: every set is first built as a running
        // product from 1 (all denominators inverted by ONE batch inversion), then the end values E_s are read back and set s
        // is rescaled by the carry prod_{t<s} E_t — the same field elements as the sequential chain, without a host round
        // trip per set. Sharded: set s is built by rank s mod world (the rank that also commits and transforms it) and the
        // end values are exchanged.
        const bool dist_sets = shard.on() && NS >= (uint32_t)ctx.world;
        auto builds = [&](uint32_t set) { return !dist_sets || shard.mine(set); };
        std::vector<uint32_t> my_sets;
        for (uint32_t set = 0; set < NS; ++set)
            if (builds(set)) my_sets.push_back(set);
        std::vector<Fr> E(NS, f_zero<FrCfg>());
        {
            DevBuf<Fr> m(my_sets.size() * n, s);
            for (size_t li = 0; li < my_sets.size(); ++li) {
                const uint32_t j0 = my_sets[li] * Shape::chunk_len, j1 = std::min(P, j0 + Shape::chunk_len);
                for (uint32_t j = j0; j < j1; ++j)
                    perm_denominator(m.get() + li * n, perm_values(j), pk.sigma_values.get() + (size_t)j * n, beta, gamma, n, j == j0, s);
            }
            fr_batch_invert(m.get(), my_sets.size() * n, s);
            for (size_t li = 0; li < my_sets.size(); ++li) {
                const uint32_t set = my_sets[li], j0 = set * Shape::chunk_len, j1 = std::min(P, j0 + Shape::chunk_len);
                Fr* z = z_polys.get() + (size_t)set * n;  // Lagrange values first, converted in place after the commit
                for (uint32_t j = j0; j < j1; ++j) perm_numerator(m.get() + li * n, perm_values(j), dpow[j], gamma, tw.t.get(), tw.log_n, sh.k, s);
                fr_prefix_product(z, m.get() + li * n, one, n, s);
                CUDA_CHECK(cudaMemcpyAsync(&E[set], z + (n - bf - 1), sizeof(Fr), cudaMemcpyDeviceToHost, s));
            }
            CUDA_CHECK(cudaStreamSynchronize(s));
        }
        if (dist_sets) {
            std::vector<Fr> all((size_t)NS * ctx.world);
            shard.host_allgather(E.data(), NS * sizeof(Fr), all.data());
            for (uint32_t set = 0; set < NS; ++set) E[set] = all[(size_t)shard.owner(set) * NS + set];
        }
        std::vector<Fr> blind((size_t)NS * bf);
        Fr carry = one;
        for (uint32_t set = 0; set < NS; ++set) {
            Fr* z = z_polys.get() + (size_t)set * n;
            for (uint32_t i = 0; i < bf; ++i) blind[(size_t)set * bf + i] = rng.next();  // every rank draws every value: the streams stay in step
            skip_unused_blinds(1);
            if (builds(set)) {
                if (set > 0) fr_scale(z, carry, n, s);
                CUDA_CHECK(cudaMemcpyAsync(z + (n - bf), blind.data() + (size_t)set * bf, bf * sizeof(Fr), cudaMemcpyHostToDevice, s));
            }
            carry = f_mul(carry, E[set]);
        }
        CUDA_CHECK(cudaStreamSynchronize(s));
        lap(tm ? &tm->products : nullptr);
    }
    // step 6: lookup grand products (D.6)
    DevBuf<Fr> lk_z_poly((size_t)L * n, s);
    {
        // sharded: lookup l is built by rank l mod world and broadcast; one batch inversion serves all of a rank's lookups
        std::vector<Fr> blind((size_t)L * bf);
        std::vector<uint32_t> my_lookups;
        for (uint32_t l = 0; l < L; ++l) {
            for (uint32_t i = 0; i < bf; ++i) blind[(size_t)l * bf + i] = rng.next();
            skip_unused_blinds(1);
            if (shard.mine(l)) my_lookups.push_back(l);
        }
        DevBuf<Fr> p(my_lookups.size() * n, s);
        for (size_t li = 0; li < my_lookups.size(); ++li) {
            const uint32_t l = my_lookups[li];
            lookup_denominator(p.get() + li * n, perm_in.get() + (size_t)l * n, perm_tab.get() + (size_t)l * n, beta, gamma, n, s);
        }
        fr_batch_invert(p.get(), my_lookups.size() * n, s);
        for (size_t li = 0; li < my_lookups.size(); ++li) {
            const uint32_t l = my_lookups[li];
            Fr* z = lk_z_poly.get() + (size_t)l * n;
            lookup_numerator(p.get() + li * n, advice.get() + (size_t)(A + l) * n, table_values, beta, gamma, n, s);
            fr_prefix_product(z, p.get() + li * n, one, n, s);
            CUDA_CHECK(cudaMemcpyAsync(z + (n - bf), blind.data() + (size_t)l * bf, bf * sizeof(Fr), cudaMemcpyHostToDevice, s));
        }
        CUDA_CHECK(cudaStreamSynchronize(s));
        shard.allgather_columns(lk_z_poly.get(), L, n);
        lap(tm ? &tm->products : nullptr);
    }
    // step 7: vanishing::commit (D.7): n sequential Fr::random draws = n consecutive ChaCha blocks, generated in place
    DevBuf<Fr> random_poly(n, s);
    if (rng.external()) {
        // the host's own RngCore: the n draws are pulled through its fill_bytes (64 bytes each, in chunks), uploaded and
        // reduced on the device; the chunk-seeded variant pulls its seeds the same way and expands them on the device
        if (shard.on()) throw std::invalid_argument("create_proof: an external random source cannot be shared by several ranks; use the seeded entry point");
        if (ctx.compat.random_poly_chunks == 0) {
            const size_t piece = (size_t)1 << 16;  // draws per upload
            std::vector<uint32_t> host(16 * std::min(piece, n));
            DevBuf<uint32_t> words(16 * std::min(piece, n), s);
            for (size_t lo = 0; lo < n; lo += piece) {
                const size_t cnt = std::min(piece, n - lo);
                rng.words(host.data(), 16 * cnt);
                CUDA_CHECK(cudaMemcpyAsync(words.get(), host.data(), 64 * cnt, cudaMemcpyHostToDevice, s));
                fr_from_u512(random_poly.get() + lo, words.get(), cnt, s);
                CUDA_CHECK(cudaStreamSynchronize(s));  // `host` is refilled next
            }
        } else {
            const size_t T = std::min<size_t>(ctx.compat.random_poly_chunks, n), chunk = n / T, n_chunks = T + (n % T != 0 ? 1 : 0);
            for (size_t c = 0; c < n_chunks; ++c) {
                uint8_t seed[32];
                rng.fill_bytes32(seed);
                const host::FrRandomStream sub = host::FrRandomStream::chacha20_from_seed(seed);
                const size_t lo = c * chunk, hi = std::min(n, (c + 1) * chunk);
                fr_random_stream(random_poly.get() + lo, hi - lo, sub.key, 0, sub.rounds, s);
            }
        }
    } else if (ctx.compat.random_poly_chunks == 0) {  // [UNVERIFIED-3] n sequential draws = n consecutive blocks of the main stream
        if (!rng.aligned()) throw std::logic_error("create_proof: random stream not block aligned");
        fr_random_stream(random_poly.get(), n, rng.key, rng.block_index(), rng.rounds, s);
        rng.skip(n);
    } else {  // one ChaCha20Rng per worker chunk, seeded from the main stream (T chunks of n / T, plus one for a remainder)
        const size_t T = std::min<size_t>(ctx.compat.random_poly_chunks, n), chunk = n / T, n_chunks = T + (n % T != 0 ? 1 : 0);
        std::vector<host::FrRandomStream> seeds;
        for (size_t c = 0; c < n_chunks; ++c) {
            uint8_t seed[32];
            rng.fill_bytes32(seed);
            seeds.push_back(host::FrRandomStream::chacha20_from_seed(seed));
        }
        for (size_t c = 0; c < n_chunks; ++c) {
            const size_t lo = c * chunk, hi = std::min(n, (c + 1) * chunk);
            fr_random_stream(random_poly.get() + lo, hi - lo, seeds[c].key, 0, seeds[c].rounds, s);
        }
    }
    skip_unused_blinds(1);
    lap(tm ? &tm->other : nullptr);
    // The permutation products, the lookup products (Lagrange basis) and the random polynomial (coefficient basis) are
    // written to the transcript back to back with no challenge in between: ONE commit batch, one bucket reduction. Sharded,
    // column j of the batch is committed by rank j mod world — for the leading z columns that is the rank that built them.
    {
        std::vector<const Fr*> cols;
        std::vector<int> basis;
        for (uint32_t set = 0; set < NS; ++set) cols.push_back(z_polys.get() + (size_t)set * n), basis.push_back(1);
        for (uint32_t l = 0; l < L; ++l) cols.push_back(lk_z_poly.get() + (size_t)l * n), basis.push_back(1);
        cols.push_back(random_poly.get()), basis.push_back(0);
        if (shard.on() && NS >= (uint32_t)ctx.world) {
            // z columns beyond the dealt part of the batch are committed by point range on EVERY rank: hand them out first
            const size_t dealt = cols.size() / ctx.world * ctx.world;
            for (size_t set = dealt; set < NS; ++set) shard.broadcast(z_polys.get() + set * n, n, shard.owner(set));
        }
        std::vector<G1Affine> cms(cols.size());
        msm_batch_srs_mixed(ctx, basis.data(), cols.data(), cols.size(), n, cms.data());
        lap(tm ? &tm->msm : nullptr);
        for (const G1Affine& cm : cms) tr.write_point(cm);
    }
    if (!shard.on()) {
        dev_lagrange_to_coeff(ctx, sh.k, z_polys.get(), NS, n);
        for (uint32_t set = 0; set < NS; ++set) dev_coeff_to_extended(ctx, sh.k, z_polys.get() + (size_t)set * n, z_cosets.get() + (size_t)set * en);
    } else {
        for (uint32_t set = 0; set < NS; ++set)
            if (shard.mine(set)) {
                dev_lagrange_to_coeff(ctx, sh.k, z_polys.get() + (size_t)set * n);
                dev_coeff_to_extended(ctx, sh.k, z_polys.get() + (size_t)set * n, z_cosets.get() + (size_t)set * en);
            }
        shard.allgather_columns_async(z_polys.get(), NS, n);  // coefficient forms are first read by the evaluations at x
        std::vector<Fr*> cs(NS);
        for (uint32_t set = 0; set < NS; ++set) cs[set] = z_cosets.get() + (size_t)set * en;
        shard.exchange_row_slices(cs.data(), NS, [&](size_t c) { return shard.owner(c); }, en, HALO_BEFORE, HALO_AFTER);
    }
    lagrange_to_coeff_dist(ctx, shard, sh.k, lk_z_poly.get(), L, n, OFF_LOOKUP_Z);
    lap(tm ? &tm->ntt : nullptr);
    const Fr y = tr.squeeze_challenge();
    // step 8/9: advice polys + cosets, h(X) (D.8)
    if (side_ntt) CUDA_CHECK(cudaStreamWaitEvent(s, ctx.ntt_done, 0));
    else advice_transforms(false);
    if (shard.on()) {  // coefficient forms: overlapped all-gather (first read by the evaluations); cosets: row slices
        shard.allgather_columns_async(advice_polys.get(), NA, n, OFF_ADVICE_NTT);
        std::vector<Fr*> cs(NA);
        for (uint32_t c = 0; c < NA; ++c) cs[c] = advice_cosets.get() + (size_t)c * en;
        shard.exchange_row_slices(cs.data(), NA, [&](size_t c) { return shard.owner(c, OFF_ADVICE_NTT); }, en, HALO_BEFORE, HALO_AFTER);
    }
    lap(tm ? &tm->ntt : nullptr);
    DevBuf<Fr> h(en, s);
    evaluate_h_dev(ctx, shard, pk, advice_cosets.get(), z_cosets.get(), lk_z_poly.get(), perm_in_poly.get(), perm_tab_poly.get(), y, beta, gamma,
                   h.get(), tm, lap);
    // step 10: vanishing::construct (D.9) — t_inv scaling already applied by the last h kernel
    DevBuf<Fr> h_coeff(3 * n, s);
    if (shard.on() && n % ctx.world == 0) {
        // the 4n-point inverse transform of h across the ranks: the four stride-4 subsequences are transformed by (up to)
        // four ranks (size n each), exchanged, and the radix-4 combine + scaling + truncation is done by slice of k
        DevBuf<Fr> Y(4 * n, s);
        NttPlan p = make_plan(ctx, sh.k, true);  // plain inverse transform: root ω⁻¹, no divisor
        p.in_stride = 4;
        for (uint32_t j = 0; j < 4; ++j)
            if (shard.mine(j)) {
                p.in_offset = j;
                ntt_run_batch(p, h.get(), Y.get() + (size_t)j * n, ctx.get_scratch(n), 1, 0, 0, 0, s);
            }
        shard.allgather_columns(Y.get(), 4, n);
        const size_t len = n / ctx.world, k_lo = len * ctx.rank;
        fr_e2c_combine(Y.get(), h_coeff.get(), n, k_lo, k_lo + len, tw.t.get(), tw.log_n, sh.k + 2, f_pow_u64(dom.extended_omega_inv, n), dom.post_e2c(), s);
        for (uint32_t q = 0; q < 3; ++q) shard.all_gather_inplace(h_coeff.get() + (size_t)q * n, len);
    } else {
        dev_extended_to_coeff(ctx, sh.k, h.get(), h_coeff.get());
    }
    h.release();
    lap(tm ? &tm->ntt : nullptr);
    skip_unused_blinds(3);
    for (const G1Affine& cm : commit_batch(ctx, 0, h_coeff.get(), n, 3, n)) tr.write_point(cm);
    lap(tm ? &tm->msm : nullptr);
    const Fr x = tr.squeeze_challenge();
    const Fr xn = f_pow_u64(x, n);
    shard.async_wait();  // the overlapped all-gathers of the z and advice coefficient forms
    // step 11: evaluations (D.10)
    const Fr x_next = dom.rotate_omega(x, 1), x_prev = dom.rotate_omega(x, -1), x_last = dom.rotate_omega(x, -(int)(bf + 1));
    const Fr x_rot2 = dom.rotate_omega(x, 2), x_rot3 = dom.rotate_omega(x, 3);
    // h_poly = sum_j xn^j h_j
    DevBuf<Fr> h_poly(n, s);
    fr_lincomb(h_poly.get(), {h_coeff.get(), h_coeff.get() + n, h_coeff.get() + 2 * n}, {one, xn, f_sqr(xn)}, n, false, s);
    // poly table for SHPLONK
    std::vector<const Fr*> polys;
    auto add_poly = [&](const Fr* p) {
        polys.push_back(p);
        return polys.size() - 1;
    };
    std::vector<size_t> id_adv(NA), id_fixed(sh.num_fixed()), id_sigma(P), id_z(NS), id_lz(L), id_la(L), id_ls(L);
    for (uint32_t c = 0; c < NA; ++c) id_adv[c] = add_poly(advice_polys.get() + (size_t)c * n);
    for (uint32_t i = 0; i < sh.num_fixed(); ++i) id_fixed[i] = add_poly(pk.fixed_polys.get() + (size_t)i * n);
    const size_t id_h = add_poly(h_poly.get()), id_rand = add_poly(random_poly.get());
    for (uint32_t j = 0; j < P; ++j) id_sigma[j] = add_poly(pk.sigma_polys.get() + (size_t)j * n);
    for (uint32_t set = 0; set < NS; ++set) id_z[set] = add_poly(z_polys.get() + (size_t)set * n);
    for (uint32_t l = 0; l < L; ++l) {
        id_lz[l] = add_poly(lk_z_poly.get() + (size_t)l * n);
        id_la[l] = add_poly(perm_in_poly.get() + (size_t)l * n);
        id_ls[l] = add_poly(perm_tab_poly.get() + (size_t)l * n);
    }
    // evaluate everything grouped by point
    std::vector<Fr> ev_x(polys.size()), ev_next(polys.size()), ev_r2(A), ev_r3(A), ev_prev(L), ev_last(NS);
    {
        eval_many_dist(ctx, shard, polys, n, x, ev_x.data());
        std::vector<const Fr*> list;
        std::vector<Fr> out;
        // x_next: gate advice columns, permutation z, lookup z
        for (uint32_t c = 0; c < A; ++c) list.push_back(polys[id_adv[c]]);
        for (uint32_t set = 0; set < NS; ++set) list.push_back(polys[id_z[set]]);
        for (uint32_t l = 0; l < L; ++l) list.push_back(polys[id_lz[l]]);
        out.resize(list.size());
        eval_many_dist(ctx, shard, list, n, x_next, out.data());
        for (uint32_t c = 0; c < A; ++c) ev_next[id_adv[c]] = out[c];
        for (uint32_t set = 0; set < NS; ++set) ev_next[id_z[set]] = out[A + set];
        for (uint32_t l = 0; l < L; ++l) ev_next[id_lz[l]] = out[A + NS + l];
        list.clear();
        for (uint32_t c = 0; c < A; ++c) list.push_back(polys[id_adv[c]]);
        eval_many_dist(ctx, shard, list, n, x_rot2, ev_r2.data());
        eval_many_dist(ctx, shard, list, n, x_rot3, ev_r3.data());
        list.clear();
        for (uint32_t l = 0; l < L; ++l) list.push_back(polys[id_la[l]]);
        eval_many_dist(ctx, shard, list, n, x_prev, ev_prev.data());
        list.clear();
        for (uint32_t set = 0; set + 1 < NS; ++set) list.push_back(polys[id_z[set]]);
        eval_many_dist(ctx, shard, list, n, x_last, ev_last.data());
    }
    std::vector<Query> q_advice, q_perm, q_lookup, q_fixed, q_sigma, q_vanish, queries;
    for (uint32_t c = 0; c < NA; ++c) {
        const size_t id = id_adv[c];
        tr.write_scalar(ev_x[id]);
        q_advice.push_back({id, x, ev_x[id]});
        if (c < A) {
            tr.write_scalar(ev_next[id]);
            tr.write_scalar(ev_r2[c]);
            tr.write_scalar(ev_r3[c]);
            q_advice.push_back({id, x_next, ev_next[id]});
            q_advice.push_back({id, x_rot2, ev_r2[c]});
            q_advice.push_back({id, x_rot3, ev_r3[c]});
        }
    }
    for (uint32_t i = 0; i < sh.num_fixed(); ++i) {
        tr.write_scalar(ev_x[id_fixed[i]]);
        q_fixed.push_back({id_fixed[i], x, ev_x[id_fixed[i]]});
    }
    tr.write_scalar(ev_x[id_rand]);
    q_vanish.push_back({id_h, x, ev_x[id_h]});
    q_vanish.push_back({id_rand, x, ev_x[id_rand]});
    for (uint32_t j = 0; j < P; ++j) {
        tr.write_scalar(ev_x[id_sigma[j]]);
        q_sigma.push_back({id_sigma[j], x, ev_x[id_sigma[j]]});
    }
    {
        std::vector<Query> lastq;
        for (uint32_t set = 0; set < NS; ++set) {
            const size_t id = id_z[set];
            tr.write_scalar(ev_x[id]);
            tr.write_scalar(ev_next[id]);
            q_perm.push_back({id, x, ev_x[id]});
            q_perm.push_back({id, x_next, ev_next[id]});
            if (set + 1 != NS) {
                tr.write_scalar(ev_last[set]);
                lastq.push_back({id, x_last, ev_last[set]});
            }
        }
        for (size_t i = lastq.size(); i-- > 0;) q_perm.push_back(lastq[i]);
    }
    for (uint32_t l = 0; l < L; ++l) {
        tr.write_scalar(ev_x[id_lz[l]]);
        tr.write_scalar(ev_next[id_lz[l]]);
        tr.write_scalar(ev_x[id_la[l]]);
        tr.write_scalar(ev_prev[l]);
        tr.write_scalar(ev_x[id_ls[l]]);
        q_lookup.push_back({id_lz[l], x, ev_x[id_lz[l]]});
        q_lookup.push_back({id_la[l], x, ev_x[id_la[l]]});
        q_lookup.push_back({id_ls[l], x, ev_x[id_ls[l]]});
        q_lookup.push_back({id_la[l], x_prev, ev_prev[l]});
        q_lookup.push_back({id_lz[l], x_next, ev_next[id_lz[l]]});
    }
    for (auto* v : {&q_advice, &q_perm, &q_lookup, &q_fixed, &q_sigma, &q_vanish}) queries.insert(queries.end(), v->begin(), v->end());
    lap(tm ? &tm->evals : nullptr);
    // step 12: SHPLONK (D.11)
    const Fr ych = tr.squeeze_challenge();
    std::vector<RotationSet> sets;
    std::vector<Fr> super;
    construct_intermediate_sets(queries, sets, super);
    const Fr vch = tr.squeeze_challenge();
    DevBuf<Fr> h_x(n, s), buf_a(n, s), buf_b(n, s);
    // single GPU: the y-combined polynomial of every rotation set (Σ_j y^j p_j) is kept, so that the final linear combination
    // runs over the sets instead of over all ≈ 75 polynomials again
    DevBuf<Fr> set_polys;
    const bool keep_sets = !shard.on();
    CUDA_CHECK(cudaMemsetAsync(h_x.get(), 0, n * sizeof(Fr), s));
    std::vector<std::vector<std::vector<Fr>>> low(sets.size());
    {
        // Sharded: the rotation sets are independent until they are summed into h(X), so they are dealt to the ranks by
        // cost (longest first onto the least loaded rank: one axpy per polynomial, a division ≈ 6 axpys per point); each rank
        // sums its sets into a partial h(X), the partials are all-gathered and added.
        std::vector<int> set_owner(sets.size(), 0);
        if (shard.on()) {
            std::vector<size_t> order(sets.size()), load(ctx.world, 0);
            for (size_t i = 0; i < sets.size(); ++i) order[i] = i;
            auto cost = [&](size_t i) { return sets[i].polys.size() + 6 * sets[i].points.size(); };
            std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return cost(a) > cost(b); });
            for (size_t i : order) {
                const size_t r = std::min_element(load.begin(), load.end()) - load.begin();
                set_owner[i] = (int)r;
                load[r] += cost(i);
            }
        }
        Fr vp = one;
        for (size_t i = 0; i < sets.size(); ++i) {
            const RotationSet& rs = sets[i];
            std::vector<const Fr*> ps;
            std::vector<Fr> cs;
            std::vector<Fr> r_comb(rs.points.size(), f_zero<FrCfg>());
            Fr yp = one;
            for (size_t j = 0; j < rs.polys.size(); ++j) {
                low[i].push_back(lagrange_interpolate(rs.points, rs.evals[j]));
                ps.push_back(polys[rs.polys[j]]);
                cs.push_back(yp);
                for (size_t t = 0; t < low[i][j].size(); ++t) r_comb[t] = f_add(r_comb[t], f_mul(low[i][j][t], yp));
                yp = f_mul(yp, ych);
            }
            if (!shard.on() || set_owner[i] == ctx.rank) {
                fr_lincomb(buf_a.get(), ps, cs, n, false, s);
                if (keep_sets) {
                    if (!set_polys.size()) set_polys.alloc(sets.size() * n, s);
                    CUDA_CHECK(cudaMemcpyAsync(set_polys.get() + i * n, buf_a.get(), n * sizeof(Fr), cudaMemcpyDeviceToDevice, s));
                }
                fr_sub_low(buf_a.get(), r_comb.data(), (uint32_t)r_comb.size(), s);
                Fr *src = buf_a.get(), *dst = buf_b.get();
                for (auto& pt : rs.points) {
                    fr_kate_division(ctx, src, dst, n, pt);
                    std::swap(src, dst);
                }
                fr_lincomb(h_x.get(), {src}, {vp}, n, true, s);
            }
            vp = f_mul(vp, vch);
        }
        if (shard.on()) {
            DevBuf<Fr> parts((size_t)ctx.world * n, s);
            CUDA_CHECK(cudaMemcpyAsync(parts.get() + (size_t)ctx.rank * n, h_x.get(), n * sizeof(Fr), cudaMemcpyDeviceToDevice, s));
            shard.all_gather_inplace(parts.get(), n);
            std::vector<const Fr*> ps(ctx.world);
            for (int r = 0; r < ctx.world; ++r) ps[r] = parts.get() + (size_t)r * n;
            fr_lincomb(h_x.get(), ps, std::vector<Fr>(ctx.world, one), n, false, s);
            CUDA_CHECK(cudaStreamSynchronize(s));
        }
    }
    lap(tm ? &tm->shplonk : nullptr);
    tr.write_point(commit_coeff(ctx, h_x.get(), n));
    lap(tm ? &tm->msm : nullptr);
    const Fr uch = tr.squeeze_challenge();
    {
        std::vector<const Fr*> ps;
        std::vector<Fr> cs;
        std::vector<Fr> z_diffs;
        Fr const_term = f_zero<FrCfg>(), vp = one;
        for (size_t i = 0; i < sets.size(); ++i) {
            const RotationSet& rs = sets[i];
            std::vector<Fr> diffs;
            for (auto& pnt : super) {
                bool in_set = false;
                for (auto& q : rs.points) in_set |= f_eq(pnt, q);
                if (!in_set) diffs.push_back(pnt);
            }
            const Fr z_i = vanishing_eval(diffs, uch);
            z_diffs.push_back(z_i);
            const Fr sc = f_mul(z_i, vp);
            Fr yp = one;
            if (keep_sets) {  // Σ_j (sc·y^j)·p_j = sc·(the set's kept y-combination): the same field elements
                ps.push_back(set_polys.get() + i * n);
                cs.push_back(sc);
            }
            for (size_t j = 0; j < rs.polys.size(); ++j) {
                if (!keep_sets) {
                    ps.push_back(polys[rs.polys[j]]);
                    cs.push_back(f_mul(sc, yp));
                }
                const_term = f_add(const_term, f_mul(f_mul(sc, yp), eval_small(low[i][j], uch)));
                yp = f_mul(yp, ych);
            }
            vp = f_mul(vp, vch);
        }
        const Fr zt_eval = vanishing_eval(super, uch);
        ps.push_back(h_x.get());
        cs.push_back(f_neg(zt_eval));
        if (shard.on() && n % ctx.world == 0) {  // the big linear combination by coefficient range, then an in-place all-gather
            const size_t len = n / ctx.world, lo = len * ctx.rank;
            for (auto& ptr : ps) ptr += lo;
            fr_lincomb(buf_a.get() + lo, ps, cs, len, false, s);
            shard.all_gather_inplace(buf_a.get(), len);
        } else {
            fr_lincomb(buf_a.get(), ps, cs, n, false, s);
        }
        fr_sub_low(buf_a.get(), &const_term, 1, s);
        fr_kate_division(ctx, buf_a.get(), buf_b.get(), n, uch);
        fr_scale(buf_b.get(), f_inv(z_diffs[0]), n, s);
    }
    lap(tm ? &tm->shplonk : nullptr);
    tr.write_point(commit_coeff(ctx, buf_b.get(), n));
    lap(tm ? &tm->msm : nullptr);
    if (tm) {  // the MSM stage includes the cross-rank exchange of partial sums: report it separately
        tm->other += ctx.exchange_seconds - exchange0;
        tm->msm -= ctx.exchange_seconds - exchange0;
        tm->comm = ctx.comm_seconds - comm0;
    }
    return tr.proof;
}

}  // namespace b200zk
