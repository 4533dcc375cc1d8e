// Context of libb200zk: one CUDA device, one stream, cached twiddle tables / domains / SRS / proving keys.
#pragma once
#include <functional>
#include <memory>

#include "common.cuh"

namespace b200zk {

// ---- host-side Fr helpers (the HD field code on its portable path) -------------------------------------
inline Fr fr_from_hex(const char* s) {
    Fr c = f_zero<FrCfg>();
    size_t n = strlen(s);
    for (size_t i = 0; i < n; ++i) {
        char ch = s[n - 1 - i];
        uint32_t v = (ch >= '0' && ch <= '9') ? ch - '0' : (ch >= 'a' && ch <= 'f') ? ch - 'a' + 10 : ch - 'A' + 10;
        c.l[i / 8] |= v << (4 * (i % 8));
    }
    return f_to_mont(c);
}
inline Fr fr_from_u64(uint64_t v) {
    Fr c = f_zero<FrCfg>();
    c.l[0] = (uint32_t)v;
    c.l[1] = (uint32_t)(v >> 32);
    return f_to_mont(c);
}
struct FrConsts {
    static constexpr uint32_t S = 28;
    static const Fr& root_of_unity() {
        static Fr v = fr_from_hex("03ddb9f5166d18b798865ea93dd31f743215cf6dd39329c8d34f1ed960c37c9c");
        return v;
    }
    static const Fr& zeta() {
        static Fr v = fr_from_hex("30644e72e131a029048b6e193fd84104cc37a73fec2bc5e9b8ca0b2d36636f23");
        return v;
    }
    static const Fr& delta() {
        static Fr v = fr_from_hex("09226b6e22c6f0ca64ec26aad4c86e715b5f898e5e963f25870e56bbe533e9a2");
        return v;
    }
    // primitive 2^log_n-th root of unity used by EvaluationDomain
    static Fr root(uint32_t log_n) {
        Fr w = root_of_unity();
        for (uint32_t i = log_n; i < S; ++i) w = f_sqr(w);
        return w;
    }
};

// EvaluationDomain::new(4, k) constants
struct Domain {
    uint32_t k = 0, extended_k = 0;
    size_t n = 0, extended_n = 0;
    Fr omega, omega_inv, extended_omega, extended_omega_inv, ifft_divisor, extended_ifft_divisor;
    Fr t_inv[4];
    // device constants: [0..3) into-coset factors {1, ζ, ζ²}; [3..6) lagrange_to_coeff post factors {1/n}×3;
    // [6..9) extended_to_coeff post factors {1/4n, ζ²/4n, ζ/4n}; [9..13) t_inv
    DevBuf<Fr> consts;
    const Fr* pre_coset() const { return consts.get(); }
    const Fr* post_l2c() const { return consts.get() + 3; }
    const Fr* post_e2c() const { return consts.get() + 6; }
    const Fr* t_inv_dev() const { return consts.get() + 9; }
    Fr rotate_omega(const Fr& x, int rotation) const {
        return rotation >= 0 ? f_mul(x, f_pow_u64(omega, (uint64_t)rotation)) : f_mul(x, f_pow_u64(omega_inv, (uint64_t)(-(int64_t)rotation)));
    }
};

struct Srs {
    uint32_t k = 0;
    size_t n = 0;
    size_t tab_lo = 0, tab_n = 0;  // point range covered by the window tables (this rank's MSM shard)
    DevBuf<G1Affine> g, g_lagrange;
    // precomputed window tables T[w][i] = 2^(tab_c·w)·P_i (msm.cu, merged-bucket mode); empty when disabled
    DevBuf<G1Affine> g_tab, gl_tab;
    uint32_t tab_c = 0;
    // multi-GPU: whether EVERY rank built the table of g / g_lagrange (-1 = not agreed yet). Table building depends on each
    // rank's free memory; ranks that disagreed would pick different MSM configurations and exchange mismatched sums.
    mutable int tab_agreed[2] = {-1, -1};
    // the verifier-side G2 points (g2, s·g2) as file bytes: RawBytes (256 B) after setup, or whatever a read file carried
    std::vector<uint8_t> g2_bytes;
    int g2_format = 0;
};
// all-gather of `bytes` from every rank into recv[world][bytes]; returns 0 on success (host buffers)
typedef int (*AllGatherFn)(void* user, const void* send, size_t bytes, void* recv);


// Upstream details that change proof BYTES and could not be checked against halo2-axiom's source in this container
// (SURVEY.md §8c items 1–4). One switch each, mirrored bit for bit by the oracle (oracle/curve.hpp `Compat`); the defaults
// are the classic PSE-halo2 behaviour. b200zk_set_compat flips them — the day the real source is consulted, a mismatch is
// a flag flip, not a rewrite.
struct Compat {
    bool draw_unused_blinds = true;    // [1] create_proof draws the Blind(..) scalars that KZG commitments ignore
    bool lookup_fill_from_end = true;  // [2] permute_expression_pair: ascending leftovers go to repeated rows popped from the END
    uint32_t random_poly_chunks = 0;   // [3] vanishing::commit: 0 = n sequential draws; T = T worker chunks, each filled from its own
                                       //     ChaCha20Rng whose 32-byte seed is drawn from the main stream
    int point_sign_bit = 6;            // [4] y-sign flag bit of compressed G1 points (6 or 7; the identity flag takes the other)
};

#ifndef B200ZK_MSM_SLOTS
#define B200ZK_MSM_SLOTS 4
#endif
constexpr int MSM_SLOTS = B200ZK_MSM_SLOTS;  // MSM columns in flight per commit batch
#ifndef B200ZK_STAGE_WORKERS
#define B200ZK_STAGE_WORKERS 6
#endif
constexpr int STAGE_WORKERS = B200ZK_STAGE_WORKERS, STAGE_SLOTS = 2;  // pinned staging of pageable uploads (upload.cuh)

struct Context {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t aux_streams[MSM_SLOTS - 1] = {};  // further MSM columns in flight (msm.cu)
    cudaEvent_t msm_events[MSM_SLOTS] = {}, msm_join[MSM_SLOTS - 1] = {}, msm_fork = nullptr;
    cudaStream_t copy_stream = nullptr;    // witness upload overlapped with the advice commitments (prover.cu)
    cudaEvent_t copy_fork = nullptr, copy_done = nullptr;
    // A commit batch may be issued while its columns are still arriving: before a column's first kernel is queued, the batch
    // calls column_gate(column index, the stream that will read it) — the gate blocks the HOST until the column's copy has
    // been queued and makes the stream wait for it (prover.cu: the witness upload). Empty = no gating.
    std::function<void(size_t, cudaStream_t)> column_gate;
    std::vector<cudaEvent_t> column_events;  // one per advice column, created on first use
    cudaEvent_t column_event(size_t c) {
        while (column_events.size() <= c) {
            cudaEvent_t e;
            CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            column_events.push_back(e);
        }
        return column_events[c];
    }
    // transforms that do not depend on the transcript (the advice columns' lagrange_to_coeff / coeff_to_extended) run on this
    // stream beside the commit batches: the MSM kernels leave ≈ 20 % of the multiplier pipe idle (latency-bound phases), which
    // the NTT kernels fill (prover.cu)
    cudaStream_t ntt_stream = nullptr;
    cudaEvent_t ntt_fork = nullptr, ntt_done = nullptr;
    DevBuf<Fr> side_scratch;
    cudaStream_t comm_stream = nullptr;    // multi-GPU: collectives whose result is needed late overlap the kernels in between
    cudaEvent_t comm_fork = nullptr, comm_done = nullptr;
    bool comm_pending = false;
    uint8_t* stage_buf = nullptr;          // pinned staging slots for pageable host buffers, allocated on first use
    cudaEvent_t stage_ev[STAGE_WORKERS * STAGE_SLOTS] = {};
    uint32_t* pinned_u32 = nullptr;        // pinned words for the entry-count read-backs (one per slot)
    std::mutex mu;
    std::string last_error;
    std::map<uint32_t, std::unique_ptr<TwiddleTable>> tables;  // standard roots, keyed by table log
    std::unique_ptr<TwiddleTable> custom_table;                // last non-standard omega
    std::map<uint32_t, std::unique_ptr<Domain>> domains;
    DevBuf<Fr> scratch;
    Arena arena;  // transient buffers of this context's stream
    std::unique_ptr<Srs> srs;
    // multi-GPU: this context is rank `rank` of `world`. The host's all-gather callback carries the NCCL
    // id once (collectives.cuh); without a communicator the MSM entry points exchange their partial sums through it
    int rank = 0, world = 1;
    AllGatherFn allgather = nullptr;
    void* allgather_user = nullptr;
    // with comm_trace set (the timings mode of create_proof) every collective is bracketed by stream synchronisations and
    // its wall time — transfer plus waiting for the slowest peer — is added to comm_seconds
    bool comm_trace = false;
    double comm_seconds = 0;
    unsigned comm_calls = 0;
    double exchange_seconds = 0;  // host time spent in the cross-rank exchange of partial MSM sums (reported under "other")
    bool msm_tables_enabled = true;
    int msm_affine_rounds = 0;  // batched-affine pre-reduction rounds of dense MSM columns (msm.cu; measured slower: off)
    std::shared_ptr<struct Nccl> nccl;  // collectives.cuh: created on first use from the bootstrap callback (one process per
                                        // GPU), or installed by b200zk_create_multi (one process, one thread per GPU)
    // this context is one rank of a multi-GPU job
    Compat compat;
    bool solo = false;  // set around calls that must run on this GPU alone although the context belongs to a device group
    bool sharded() const { return !solo && world > 1 && (allgather != nullptr || nccl != nullptr); }

    // table of the standard 2^t-th root with t >= log_n
    const TwiddleTable& std_table(uint32_t log_n) {
        for (auto& kv : tables)
            if (kv.first >= log_n) return *kv.second;
        auto t = std::make_unique<TwiddleTable>();
        t->log_n = log_n;
        t->omega = FrConsts::root(log_n);
        t->t.alloc_persistent(log_n == 0 ? 1 : (size_t)1 << (log_n - 1), stream);
        build_twiddle_table(t->t.get(), t->omega, log_n, stream);
        auto& ref = tables[log_n];
        ref = std::move(t);
        return *ref;
    }
    Fr* get_scratch(size_t n) {
        if (scratch.size() < n) {
            CUDA_CHECK(cudaStreamSynchronize(stream));
            scratch.alloc_persistent(n, stream);
        }
        return scratch.get();
    }
    Fr* get_side_scratch(size_t n) {  // scratch of the transforms on ntt_stream
        if (side_scratch.size() < n) {
            CUDA_CHECK(cudaStreamSynchronize(ntt_stream));
            side_scratch.alloc_persistent(n, ntt_stream);
        }
        return side_scratch.get();
    }
    const Domain& domain(uint32_t k) {
        auto it = domains.find(k);
        if (it != domains.end()) return *it->second;
        auto d = std::make_unique<Domain>();
        d->k = k;
        d->extended_k = k + 2;
        d->n = (size_t)1 << k;
        d->extended_n = (size_t)4 << k;
        d->extended_omega = FrConsts::root(k + 2);
        d->omega = f_sqr(f_sqr(d->extended_omega));
        d->omega_inv = f_inv(d->omega);
        d->extended_omega_inv = f_inv(d->extended_omega);
        d->ifft_divisor = f_inv(fr_from_u64(d->n));
        d->extended_ifft_divisor = f_inv(fr_from_u64(d->extended_n));
        const Fr zeta = FrConsts::zeta(), zeta2 = f_sqr(zeta), one = f_one<FrCfg>();
        const Fr orig = f_pow_u64(zeta, d->n), step = f_pow_u64(d->extended_omega, d->n);
        Fr cur = orig;
        for (int j = 0; j < 4; ++j) {
            d->t_inv[j] = f_inv(f_sub(cur, one));
            cur = f_mul(cur, step);
        }
        Fr h[13] = {one, zeta, zeta2,
                    d->ifft_divisor, d->ifft_divisor, d->ifft_divisor,
                    d->extended_ifft_divisor, f_mul(d->extended_ifft_divisor, zeta2), f_mul(d->extended_ifft_divisor, zeta),
                    d->t_inv[0], d->t_inv[1], d->t_inv[2], d->t_inv[3]};
        d->consts.alloc_persistent(13, stream);
        CUDA_CHECK(cudaMemcpyAsync(d->consts.get(), h, sizeof(h), cudaMemcpyHostToDevice, stream));
        CUDA_CHECK(cudaStreamSynchronize(stream));
        auto& ref = domains[k];
        ref = std::move(d);
        return *ref;
    }
};

// `batch` transforms of one plan; out may alias in; scratch must hold batch·2^log_n elements when more than one pass is needed
void ntt_run_batch(const NttPlan& plan, const Fr* in, Fr* out, Fr* scratch, uint32_t batch, size_t stride_in, size_t stride_out,
                   size_t stride_scratch, cudaStream_t stream);

// domain transforms on device buffers (ntt.cu semantics; see b200zk.h)
inline NttPlan make_plan(Context& ctx, uint32_t log_n, bool inverse) {
    const TwiddleTable& t = ctx.std_table(log_n);
    NttPlan p{};
    p.table = t.t.get();
    p.table_log = t.log_n;
    p.log_n = log_n;
    p.inverse = inverse;
    return p;
}
// `side`: run on the context's ntt_stream with its own scratch (see Context::ntt_stream) instead of the main stream
inline void dev_lagrange_to_coeff(Context& ctx, uint32_t k, Fr* a, uint32_t batch = 1, size_t stride = 0, bool side = false) {
    const Domain& d = ctx.domain(k);
    NttPlan p = make_plan(ctx, k, true);
    p.post_scale3 = d.post_l2c();
    Fr* scratch = ntt_num_passes(k) > 1 ? (side ? ctx.get_side_scratch(d.n * batch) : ctx.get_scratch(d.n * batch)) : nullptr;
    ntt_run_batch(p, a, a, scratch, batch, stride, stride, d.n, side ? ctx.ntt_stream : ctx.stream);
}
inline void dev_coeff_to_extended(Context& ctx, uint32_t k, const Fr* in, Fr* out, uint32_t batch = 1, size_t stride_in = 0,
                                  size_t stride_out = 0, bool side = false) {
    const Domain& d = ctx.domain(k);
    NttPlan p = make_plan(ctx, k + 2, false);
    p.pre_scale3 = d.pre_coset();
    p.in_len = d.n;
    // multi-pass transforms bounce through scratch; the output buffer itself can serve when batch == 1
    Fr* scratch = side ? ctx.get_side_scratch(d.extended_n * batch) : ctx.get_scratch(d.extended_n * batch);
    ntt_run_batch(p, in, out, scratch, batch, stride_in, stride_out, d.extended_n, side ? ctx.ntt_stream : ctx.stream);
}
inline void dev_extended_to_coeff(Context& ctx, uint32_t k, const Fr* in, Fr* out) {
    const Domain& d = ctx.domain(k);
    NttPlan p = make_plan(ctx, k + 2, true);
    p.post_scale3 = d.post_e2c();
    p.out_len = 3 * d.n;
    Fr* scratch = ctx.get_scratch(d.extended_n);
    ntt_run_batch(p, in, out, scratch, 1, 0, 0, 0, ctx.stream);
}

}  // namespace b200zk
