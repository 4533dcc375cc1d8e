// ORACLE — TEST INFRASTRUCTURE ONLY. CPU restatement of halo2_proofs::poly::{EvaluationDomain,
// kzg::commitment::ParamsKZG} (crate halo2-axiom, un-vendored; SURVEY.md §8a rows D, K and
// Appendix A.2, A.4). Reached from the reference through halo2-base `bench_builder`
// (verifier/src/stark/mod.rs:543,593). PARITY UNPINNED vs upstream; pinned by the KATs of SURVEY §8c
// (NTT∘iNTT = id, extended round trip, commit_lagrange(lagrange(p)) == commit(p), commitment == p(s)·G).
#pragma once
#include "arith.hpp"
#include "rng.hpp"

namespace oracle {

typedef std::vector<Fr> Poly;

struct Domain {
    uint32_t k, extended_k;
    size_t n, extended_n;
    Fr omega, omega_inv, extended_omega, extended_omega_inv;
    Fr g_coset, g_coset_inv;
    Fr ifft_divisor, extended_ifft_divisor;
    std::vector<Fr> t_evaluations;  // already inverted
    uint32_t quotient_poly_degree;

    // EvaluationDomain::new(j, k)
    Domain(uint32_t j, uint32_t k_) : k(k_) {
        quotient_poly_degree = j - 1;
        n = (size_t)1 << k;
        extended_k = k;
        while (((size_t)1 << extended_k) < n * quotient_poly_degree) ++extended_k;
        extended_n = (size_t)1 << extended_k;
        extended_omega = FrConst::root_of_unity();
        for (uint32_t i = extended_k; i < (uint32_t)FrConst::S; ++i) extended_omega = extended_omega.sqr();
        extended_omega_inv = extended_omega.inv();
        omega = extended_omega;
        for (uint32_t i = k; i < extended_k; ++i) omega = omega.sqr();
        omega_inv = omega.inv();
        g_coset = FrConst::zeta();
        g_coset_inv = g_coset.sqr();
        Fr orig = g_coset.pow_u64(n), step = extended_omega.pow_u64(n), cur = orig;
        do {
            t_evaluations.push_back(cur);
            cur *= step;
        } while (cur != orig);
        for (auto& t : t_evaluations) t -= Fr::one();
        batch_invert(t_evaluations.data(), t_evaluations.size());
        ifft_divisor = Fr::from_u64(n).inv();
        extended_ifft_divisor = Fr::from_u64(extended_n).inv();
    }
    Fr rotate_omega(const Fr& x, int rotation) const {
        return rotation >= 0 ? x * omega.pow_u64((u64)rotation) : x * omega_inv.pow_u64((u64)(-(int64_t)rotation));
    }
    void ifft(Fr* a, const Fr& w_inv, uint32_t log_n, const Fr& divisor) const {
        best_fft(a, w_inv, log_n);
        size_t len = (size_t)1 << log_n;
        parallel_chunks(len, [&](size_t b, size_t e, int) {
            for (size_t i = b; i < e; ++i) a[i] *= divisor;
        });
    }
    Poly lagrange_to_coeff(Poly a) const {
        ifft(a.data(), omega_inv, k, ifft_divisor);
        return a;
    }
    Poly coeff_to_lagrange(Poly a) const {
        best_fft(a.data(), omega, k);
        return a;
    }
    void distribute_powers_zeta(Poly& a, bool into_coset) const {
        Fr cp[2] = {into_coset ? g_coset : g_coset_inv, into_coset ? g_coset_inv : g_coset};
        parallel_chunks(a.size(), [&](size_t b, size_t e, int) {
            for (size_t i = b; i < e; ++i) {
                size_t m = i % 3;
                if (m) a[i] *= cp[m - 1];
            }
        });
    }
    Poly coeff_to_extended(Poly a) const {
        distribute_powers_zeta(a, true);
        a.resize(extended_n, Fr::zero());
        best_fft(a.data(), extended_omega, extended_k);
        return a;
    }
    Poly extended_to_coeff(Poly a) const {
        ifft(a.data(), extended_omega_inv, extended_k, extended_ifft_divisor);
        distribute_powers_zeta(a, false);
        a.resize(n * quotient_poly_degree);
        return a;
    }
    void divide_by_vanishing_poly(Poly& a) const {
        size_t m = t_evaluations.size();
        parallel_chunks(a.size(), [&](size_t b, size_t e, int) {
            for (size_t i = b; i < e; ++i) a[i] *= t_evaluations[i % m];
        });
    }
    // l_i_range(x, xn, rotations): l_i(x) = (x^n - 1) * omega^i / (n * (x - omega^i))
    std::vector<Fr> l_i_range(const Fr& x, const Fr& xn, int rot_begin, int rot_end_inclusive) const {
        std::vector<Fr> res;
        Fr common = (xn - Fr::one()) * ifft_divisor;
        for (int r = rot_begin; r <= rot_end_inclusive; ++r) {
            Fr w = rotate_omega(Fr::one(), r);
            res.push_back(common * w * (x - w).inv());
        }
        return res;
    }
};

// ParamsKZG<Bn256>: only the G1 side matters to the prover; the trapdoor is kept so that the oracle
// verifier can check openings without a pairing (e(A,[1]_2) == e(B,[s]_2)  <=>  A == s·B).
struct Params {
    uint32_t k;
    size_t n;
    std::vector<G1Affine> g, g_lagrange;
    Fr s;

    // fixed-base windowed multiplication table for the generator (8-bit windows)
    struct FixedBase {
        std::vector<G1Affine> table;  // [32][255]
        FixedBase() {
            std::vector<G1> t(32 * 255);
            G1 base = G1::from_affine(G1Affine::generator());
            for (int w = 0; w < 32; ++w) {
                G1 acc = base;
                for (int d = 1; d <= 255; ++d) {
                    t[w * 255 + d - 1] = acc;
                    acc = acc.add(base);
                }
                base = acc;  // 256 * previous base
            }
            table.resize(t.size());
            batch_normalize(t.data(), table.data(), t.size());
        }
        G1 mul(const Fr& s) const {
            U256 e = s.to_canonical();
            G1 acc = G1::identity();
            for (int w = 0; w < 32; ++w) {
                unsigned d = (e.l[w / 8] >> (8 * (w % 8))) & 0xff;
                if (d) acc = acc.add_affine(table[w * 255 + d - 1]);
            }
            return acc;
        }
    };

    // ParamsKZG::setup(k, rng): s = Fr::random(rng); g[i] = s^i G; g_lagrange[i] = L_i(s) G.
    // Upstream derives g_lagrange with a group inverse FFT of g; with s known the same group elements
    // are L_i(s)·G, L_i(s) = (s^n - 1) ω^i / (n (s - ω^i)) — checked equal at small k in the tests
    // (`setup_via_group_fft`).
    static Params setup(uint32_t k, ChaChaRng& rng) { return from_trapdoor(k, rng.random_fr()); }
    static Params from_trapdoor(uint32_t k, const Fr& s) {
        Params p;
        p.k = k;
        p.n = (size_t)1 << k;
        p.s = s;
        size_t n = p.n;
        Domain dom(2, k);
        std::vector<Fr> pw(n), lag(n);
        pw[0] = Fr::one();
        for (size_t i = 1; i < n; ++i) pw[i] = pw[i - 1] * s;
        Fr sn = pw[n - 1] * s;
        Fr common = (sn - Fr::one()) * dom.ifft_divisor;
        {
            Fr w = Fr::one();
            for (size_t i = 0; i < n; ++i) {
                lag[i] = s - w;
                w *= dom.omega;
            }
            batch_invert(lag.data(), n);
            w = Fr::one();
            for (size_t i = 0; i < n; ++i) {
                lag[i] = lag[i] * common * w;
                w *= dom.omega;
            }
        }
        static const FixedBase fb;
        std::vector<G1> gp(n), glp(n);
        parallel_chunks(n, [&](size_t b, size_t e, int) {
            for (size_t i = b; i < e; ++i) {
                gp[i] = fb.mul(pw[i]);
                glp[i] = fb.mul(lag[i]);
            }
        });
        p.g.resize(n);
        p.g_lagrange.resize(n);
        batch_normalize(gp.data(), p.g.data(), n);
        batch_normalize(glp.data(), p.g_lagrange.data(), n);
        return p;
    }
    // the upstream derivation of g_lagrange (group iFFT) — small k only, used as a KAT
    static std::vector<G1Affine> lagrange_via_group_fft(const std::vector<G1Affine>& g, uint32_t k) {
        size_t n = g.size();
        Domain dom(2, k);
        std::vector<G1> a(n);
        for (size_t i = 0; i < n; ++i) a[i] = G1::from_affine(g[i]);
        struct GW {
            G1 p;
            GW operator+(const GW& o) const { return GW{p.add(o.p)}; }
            GW operator-(const GW& o) const { return GW{p.add(o.p.neg())}; }
        };
        std::vector<GW> w(n);
        for (size_t i = 0; i < n; ++i) w[i].p = a[i];
        best_fft_generic(w.data(), n, dom.omega_inv, k, [](const GW& x, const Fr& t) { return GW{x.p.mul(t)}; });
        std::vector<G1> out(n);
        for (size_t i = 0; i < n; ++i) out[i] = w[i].p.mul(dom.ifft_divisor);
        std::vector<G1Affine> res(n);
        batch_normalize(out.data(), res.data(), n);
        return res;
    }
    G1 commit(const Poly& coeffs) const { return best_multiexp(coeffs.data(), g.data(), coeffs.size()); }
    G1 commit_lagrange(const Poly& evals) const { return best_multiexp(evals.data(), g_lagrange.data(), evals.size()); }
};

}  // namespace oracle
