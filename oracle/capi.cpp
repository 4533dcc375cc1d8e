// ORACLE — TEST INFRASTRUCTURE ONLY (see field.hpp / plonk.hpp headers). C entry points for ctypes.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
#include <array>
#include <chrono>

#include "plonk.hpp"

using namespace oracle;

#define EXPORT extern "C" __attribute__((visibility("default")))

static std::string g_err;

EXPORT const char* oracle_last_error() { return g_err.c_str(); }
EXPORT void oracle_set_threads(int n) { num_threads_ref() = n > 0 ? n : (int)std::max(1u, std::thread::hardware_concurrency()); }
EXPORT int oracle_get_threads() { return num_threads(); }
// the [UNVERIFIED-1..4] switches (curve.hpp `Compat`): bit 0 = do NOT draw unused Blind scalars, bit 1 = ascending lookup
// fill, bit 2 = y-sign flag in bit 7; random_poly_chunks as in Compat. (0, 0) restores the defaults.
EXPORT void oracle_set_compat(uint32_t flags, uint32_t random_poly_chunks) {
    Compat& c = compat();
    c.draw_unused_blinds = !(flags & 1u);
    c.lookup_fill_from_end = !(flags & 2u);
    c.point_sign_bit = (flags & 4u) ? 7 : 6;
    c.random_poly_chunks = random_poly_chunks;
}

// ---- fields (which: 0 = Fr, 1 = Fq); all elements are 4×u64 Montgomery limbs ------------------------
template <class F>
static void field_op(int op, const u64* a, const u64* b, u64* out) {
    F x, y, r;
    memcpy(x.l, a, 32);
    if (b) memcpy(y.l, b, 32);
    switch (op) {
        case 0: r = x + y; break;
        case 1: r = x - y; break;
        case 2: r = x * y; break;
        case 3: r = x.inv(); break;
        case 4: r = -x; break;
        case 5: {  // to canonical
            U256 c = x.to_canonical();
            memcpy(out, c.l, 32);
            return;
        }
        case 6: {  // from canonical
            U256 c;
            memcpy(c.l, a, 32);
            r = F::from_canonical(c);
            break;
        }
        default: r = F::zero();
    }
    memcpy(out, r.l, 32);
}
EXPORT void oracle_field_op(int which, int op, const u64* a, const u64* b, u64* out) {
    if (which == 0) field_op<Fr>(op, a, b, out);
    else field_op<Fq>(op, a, b, out);
}
EXPORT void oracle_field_params(int which, u64* p, u64* inv, u64* r1, u64* r2, u64* r3) {
    const FieldParams& P = which == 0 ? Fr::P() : Fq::P();
    memcpy(p, P.p, 32);
    *inv = P.inv;
    memcpy(r1, P.r1, 32);
    memcpy(r2, P.r2, 32);
    memcpy(r3, P.r3, 32);
}
EXPORT void oracle_fr_from_u512(const u64* w, u64* out) {
    Fr r = Fr::from_u512(w);
    memcpy(out, r.l, 32);
}
EXPORT void oracle_fr_constants(u64* root, u64* zeta, u64* delta) {
    memcpy(root, FrConst::root_of_unity().l, 32);
    memcpy(zeta, FrConst::zeta().l, 32);
    memcpy(delta, FrConst::delta().l, 32);
}
EXPORT void oracle_fr_batch_invert(u64* a, size_t n) { batch_invert((Fr*)a, n); }

// ---- G1 ---------------------------------------------------------------------------------------------
EXPORT void oracle_g1_mul(const u64* p, const u64* s, u64* out) {
    G1Affine a;
    Fr k;
    memcpy(&a, p, 64);
    memcpy(k.l, s, 32);
    G1Affine r = G1::from_affine(a).mul(k).to_affine();
    memcpy(out, &r, 64);
}
EXPORT void oracle_g1_add(const u64* p, const u64* q, u64* out) {
    G1Affine a, b;
    memcpy(&a, p, 64);
    memcpy(&b, q, 64);
    G1Affine r = G1::from_affine(a).add(G1::from_affine(b)).to_affine();
    memcpy(out, &r, 64);
}
EXPORT void oracle_g1_madd(const u64* p, const u64* q, u64* out) {
    G1Affine a, b;
    memcpy(&a, p, 64);
    memcpy(&b, q, 64);
    G1Affine r = G1::from_affine(a).add_affine(b).to_affine();
    memcpy(out, &r, 64);
}
EXPORT int oracle_g1_on_curve(const u64* p) {
    G1Affine a;
    memcpy(&a, p, 64);
    return a.is_on_curve();
}
EXPORT void oracle_g1_to_bytes(const u64* p, uint8_t* out) {
    G1Affine a;
    memcpy(&a, p, 64);
    a.to_bytes(out);
}
EXPORT int oracle_g1_from_bytes(const uint8_t* in, u64* out) {
    G1Affine a;
    if (!G1Affine::from_bytes(in, a)) return 0;
    memcpy(out, &a, 64);
    return 1;
}
EXPORT void oracle_msm(const u64* scalars, const u64* bases, size_t n, u64* out) {
    G1Affine r = best_multiexp((const Fr*)scalars, (const G1Affine*)bases, n).to_affine();
    memcpy(out, &r, 64);
}
EXPORT void oracle_naive_msm(const u64* scalars, const u64* bases, size_t n, u64* out) {
    G1Affine r = naive_msm((const Fr*)scalars, (const G1Affine*)bases, n).to_affine();
    memcpy(out, &r, 64);
}

// ---- FFT / domain -----------------------------------------------------------------------------------
EXPORT void oracle_best_fft(u64* a, uint32_t log_n, const u64* omega) {
    Fr w;
    memcpy(w.l, omega, 32);
    best_fft((Fr*)a, w, log_n);
}
EXPORT void oracle_naive_dft(const u64* a, u64* out, size_t n, const u64* omega) {
    Fr w;
    memcpy(w.l, omega, 32);
    naive_dft((const Fr*)a, (Fr*)out, n, w);
}
// which: 0 omega, 1 omega_inv, 2 extended_omega, 3 extended_omega_inv, 4 ifft_divisor, 5 extended_ifft_divisor,
//        6.. t_evaluations[which-6] (inverted)
EXPORT void oracle_domain_constant(uint32_t k, int which, u64* out) {
    Domain d(Shape::degree, k);
    Fr v;
    switch (which) {
        case 0: v = d.omega; break;
        case 1: v = d.omega_inv; break;
        case 2: v = d.extended_omega; break;
        case 3: v = d.extended_omega_inv; break;
        case 4: v = d.ifft_divisor; break;
        case 5: v = d.extended_ifft_divisor; break;
        default: v = d.t_evaluations.at(which - 6);
    }
    memcpy(out, v.l, 32);
}
EXPORT void oracle_lagrange_to_coeff(uint32_t k, u64* a) {
    Domain d(Shape::degree, k);
    Poly p((Fr*)a, (Fr*)a + d.n);
    p = d.lagrange_to_coeff(std::move(p));
    memcpy(a, p.data(), 32 * d.n);
}
EXPORT void oracle_coeff_to_extended(uint32_t k, const u64* in, u64* out) {
    Domain d(Shape::degree, k);
    Poly p((const Fr*)in, (const Fr*)in + d.n);
    p = d.coeff_to_extended(std::move(p));
    memcpy(out, p.data(), 32 * d.extended_n);
}
EXPORT void oracle_extended_to_coeff(uint32_t k, const u64* in, u64* out) {
    Domain d(Shape::degree, k);
    Poly p((const Fr*)in, (const Fr*)in + d.extended_n);
    p = d.extended_to_coeff(std::move(p));
    memcpy(out, p.data(), 32 * p.size());
}
EXPORT void oracle_eval_polynomial(const u64* poly, size_t n, const u64* point, u64* out) {
    Fr x;
    memcpy(x.l, point, 32);
    Fr r = eval_polynomial((const Fr*)poly, n, x);
    memcpy(out, r.l, 32);
}
EXPORT void oracle_kate_division(const u64* a, size_t n, const u64* b, u64* out) {
    Fr x;
    memcpy(x.l, b, 32);
    auto q = kate_division((const Fr*)a, n, x);
    memcpy(out, q.data(), 32 * q.size());
}

// ---- RNG / hashing ----------------------------------------------------------------------------------
EXPORT void oracle_chacha_words(const uint8_t* seed, int rounds, size_t nwords, uint32_t* out) {
    ChaChaRng r(seed, rounds);
    for (size_t i = 0; i < nwords; ++i) out[i] = r.next_u32();
}
EXPORT void oracle_chacha_block(const uint32_t* in, uint32_t* out, int rounds) { chacha_block(in, out, rounds); }
EXPORT void oracle_std_rng_seed(u64 state, uint8_t* seed_out) {
    ChaChaRng r = ChaChaRng::std_rng_seed_from_u64(state);
    memcpy(seed_out, r.key, 32);
}
EXPORT void oracle_std_rng_random_fr(u64 state, size_t count, u64* out) {
    ChaChaRng r = ChaChaRng::std_rng_seed_from_u64(state);
    for (size_t i = 0; i < count; ++i) {
        Fr v = r.random_fr();
        memcpy(out + 4 * i, v.l, 32);
    }
}
EXPORT void oracle_blake2b(const char* personal16, const uint8_t* data, size_t len, uint8_t* out64) {
    Blake2b h(64, personal16);
    h.update(data, len);
    h.finalize(out64);
}
// absorbs: kind 0 = squeeze (writes 32-byte challenge to out, advancing by 32), 1 = point (64 B in), 2 = scalar (32 B in)
EXPORT size_t oracle_transcript_script(const uint8_t* kinds, size_t nops, const u64* in, u64* out, uint8_t* proof_out) {
    TranscriptWrite tr;
    size_t ip = 0, op = 0;
    for (size_t i = 0; i < nops; ++i) {
        if (kinds[i] == 0) {
            Fr c = tr.squeeze_challenge();
            memcpy(out + op, c.l, 32);
            op += 4;
        } else if (kinds[i] == 1) {
            G1Affine p;
            memcpy(&p, in + ip, 64);
            ip += 8;
            tr.write_point(p);
        } else {
            Fr s;
            memcpy(s.l, in + ip, 32);
            ip += 4;
            tr.write_scalar(s);
        }
    }
    memcpy(proof_out, tr.proof.data(), tr.proof.size());
    return tr.proof.size();
}

// ---- KZG params / keygen / prover / verifier (opaque handles) -----------------------------------------
EXPORT void* oracle_params_setup(uint32_t k, const uint8_t* chacha20_seed) {
    ChaChaRng rng = ChaChaRng::chacha20_from_seed(chacha20_seed);
    return new Params(Params::setup(k, rng));
}
EXPORT void* oracle_params_from_trapdoor(uint32_t k, const u64* s) {
    Fr t;
    memcpy(t.l, s, 32);
    return new Params(Params::from_trapdoor(k, t));
}
// load externally produced bases (e.g. downloaded from the device setup) with a known trapdoor
EXPORT void* oracle_params_load(uint32_t k, const u64* s, const u64* g, const u64* g_lagrange) {
    Params* p = new Params;
    p->k = k;
    p->n = (size_t)1 << k;
    memcpy(p->s.l, s, 32);
    p->g.assign((const G1Affine*)g, (const G1Affine*)g + p->n);
    p->g_lagrange.assign((const G1Affine*)g_lagrange, (const G1Affine*)g_lagrange + p->n);
    return p;
}
EXPORT void oracle_params_free(void* p) { delete (Params*)p; }
EXPORT void oracle_params_get(void* p, u64* s, u64* g, u64* g_lagrange) {
    Params* P = (Params*)p;
    if (s) memcpy(s, P->s.l, 32);
    if (g) memcpy(g, P->g.data(), 64 * P->n);
    if (g_lagrange) memcpy(g_lagrange, P->g_lagrange.data(), 64 * P->n);
}
EXPORT void oracle_lagrange_via_group_fft(void* p, u64* out) {
    Params* P = (Params*)p;
    auto r = Params::lagrange_via_group_fft(P->g, P->k);
    memcpy(out, r.data(), 64 * P->n);
}
EXPORT void oracle_commit(void* p, int lagrange, const u64* poly, u64* out) {
    Params* P = (Params*)p;
    Poly a((const Fr*)poly, (const Fr*)poly + P->n);
    G1Affine r = (lagrange ? P->commit_lagrange(a) : P->commit(a)).to_affine();
    memcpy(out, &r, 64);
}

static std::vector<Poly> unpack_cols(const u64* data, size_t ncols, size_t n) {
    std::vector<Poly> cols(ncols);
    for (size_t c = 0; c < ncols; ++c) cols[c].assign((const Fr*)data + c * n, (const Fr*)data + (c + 1) * n);
    return cols;
}

// fixed: num_fixed columns × n, column-major contiguous. copies: ncopies × 4 u32.
EXPORT void* oracle_keygen(void* params, uint32_t k, uint32_t A, uint32_t L, uint32_t F, const u64* fixed, const uint32_t* copies,
                           size_t ncopies) {
    try {
        Shape sh{k, A, L, F};
        std::vector<Copy> cp(ncopies);
        for (size_t i = 0; i < ncopies; ++i) cp[i] = Copy{copies[4 * i], copies[4 * i + 1], copies[4 * i + 2], copies[4 * i + 3]};
        return new ProvingKey(keygen(*(Params*)params, sh, unpack_cols(fixed, sh.num_fixed(), sh.n()), cp));
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
EXPORT void oracle_pk_free(void* pk) { delete (ProvingKey*)pk; }
// custom gates for this key (same encoding as b200zk_pk_set_gates: 7 uint32 per calculation); ncalcs = 0 clears them
EXPORT void oracle_pk_set_gates(void* pkp, const uint32_t* calcs, size_t ncalcs, const u64* constants, size_t nconstants, const uint32_t* results,
                                size_t nresults) {
    GateProgram& g = ((ProvingKey*)pkp)->vk.gates;
    g = GateProgram();
    for (size_t j = 0; j < ncalcs; ++j) {
        const uint32_t* w = calcs + 7 * j;
        g.calcs.push_back(GateCalculation{w[0], GateSource{w[1], w[2], (int32_t)w[3]}, GateSource{w[4], w[5], (int32_t)w[6]}});
    }
    for (size_t j = 0; j < nconstants; ++j) {
        Fr c;
        memcpy(c.l, constants + 4 * j, 32);
        g.constants.push_back(c);
    }
    g.results.assign(results, results + nresults);
}
EXPORT void oracle_pk_transcript_repr(void* pk, u64* out) { memcpy(out, ((ProvingKey*)pk)->vk.transcript_repr.l, 32); }
// which: 0 fixed commitments, 1 permutation commitments, 2 sigma values column j (n Fr), 3 fixed coset i, 4 sigma coset j,
//        5 l0, 6 l_last, 7 l_active_row
EXPORT void oracle_pk_get(void* pkp, int which, uint32_t idx, u64* out) {
    ProvingKey* pk = (ProvingKey*)pkp;
    auto cp = [&](const void* src, size_t bytes) { memcpy(out, src, bytes); };
    switch (which) {
        case 0: cp(pk->vk.fixed_commitments.data(), 64 * pk->vk.fixed_commitments.size()); break;
        case 1: cp(pk->vk.perm_commitments.data(), 64 * pk->vk.perm_commitments.size()); break;
        case 2: cp(pk->sigma_values[idx].data(), 32 * pk->domain.n); break;
        case 3: cp(pk->fixed_cosets[idx].data(), 32 * pk->domain.extended_n); break;
        case 4: cp(pk->sigma_cosets[idx].data(), 32 * pk->domain.extended_n); break;
        case 5: cp(pk->l0.data(), 32 * pk->domain.extended_n); break;
        case 6: cp(pk->l_last.data(), 32 * pk->domain.extended_n); break;
        case 7: cp(pk->l_active_row.data(), 32 * pk->domain.extended_n); break;
    }
}
EXPORT size_t oracle_proof_size(uint32_t k, uint32_t A, uint32_t L, uint32_t F) { return Shape{k, A, L, F}.proof_size(); }

// returns proof length (0 on error). rng = StdRng::seed_from_u64(rng_seed). `seconds_out` = wall time of create_proof.
EXPORT size_t oracle_create_proof(void* params, void* pkp, const u64* advice, u64 rng_seed, uint8_t* proof_out, double* seconds_out) {
    try {
        ProvingKey* pk = (ProvingKey*)pkp;
        const Shape& sh = pk->vk.shape;
        ChaChaRng rng = ChaChaRng::std_rng_seed_from_u64(rng_seed);
        auto cols = unpack_cols(advice, sh.num_advice(), sh.n());
        auto t0 = std::chrono::steady_clock::now();
        auto proof = create_proof(*(Params*)params, *pk, std::move(cols), rng);
        auto t1 = std::chrono::steady_clock::now();
        if (seconds_out) *seconds_out = std::chrono::duration<double>(t1 - t0).count();
        memcpy(proof_out, proof.data(), proof.size());
        return proof.size();
    } catch (const std::exception& e) {
        g_err = e.what();
        return 0;
    }
}
// evaluate_h on its own: coefficient-form inputs (advice (A+L)×n, permutation products num_sets×n, per lookup Z, a', s'
// = L×3×n), challenges y/beta/gamma; h_out = 4n extended-domain values divided by the vanishing polynomial.
EXPORT int oracle_evaluate_h(void* pkp, const u64* advice_coeff, const u64* z_coeff, const u64* lookup_coeff, const u64* y, const u64* beta,
                             const u64* gamma, u64* h_out) {
    try {
        ProvingKey* pk = (ProvingKey*)pkp;
        const Shape& sh = pk->vk.shape;
        const Domain& dom = pk->domain;
        const size_t n = sh.n();
        std::vector<Poly> advice_cosets, z_cosets, lz(sh.L), la(sh.L), ls(sh.L);
        for (auto& p : unpack_cols(advice_coeff, sh.num_advice(), n)) advice_cosets.push_back(dom.coeff_to_extended(p));
        for (auto& p : unpack_cols(z_coeff, sh.num_sets(), n)) z_cosets.push_back(dom.coeff_to_extended(p));
        if (sh.L) {
            auto lk = unpack_cols(lookup_coeff, 3 * sh.L, n);
            for (uint32_t l = 0; l < sh.L; ++l) {
                lz[l] = lk[3 * l];
                la[l] = lk[3 * l + 1];
                ls[l] = lk[3 * l + 2];
            }
        }
        Challenges ch{};
        memcpy(ch.y.l, y, 32);
        memcpy(ch.beta.l, beta, 32);
        memcpy(ch.gamma.l, gamma, 32);
        Poly h = evaluate_h(*pk, dom, advice_cosets, z_cosets, lz, la, ls, ch);
        memcpy(h_out, h.data(), h.size() * 32);
        return 1;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 0;
    }
}
// 1 = accepted
EXPORT int oracle_verify_proof(void* params, void* pkp, const uint8_t* proof, size_t len) {
    ProvingKey* pk = (ProvingKey*)pkp;
    g_err = verify_proof(*(Params*)params, pk->vk, proof, len);
    return g_err.empty();
}
// same with the real pairing equation (slower: two Miller loops and one final exponentiation)
EXPORT int oracle_verify_proof_pairing(void* params, void* pkp, const uint8_t* proof, size_t len) {
    ProvingKey* pk = (ProvingKey*)pkp;
    g_err = verify_proof(*(Params*)params, pk->vk, proof, len, nullptr, true);
    return g_err.empty();
}
// pairing KATs: returns a bit mask of the identities that hold:
//  1: e(aP, bQ) == e(P, Q)^(ab)   2: e(P, Q) != 1   4: e(P, Q)^r == 1   8: e(P1 + P2, Q) == e(P1, Q)·e(P2, Q)
//  16: e(P, Q1 + Q2) == e(P, Q1)·e(P, Q2)   32: s·g2 is on the twist
EXPORT int oracle_pairing_selfcheck(const u64* a_mont, const u64* b_mont) {
    Fr a, b;
    memcpy(a.l, a_mont, 32);
    memcpy(b.l, b_mont, 32);
    const G1Affine P = G1Affine::generator();
    const G2Affine Q = G2Affine::generator();
    const Fq12 e = pairing(P, Q);
    int mask = 0;
    const G1Affine aP = G1::from_affine(P).mul(a).to_affine();
    const G2Affine bQ = Q.mul(b);
    if (pairing(aP, bQ) == e.pow_fr(a * b)) mask |= 1;
    if (!(e == Fq12::one())) mask |= 2;
    {
        // r = -1 + ... : e^r == 1  <=>  e^(r-1) · e == 1
        Fr minus_one = -Fr::one();
        if (e.pow_fr(minus_one) * e == Fq12::one()) mask |= 4;
    }
    const G1Affine P2 = G1::from_affine(P).mul(Fr::from_u64(7)).to_affine();
    const G1Affine Psum = G1::from_affine(aP).add_affine(P2).to_affine();
    if (pairing(Psum, Q) == pairing(aP, Q) * pairing(P2, Q)) mask |= 8;
    const G2Affine Q2 = Q.mul(Fr::from_u64(11));
    if (pairing(P, bQ.add(Q2)) == pairing(P, bQ) * pairing(P, Q2)) mask |= 16;
    if (bQ.is_on_curve() && Q.is_on_curve()) mask |= 32;
    return mask;
}
// e(p0, q0)·e(p1, q1) == 1 for raw inputs: G1 = 8 u64 (x, y Montgomery limbs), G2 = 16 u64 (x.c0, x.c1, y.c0, y.c1), the
// RawBytes layout of a ParamsKZG file. Returns 1/0, or -1 when a point is off its curve.
EXPORT int oracle_pairing_product_is_one(const u64* p0, const u64* q0, const u64* p1, const u64* q1) {
    auto g1 = [](const u64* p) {
        G1Affine a;
        memcpy(a.x.l, p, 32);
        memcpy(a.y.l, p + 4, 32);
        return a;
    };
    auto g2 = [](const u64* q) {
        G2Affine a;
        memcpy(a.x.c0.l, q, 32);
        memcpy(a.x.c1.l, q + 4, 32);
        memcpy(a.y.c0.l, q + 8, 32);
        memcpy(a.y.c1.l, q + 12, 32);
        a.inf = false;
        return a;
    };
    const G1Affine a0 = g1(p0), a1 = g1(p1);
    const G2Affine b0 = g2(q0), b1 = g2(q1);
    if (!a0.is_on_curve() || !a1.is_on_curve() || !b0.is_on_curve() || !b1.is_on_curve()) return -1;
    return pairing_product_is_one(a0, b0, a1, b1) ? 1 : 0;
}
EXPORT int oracle_mock_check(uint32_t k, uint32_t A, uint32_t L, uint32_t F, const u64* fixed, const u64* advice, const uint32_t* copies,
                             size_t ncopies) {
    Shape sh{k, A, L, F};
    std::vector<Copy> cp(ncopies);
    for (size_t i = 0; i < ncopies; ++i) cp[i] = Copy{copies[4 * i], copies[4 * i + 1], copies[4 * i + 2], copies[4 * i + 3]};
    g_err = mock_check(sh, unpack_cols(fixed, sh.num_fixed(), sh.n()), unpack_cols(advice, sh.num_advice(), sh.n()), cp);
    return g_err.empty();
}
EXPORT int oracle_permute_expression_pair(uint32_t k, const u64* input, const u64* table, u64* a_out, u64* s_out) {
    Shape sh{k, 1, 1, 1};
    Poly in((const Fr*)input, (const Fr*)input + sh.n()), tab((const Fr*)table, (const Fr*)table + sh.n()), a, s;
    if (!permute_expression_pair(sh, in, tab, a, s)) return 0;
    memcpy(a_out, a.data(), 32 * a.size());
    memcpy(s_out, s.data(), 32 * s.size());
    return 1;
}

// ---- verifier-only handles: check a proof against externally produced vk commitments (no oracle keygen needed) -------
// params: trapdoor only (the verifier here never touches the G1 bases); vk: shape + commitments + transcript_repr.
EXPORT void* oracle_verifier_new(uint32_t k, uint32_t A, uint32_t L, uint32_t F, const u64* trapdoor, const u64* fixed_commitments,
                                 const u64* perm_commitments, const u64* transcript_repr) {
    struct Bundle {
        Params params;
        VerifyingKey vk;
    };
    Bundle* b = new Bundle;
    b->params.k = k;
    b->params.n = (size_t)1 << k;
    memcpy(b->params.s.l, trapdoor, 32);
    b->vk.shape = Shape{k, A, L, F};
    b->vk.fixed_commitments.assign((const G1Affine*)fixed_commitments, (const G1Affine*)fixed_commitments + b->vk.shape.num_fixed());
    b->vk.perm_commitments.assign((const G1Affine*)perm_commitments, (const G1Affine*)perm_commitments + b->vk.shape.num_perm());
    memcpy(b->vk.transcript_repr.l, transcript_repr, 32);
    return b;
}
EXPORT int oracle_verifier_verify(void* h, const uint8_t* proof, size_t len, int use_pairing) {
    struct Bundle {
        Params params;
        VerifyingKey vk;
    };
    Bundle* b = (Bundle*)h;
    g_err = verify_proof(b->params, b->vk, proof, len, nullptr, use_pairing != 0);
    return g_err.empty();
}
EXPORT void oracle_verifier_free(void* h) {
    struct Bundle {
        Params params;
        VerifyingKey vk;
    };
    delete (Bundle*)h;
}
