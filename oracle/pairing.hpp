// ORACLE — TEST INFRASTRUCTURE ONLY. BN254 optimal ate pairing, written for clarity not speed (≈20 ms per pairing),
// so that the oracle's verify_proof can check the SHPLONK opening with the real equation
//     e(L, [1]_2) · e(−H', [s]_2) == 1
// instead of the trapdoor shortcut — i.e. the check halo2_proofs::plonk::verify_proof performs inside halo2-base's
// `check_proof` (the only assertion of the reference's bench tests, verifier/src/stark/mod.rs:543,593).
// PARITY UNPINNED vs upstream (no pairing vector in the reference); pinned by bilinearity, non-degeneracy and
// e(P,Q)^r = 1 in tests/test_oracle_kat.py, and by agreement with the trapdoor check on every proof.
//
// Construction: Fq2 = Fq[u]/(u²+1); Fq12 as the flat ring Fq[w]/(w¹² − 18w⁶ + 82) (w⁶ = ξ = 9 + u); G2 on the D-type
// twist y² = x³ + 3/ξ with untwist (x, y) -> (x·w², y·w³). Miller loop over 6x+2 with affine steps on the twist (Fq2
// inversions) and sparse line values y_P − (λ·x_P)·w + (λ·x_T − y_T)·w³, the two Frobenius lines of the optimal ate
// pairing, and a final exponentiation by (p¹² − 1)/r as ONE plain square-and-multiply (constant derived from p and r).
#pragma once
#include "curve.hpp"

namespace oracle {

struct Fq2 {
    Fq c0, c1;
    static Fq2 zero() { return {Fq::zero(), Fq::zero()}; }
    static Fq2 one() { return {Fq::one(), Fq::zero()}; }
    bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    bool operator==(const Fq2& o) const { return c0 == o.c0 && c1 == o.c1; }
    Fq2 operator+(const Fq2& o) const { return {c0 + o.c0, c1 + o.c1}; }
    Fq2 operator-(const Fq2& o) const { return {c0 - o.c0, c1 - o.c1}; }
    Fq2 operator-() const { return {-c0, -c1}; }
    Fq2 operator*(const Fq2& o) const { return {c0 * o.c0 - c1 * o.c1, c0 * o.c1 + c1 * o.c0}; }
    Fq2 scale(const Fq& k) const { return {c0 * k, c1 * k}; }
    Fq2 sqr() const { return *this * *this; }
    Fq2 dbl() const { return *this + *this; }
    Fq2 conj() const { return {c0, -c1}; }
    Fq2 inv() const {
        Fq d = (c0.sqr() + c1.sqr()).inv();
        return {c0 * d, -(c1 * d)};
    }
    Fq2 pow_hex(const char* hex) const {  // exponent as a big-endian hex string
        Fq2 r = one();
        for (const char* p = hex; *p; ++p) {
            int v = (*p >= '0' && *p <= '9') ? *p - '0' : (*p >= 'a' && *p <= 'f') ? *p - 'a' + 10 : *p - 'A' + 10;
            for (int b = 3; b >= 0; --b) {
                r = r.sqr();
                if ((v >> b) & 1) r = r * *this;
            }
        }
        return r;
    }
};

struct G2Affine {
    Fq2 x, y;
    bool inf = false;
    static G2Affine generator() {  // EIP-197
        return {{Fq::from_hex("1800deef121f1e76426a00665e5c4479674322d4f75edadd46debd5cd992f6ed"),
                 Fq::from_hex("198e9393920d483a7260bfb731fb5d25f1aa493335a9e71297e485b7aef312c2")},
                {Fq::from_hex("12c85ea5db8c6deb4aab71808dcb408fe3d1e7690c43d37b4ce6cc0166fa7daa"),
                 Fq::from_hex("090689d0585ff075ec9e99ad690c3395bc4b313370b38ef355acdadcd122975b")},
                false};
    }
    static Fq2 twist_b() { return Fq2{Fq::from_u64(3), Fq::zero()} * Fq2{Fq::from_u64(9), Fq::one()}.inv(); }
    bool is_on_curve() const { return inf || y.sqr() == x.sqr() * x + twist_b(); }
    G2Affine neg() const { return {x, -y, inf}; }
    // affine addition / doubling; `slope` receives lambda (undefined when the result is the point at infinity)
    G2Affine add(const G2Affine& o, Fq2* slope = nullptr) const {
        if (inf) return o;
        if (o.inf) return *this;
        Fq2 lam;
        if (x == o.x) {
            if (!(y == o.y) || y.is_zero()) return {Fq2::zero(), Fq2::zero(), true};
            lam = (x.sqr().dbl() + x.sqr()) * y.dbl().inv();
        } else {
            lam = (o.y - y) * (o.x - x).inv();
        }
        if (slope) *slope = lam;
        Fq2 x3 = lam.sqr() - x - o.x;
        return {x3, lam * (x - x3) - y, false};
    }
    G2Affine mul(const Fr& s) const {
        U256 e = s.to_canonical();
        G2Affine r{Fq2::zero(), Fq2::zero(), true};
        for (int i = 255; i >= 0; --i) {
            r = r.add(r);
            if ((e.l[i / 64] >> (i % 64)) & 1) r = r.add(*this);
        }
        return r;
    }
};

struct Fq12 {
    Fq c[12];
    static Fq12 one() {
        Fq12 r;
        for (auto& x : r.c) x = Fq::zero();
        r.c[0] = Fq::one();
        return r;
    }
    bool operator==(const Fq12& o) const {
        for (int i = 0; i < 12; ++i)
            if (c[i] != o.c[i]) return false;
        return true;
    }
    Fq12 operator*(const Fq12& o) const {
        Fq t[23];
        for (auto& x : t) x = Fq::zero();
        for (int i = 0; i < 12; ++i) {
            if (c[i].is_zero()) continue;
            for (int j = 0; j < 12; ++j) t[i + j] += c[i] * o.c[j];
        }
        const Fq k18 = Fq::from_u64(18), k82 = Fq::from_u64(82);
        for (int i = 22; i >= 12; --i) {  // w^12 = 18 w^6 - 82
            t[i - 6] += t[i] * k18;
            t[i - 12] -= t[i] * k82;
        }
        Fq12 r;
        for (int i = 0; i < 12; ++i) r.c[i] = t[i];
        return r;
    }
    Fq12 sqr() const { return *this * *this; }
    Fq12 pow_hex(const char* hex) const {
        Fq12 r = one();
        for (const char* p = hex; *p; ++p) {
            int v = (*p >= '0' && *p <= '9') ? *p - '0' : (*p >= 'a' && *p <= 'f') ? *p - 'a' + 10 : *p - 'A' + 10;
            for (int b = 3; b >= 0; --b) {
                r = r.sqr();
                if ((v >> b) & 1) r = r * *this;
            }
        }
        return r;
    }
    Fq12 pow_fr(const Fr& s) const {
        U256 e = s.to_canonical();
        Fq12 r = one();
        for (int i = 255; i >= 0; --i) {
            r = r.sqr();
            if ((e.l[i / 64] >> (i % 64)) & 1) r = r * *this;
        }
        return r;
    }
    // adds (a + b·u)·w^k with u = w^6 − 9, k + 6 < 12
    void add_fq2_at(const Fq2& v, int k) {
        c[k] += v.c0 - v.c1 * Fq::from_u64(9);
        c[k + 6] += v.c1;
    }
};

inline const char* final_exp_hex() {
    return "2f4b6dc97020fddadf107d20bc842d43bf6369b1ff6a1c71015f3f7be2e1e30a73bb94fec0daf15466b2383a5d3ec3d15ad524d8f70c54efee1bd8c3b21377e5"
           "63a09a1b705887e72eceaddea3790364a61f676baaf977870e88d5c6c8fef0781361e443ae77f5b63a2a2264487f2940a8b1ddb3d15062cd0fb2015dfc666844"
           "9aed3cc48a82d0d602d268c7daab6a41294c0cc4ebe5664568dfc50e1648a45a4a1e3a5195846a3ed011a337a02088ec80e0ebae8755cfe107acf3aafb40494e"
           "406f804216bb10cf430b0f37856b42db8dc5514724ee93dfb10826f0dd4a0364b9580291d2cd65664814fde37ca80bb4ea44eacc5e641bbadf423f9a2cbf813b"
           "8d145da90029baee7ddadda71c7f3811c4105262945bba1668c3be69a3c230974d83561841d766f9c9d570bb7fbe04c7e8a6c3c760c0de81def35692da361102"
           "b6b9b2b918837fa97896e84abb40a4efb7e54523a486964b64ca86f120";
}

// line through T (slope lambda on the twist) evaluated at P in G1: y_P − (λ x_P)·w + (λ x_T − y_T)·w³
inline Fq12 line_value(const G2Affine& T, const Fq2& lam, const G1Affine& P) {
    Fq12 l;
    for (auto& x : l.c) x = Fq::zero();
    l.c[0] = P.y;
    l.add_fq2_at(-(lam.scale(P.x)), 1);
    l.add_fq2_at(lam * T.x - T.y, 3);
    return l;
}

// Miller loop of the optimal ate pairing (no final exponentiation)
inline Fq12 miller_loop(const G1Affine& P, const G2Affine& Q) {
    Fq12 f = Fq12::one();
    if (P.is_identity() || Q.inf) return f;
    static const Fq2 xi{Fq::from_u64(9), Fq::one()};
    static const Fq2 g12 = xi.pow_hex("10216f7ba065e00de81ac1e7808072c9dd2b2385cd7b438469602eb24829a9c2");   // xi^((p-1)/3)
    static const Fq2 g13 = xi.pow_hex("183227397098d014dc2822db40c0ac2ecbc0b548b438e5469e10460b6c3e7ea3");   // xi^((p-1)/2)
    static const Fq2 g22 = xi.pow_hex("30c96e8276995341dde2529566d9b5ee5592c705cbd1cacb7a4a8c966ece68456cd8a31d35b6b9818c55d8979dcee498cab57b9adf8eb00691c1d8b62747890");   // xi^((p^2-1)/3)
    static const Fq2 g23 = xi.pow_hex("492e25c3b1e5fce2ccd37be01a4690e5805c2a88b1bab031376fd2e1a6359c682344f4abd09216425280c4e36cb656e5301039684f560809daa2c5113aeb4d8");   // xi^((p^2-1)/2)
    const char* bits = "11001110101111001011100000011100110111110011101100011101110101000";  // 6x + 2
    G2Affine T = Q;
    for (const char* b = bits + 1; *b; ++b) {
        Fq2 lam;
        G2Affine T2 = T.add(T, &lam);
        f = f.sqr() * line_value(T, lam, P);
        T = T2;
        if (*b == '1') {
            G2Affine T3 = T.add(Q, &lam);
            f = f * line_value(T, lam, P);
            T = T3;
        }
    }
    const G2Affine Q1{Q.x.conj() * g12, Q.y.conj() * g13, false};   // pi(Q)
    const G2Affine Q2{Q.x * g22, -(Q.y * g23), false};              // -pi^2(Q)
    Fq2 lam;
    G2Affine T1 = T.add(Q1, &lam);
    f = f * line_value(T, lam, P);
    T = T1;
    T.add(Q2, &lam);
    f = f * line_value(T, lam, P);
    return f;
}
inline Fq12 final_exponentiation(const Fq12& f) { return f.pow_hex(final_exp_hex()); }
inline Fq12 pairing(const G1Affine& P, const G2Affine& Q) { return final_exponentiation(miller_loop(P, Q)); }
// prod e(P_i, Q_i) == 1
inline bool pairing_product_is_one(const G1Affine& p0, const G2Affine& q0, const G1Affine& p1, const G2Affine& q1) {
    return final_exponentiation(miller_loop(p0, q0) * miller_loop(p1, q1)) == Fq12::one();
}

}  // namespace oracle
