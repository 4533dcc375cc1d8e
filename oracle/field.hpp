// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product path.
// CPU restatement of halo2curves::bn256::{Fr,Fq} (crate halo2curves-axiom, un-vendored, floating
// pin — see SURVEY.md §8c). PARITY UNPINNED against upstream: the reference repo holds no golden
// vector for this path; this file is pinned instead by independent known-answer tests
// (tests/test_oracle_*.py: Python big-int arithmetic, re-derived constants of SURVEY.md §8a row A).
//
// Representation follows halo2curves: 4×u64 little-endian limbs, Montgomery form with R = 2^256,
// canonical (< modulus). Element type named by the reference at verifier/src/field/goldilocks/base.rs:472.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace oracle {

typedef uint64_t u64;
typedef unsigned __int128 u128;

struct U256 {
    u64 l[4];
};

inline bool u256_geq(const u64* a, const u64* b) {
    for (int i = 3; i >= 0; --i) {
        if (a[i] != b[i]) return a[i] > b[i];
    }
    return true;
}
inline u64 u256_add(u64* r, const u64* a, const u64* b) {
    u128 c = 0;
    for (int i = 0; i < 4; ++i) {
        c += (u128)a[i] + b[i];
        r[i] = (u64)c;
        c >>= 64;
    }
    return (u64)c;
}
inline u64 u256_sub(u64* r, const u64* a, const u64* b) {
    u64 borrow = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a[i] - b[i] - borrow;
        r[i] = (u64)d;
        borrow = (u64)(d >> 64) & 1;
    }
    return borrow;
}
inline U256 u256_from_hex(const char* s) {
    U256 r{};
    if (s[0] == '0' && (s[1] == 'x' || s[1] == 'X')) s += 2;
    size_t n = strlen(s);
    for (size_t i = 0; i < n; ++i) {
        char ch = s[n - 1 - i];
        u64 v = (ch >= '0' && ch <= '9') ? ch - '0' : (ch >= 'a' && ch <= 'f') ? ch - 'a' + 10 : ch - 'A' + 10;
        r.l[i / 16] |= v << (4 * (i % 16));
    }
    return r;
}

// Parameters derived at start-up from the modulus alone (R, R^2, R^3, -p^-1 mod 2^64); the hard
// coded upstream constants are checked against these in the tests.
struct FieldParams {
    u64 p[4];
    u64 inv;
    u64 r1[4], r2[4], r3[4];
    explicit FieldParams(const char* modulus_hex) {
        U256 m = u256_from_hex(modulus_hex);
        memcpy(p, m.l, 32);
        u64 x = 1;  // Newton iteration for p^-1 mod 2^64
        for (int i = 0; i < 6; ++i) x *= 2 - p[0] * x;
        inv = (u64)0 - x;
        u64 t[4] = {1, 0, 0, 0};
        for (int i = 0; i < 768; ++i) {
            u64 carry = u256_add(t, t, t);
            if (carry || u256_geq(t, p)) u256_sub(t, t, p);
            if (i == 255) memcpy(r1, t, 32);
            if (i == 511) memcpy(r2, t, 32);
            if (i == 767) memcpy(r3, t, 32);
        }
    }
};

struct FrTag {
    static const FieldParams& P() {
        static FieldParams fp("30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001");
        return fp;
    }
};
struct FqTag {
    static const FieldParams& P() {
        static FieldParams fp("30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47");
        return fp;
    }
};

template <class Tag>
struct Fp {
    u64 l[4];

    static const FieldParams& P() { return Tag::P(); }
    static Fp zero() { return Fp{{0, 0, 0, 0}}; }
    static Fp one() {
        Fp r;
        memcpy(r.l, P().r1, 32);
        return r;
    }
    bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
    bool operator==(const Fp& o) const { return memcmp(l, o.l, 32) == 0; }
    bool operator!=(const Fp& o) const { return !(*this == o); }

    Fp operator+(const Fp& o) const {
        Fp r;
        u64 c = u256_add(r.l, l, o.l);
        if (c || u256_geq(r.l, P().p)) u256_sub(r.l, r.l, P().p);
        return r;
    }
    Fp operator-(const Fp& o) const {
        Fp r;
        if (u256_sub(r.l, l, o.l)) u256_add(r.l, r.l, P().p);
        return r;
    }
    Fp operator-() const {
        if (is_zero()) return *this;
        Fp r;
        u256_sub(r.l, P().p, l);
        return r;
    }
    Fp dbl() const { return *this + *this; }

    // Montgomery reduction of an 8-limb product (the halo2curves `montgomery_reduce` shape).
    static Fp mont_reduce(u64 t[8]) {
        const u64* p = P().p;
        u64 carry2 = 0;
        for (int i = 0; i < 4; ++i) {
            u64 k = t[i] * P().inv;
            u128 c = 0;
            for (int j = 0; j < 4; ++j) {
                c += (u128)k * p[j] + t[i + j];
                t[i + j] = (u64)c;
                c >>= 64;
            }
            c += (u128)t[i + 4] + carry2;
            t[i + 4] = (u64)c;
            carry2 = (u64)(c >> 64);
        }
        Fp r;
        memcpy(r.l, t + 4, 32);
        if (carry2 || u256_geq(r.l, p)) u256_sub(r.l, r.l, p);
        return r;
    }
    Fp operator*(const Fp& o) const {
        u64 t[8] = {0};
        for (int i = 0; i < 4; ++i) {
            u128 c = 0;
            for (int j = 0; j < 4; ++j) {
                c += (u128)l[i] * o.l[j] + t[i + j];
                t[i + j] = (u64)c;
                c >>= 64;
            }
            t[i + 4] = (u64)c;
        }
        return mont_reduce(t);
    }
    Fp sqr() const { return *this * *this; }
    Fp& operator+=(const Fp& o) { return *this = *this + o; }
    Fp& operator-=(const Fp& o) { return *this = *this - o; }
    Fp& operator*=(const Fp& o) { return *this = *this * o; }

    // canonical (non-Montgomery) little-endian limbs == to_repr()
    U256 to_canonical() const {
        u64 t[8] = {l[0], l[1], l[2], l[3], 0, 0, 0, 0};
        Fp r = mont_reduce(t);
        U256 o;
        memcpy(o.l, r.l, 32);
        return o;
    }
    // value must be < p
    static Fp from_canonical(const U256& v) {
        Fp a, b;
        memcpy(a.l, v.l, 32);
        memcpy(b.l, P().r2, 32);
        return a * b;
    }
    static Fp from_u64(u64 v) { return from_canonical(U256{{v, 0, 0, 0}}); }
    static Fp from_hex(const char* s) { return from_canonical(u256_from_hex(s)); }
    // halo2curves `from_u512`: (lo + 2^256·hi) mod p, lo = limbs[0..4], hi = limbs[4..8]
    static Fp from_u512(const u64 w[8]) {
        Fp d0, d1, R2, R3;
        memcpy(d0.l, w, 32);
        memcpy(d1.l, w + 4, 32);
        memcpy(R2.l, P().r2, 32);
        memcpy(R3.l, P().r3, 32);
        return d0 * R2 + d1 * R3;  // mont mul accepts unreduced inputs < 2^256
    }
    void to_bytes(uint8_t out[32]) const {
        U256 c = to_canonical();
        memcpy(out, c.l, 32);  // little-endian host
    }
    // returns false when the encoding is not canonical
    static bool from_bytes(const uint8_t in[32], Fp& out) {
        U256 c;
        memcpy(c.l, in, 32);
        if (u256_geq(c.l, P().p)) return false;
        out = from_canonical(c);
        return true;
    }

    Fp pow(const u64* e, int nlimbs) const {
        Fp r = one();
        for (int i = nlimbs - 1; i >= 0; --i)
            for (int b = 63; b >= 0; --b) {
                r = r.sqr();
                if ((e[i] >> b) & 1) r = r * *this;
            }
        return r;
    }
    Fp pow_u64(u64 e) const { return pow(&e, 1); }
    // Fermat inversion; inverse of zero is zero (callers that care check first)
    Fp inv() const {
        u64 e[4];
        u64 two[4] = {2, 0, 0, 0};
        u256_sub(e, P().p, two);
        return pow(e, 4);
    }
    // numeric order of the canonical integer
    static int cmp(const Fp& a, const Fp& b) {
        U256 x = a.to_canonical(), y = b.to_canonical();
        for (int i = 3; i >= 0; --i)
            if (x.l[i] != y.l[i]) return x.l[i] < y.l[i] ? -1 : 1;
        return 0;
    }
};

typedef Fp<FrTag> Fr;
typedef Fp<FqTag> Fq;

// Montgomery-trick batch inversion; zeros are left as zero (halo2 `batch_invert` semantics).
template <class F>
inline void batch_invert(F* a, size_t n) {
    std::vector<F> pre(n);
    F acc = F::one();
    for (size_t i = 0; i < n; ++i) {
        pre[i] = acc;
        if (!a[i].is_zero()) acc = acc * a[i];
    }
    acc = acc.inv();
    for (size_t i = n; i-- > 0;) {
        if (a[i].is_zero()) continue;
        F t = a[i];
        a[i] = acc * pre[i];
        acc = acc * t;
    }
}

// Fr constants of halo2curves::bn256::Fr; verified algebraically in tests.
struct FrConst {
    static constexpr int S = 28;
    static Fr generator() { return Fr::from_u64(7); }
    static Fr root_of_unity() { return Fr::from_hex("03ddb9f5166d18b798865ea93dd31f743215cf6dd39329c8d34f1ed960c37c9c"); }
    static Fr zeta() { return Fr::from_hex("30644e72e131a029048b6e193fd84104cc37a73fec2bc5e9b8ca0b2d36636f23"); }
    static Fr delta() { return Fr::from_hex("09226b6e22c6f0ca64ec26aad4c86e715b5f898e5e963f25870e56bbe533e9a2"); }
};

}  // namespace oracle
