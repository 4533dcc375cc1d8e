// ORACLE — TEST INFRASTRUCTURE ONLY. CPU restatement of halo2curves::bn256::{G1Affine, G1}
// (y^2 = x^3 + 3 over Fq, generator (1,2), Jacobian projective arithmetic). PARITY UNPINNED vs upstream
// (un-vendored crate); pinned by KATs: 2·G (EIP-196 vector), r·G = ∞, on-curve checks.
#pragma once
#include "field.hpp"

namespace oracle {

// Upstream details that change proof BYTES and could not be checked against halo2-axiom's source in this container
// (SURVEY.md §8c items 1–4). One switch each, mirrored bit for bit by the product (csrc/context.cuh `Compat`,
// b200zk_set_compat); the defaults are the classic PSE-halo2 behaviour. Set through oracle_set_compat.
struct Compat {
    bool draw_unused_blinds = true;    // [1] create_proof draws the Blind(..) scalars that KZG commitments ignore
    bool lookup_fill_from_end = true;  // [2] permute_expression_pair: ascending leftovers go to repeated rows popped from the END
    uint32_t random_poly_chunks = 0;   // [3] vanishing::commit: 0 = n sequential draws; T = T worker chunks, each filled from its own
                                       //     ChaCha20Rng whose 32-byte seed is drawn from the main stream (thread-count dependent upstream)
    int point_sign_bit = 6;            // [4] y-sign flag bit of compressed G1 points (6 or 7)
};
inline Compat& compat() {
    static Compat c;
    return c;
}


struct G1Affine {
    Fq x, y;  // identity is (0,0) — halo2curves convention
    static G1Affine identity() { return G1Affine{Fq::zero(), Fq::zero()}; }
    static G1Affine generator() { return G1Affine{Fq::from_u64(1), Fq::from_u64(2)}; }
    bool is_identity() const { return x.is_zero() && y.is_zero(); }
    bool operator==(const G1Affine& o) const { return x == o.x && y == o.y; }
    bool is_on_curve() const {
        if (is_identity()) return true;
        return y.sqr() == x.sqr() * x + Fq::from_u64(3);
    }
    G1Affine neg() const { return G1Affine{x, -y}; }

    // GroupEncoding of halo2curves' `new_curve_impl!` (32 bytes: LE x, bit 6 of byte 31 = y is odd,
    // bit 7 of byte 31 = identity). SURVEY.md §8c unverified item 4 — isolated here.
    // [UNVERIFIED-4] flag bits of the 32-byte compressed form: y-sign in bit compat().point_sign_bit (6 by default) of
    // byte 31, identity flag in the other of bits 6/7
    void to_bytes(uint8_t out[32]) const {
        const int sb = compat().point_sign_bit, ib = sb == 6 ? 7 : 6;
        if (is_identity()) {
            memset(out, 0, 32);
            out[31] |= (uint8_t)(1u << ib);
            return;
        }
        x.to_bytes(out);
        uint8_t yb[32];
        y.to_bytes(yb);
        out[31] |= (uint8_t)((yb[0] & 1) << sb);
    }
    static bool from_bytes(const uint8_t in[32], G1Affine& out) {
        const int sb = compat().point_sign_bit, ib = sb == 6 ? 7 : 6;
        uint8_t tmp[32];
        memcpy(tmp, in, 32);
        bool is_inf = (tmp[31] >> ib) & 1;
        int ysign = (tmp[31] >> sb) & 1;
        tmp[31] &= 0x3f;
        Fq x;
        if (!Fq::from_bytes(tmp, x)) return false;
        if (is_inf) {
            if (!x.is_zero() || ysign) return false;
            out = identity();
            return true;
        }
        Fq rhs = x.sqr() * x + Fq::from_u64(3);
        // q ≡ 3 (mod 4): sqrt = rhs^((q+1)/4)
        u64 e[4];
        u64 one[4] = {1, 0, 0, 0};
        u256_add(e, Fq::P().p, one);
        for (int i = 0; i < 4; ++i) e[i] = (e[i] >> 2) | (i < 3 ? e[i + 1] << 62 : 0);
        Fq y = rhs.pow(e, 4);
        if (y.sqr() != rhs) return false;
        uint8_t yb[32];
        y.to_bytes(yb);
        if ((yb[0] & 1) != ysign) y = -y;
        out = G1Affine{x, y};
        return true;
    }
};

struct G1 {
    Fq x, y, z;  // Jacobian: (x/z^2, y/z^3); identity has z = 0
    static G1 identity() { return G1{Fq::zero(), Fq::one(), Fq::zero()}; }
    static G1 from_affine(const G1Affine& a) {
        if (a.is_identity()) return identity();
        return G1{a.x, a.y, Fq::one()};
    }
    bool is_identity() const { return z.is_zero(); }

    G1 dbl() const {
        if (is_identity()) return *this;
        // a = 0 doubling (dbl-2009-l)
        Fq a = x.sqr(), b = y.sqr(), c = b.sqr();
        Fq d = ((x + b).sqr() - a - c).dbl();
        Fq e = a.dbl() + a, f = e.sqr();
        G1 r;
        r.x = f - d.dbl();
        r.z = (y * z).dbl();
        r.y = e * (d - r.x) - c.dbl().dbl().dbl();
        return r;
    }
    G1 add(const G1& o) const {
        if (is_identity()) return o;
        if (o.is_identity()) return *this;
        Fq z1z1 = z.sqr(), z2z2 = o.z.sqr();
        Fq u1 = x * z2z2, u2 = o.x * z1z1;
        Fq s1 = y * z2z2 * o.z, s2 = o.y * z1z1 * z;
        if (u1 == u2) {
            if (s1 == s2) return dbl();
            return identity();
        }
        Fq h = u2 - u1, i = h.dbl().sqr(), j = h * i, rr = (s2 - s1).dbl(), v = u1 * i;
        G1 r;
        r.x = rr.sqr() - j - v.dbl();
        r.y = rr * (v - r.x) - (s1 * j).dbl();
        r.z = ((z + o.z).sqr() - z1z1 - z2z2) * h;
        return r;
    }
    G1 add_affine(const G1Affine& o) const {
        if (o.is_identity()) return *this;
        if (is_identity()) return from_affine(o);
        Fq z1z1 = z.sqr();
        Fq u2 = o.x * z1z1, s2 = o.y * z1z1 * z;
        if (x == u2) {
            if (y == s2) return dbl();
            return identity();
        }
        Fq h = u2 - x, hh = h.sqr(), i = hh.dbl().dbl(), j = h * i, rr = (s2 - y).dbl(), v = x * i;
        G1 r;
        r.x = rr.sqr() - j - v.dbl();
        r.y = rr * (v - r.x) - (y * j).dbl();
        r.z = (z + h).sqr() - z1z1 - hh;
        return r;
    }
    G1 neg() const { return G1{x, -y, z}; }
    G1Affine to_affine() const {
        if (is_identity()) return G1Affine::identity();
        Fq zi = z.inv(), zi2 = zi.sqr();
        return G1Affine{x * zi2, y * zi2 * zi};
    }
    // double-and-add over the canonical scalar, MSB first
    G1 mul(const Fr& s) const {
        U256 e = s.to_canonical();
        G1 r = identity();
        for (int i = 255; i >= 0; --i) {
            r = r.dbl();
            if ((e.l[i / 64] >> (i % 64)) & 1) r = r.add(*this);
        }
        return r;
    }
    bool eq(const G1& o) const {
        G1Affine a = to_affine(), b = o.to_affine();
        return a == b;
    }
};

// halo2curves `batch_normalize`: one inversion for the whole slice.
inline void batch_normalize(const G1* in, G1Affine* out, size_t n) {
    std::vector<Fq> zs(n);
    for (size_t i = 0; i < n; ++i) zs[i] = in[i].z;
    batch_invert(zs.data(), n);
    for (size_t i = 0; i < n; ++i) {
        if (in[i].is_identity()) {
            out[i] = G1Affine::identity();
            continue;
        }
        Fq zi2 = zs[i].sqr();
        out[i] = G1Affine{in[i].x * zi2, in[i].y * zi2 * zs[i]};
    }
}

}  // namespace oracle
