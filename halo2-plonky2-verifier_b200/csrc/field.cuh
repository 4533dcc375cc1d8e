// BN254 Fr / Fq Montgomery arithmetic for sm_100a (replaces
// halo2curves::bn256::{Fr,Fq} — the element type the reference names at
// verifier/src/field/goldilocks/base.rs:472).
//
// Memory layout is the halo2curves one (4×u64 LE limbs, Montgomery R = 2^256, canonical) viewed as
// 8×u32 — no conversion at the C-ABI boundary.
//
// Device multiply: word-serial Montgomery (CIOS) over 32-bit limbs with the accumulator split in two
// carry-save halves — products a[j]·x for even j land in E, odd j in O (one limb higher) — so every row is
// two independent mad.lo.cc/madc.hi.cc chains of four wide multiply-adds that ptxas fuses into
// IMAD.WIDE.U32(.X) carry chains: 16 wide IMADs + 1 IMAD per limb of b, 136 per product. Retiring one limb per
// row swaps the roles of E and O; the one-limb merge carry is fed into the next row's first chain.
// Device square: product scanning with 36 instead of 64 partial products (f_sqr_comba). Three other multipliers were
// measured and are kept as tested alternatives: product scanning with predicate carries (f_mul_comba), 29-bit unsaturated
// limbs (f_mul_u29) — both slower than the carry chains on B200, see the comments at their definitions.
// The host bodies of the same row primitives (explicit carry variable) are what the CPU tests exercise and
// what the host side of the prover uses for its handful of scalar operations.
#pragma once
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#define DEV __device__ __forceinline__
#else
#define HD inline
#define DEV inline
#endif

namespace b200zk {

struct FrCfg {
    // r = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
    static HD constexpr uint32_t P(int i) {
        constexpr uint32_t t[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return t[i];
    }
    static constexpr uint32_t INV = 0xefffffffu;  // -r^-1 mod 2^32
    static HD constexpr uint32_t R1(int i) {
        constexpr uint32_t t[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return t[i];
    }
    static HD constexpr uint32_t R2(int i) {
        constexpr uint32_t t[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
        return t[i];
    }
    static HD constexpr uint32_t R3(int i) {
        constexpr uint32_t t[8] = {0xb4bf0040u, 0x5e94d8e1u, 0x1cfbb6b8u, 0x2a489cbeu, 0xa19fcfedu, 0x893cc664u, 0x7fcc657cu, 0x0cf8594bu};
        return t[i];
    }
};
struct FqCfg {
    // q = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
    static HD constexpr uint32_t P(int i) {
        constexpr uint32_t t[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return t[i];
    }
    static constexpr uint32_t INV = 0xe4866389u;
    static HD constexpr uint32_t R1(int i) {
        constexpr uint32_t t[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return t[i];
    }
    static HD constexpr uint32_t R2(int i) {
        constexpr uint32_t t[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u, 0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
        return t[i];
    }
    static HD constexpr uint32_t R3(int i) {
        constexpr uint32_t t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        return t[i];
    }  // unused for Fq
};

template <class C>
struct alignas(16) Field {
    uint32_t l[8];
};
typedef Field<FrCfg> Fr;
typedef Field<FqCfg> Fq;

// ------------------------------------------------------------------------------------------------
// row primitives: device = PTX carry chains, host = the same arithmetic with an explicit carry.
// Every chain is ONE asm statement, so the carry flag never crosses a statement boundary. Read-write operands
// are early-clobber ("+&r"): they are written before the last input is read, and without '&' the compiler may
// place an input that happens to hold the same value (e.g. a zero) in the same register.
// ------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
// acc[0..7] += (v0,v2,v4,v6)·x ; top = top_in + carry
DEV void chain_mad8(uint32_t* acc, uint32_t& top, uint32_t top_in, uint32_t v0, uint32_t v2, uint32_t v4, uint32_t v6, uint32_t x) {
    asm("mad.lo.cc.u32 %0, %10, %14, %0;\n\t"
        "madc.hi.cc.u32 %1, %10, %14, %1;\n\t"
        "madc.lo.cc.u32 %2, %11, %14, %2;\n\t"
        "madc.hi.cc.u32 %3, %11, %14, %3;\n\t"
        "madc.lo.cc.u32 %4, %12, %14, %4;\n\t"
        "madc.hi.cc.u32 %5, %12, %14, %5;\n\t"
        "madc.lo.cc.u32 %6, %13, %14, %6;\n\t"
        "madc.hi.cc.u32 %7, %13, %14, %7;\n\t"
        "addc.u32 %8, %9, 0;"
        : "+&r"(acc[0]), "+&r"(acc[1]), "+&r"(acc[2]), "+&r"(acc[3]), "+&r"(acc[4]), "+&r"(acc[5]), "+&r"(acc[6]), "+&r"(acc[7]), "=r"(top)
        : "r"(top_in), "r"(v0), "r"(v2), "r"(v4), "r"(v6), "r"(x));
}
// same without a carry-out limb (caller guarantees no overflow past acc[7])
DEV void chain_mad8_nocarry(uint32_t* acc, uint32_t v0, uint32_t v2, uint32_t v4, uint32_t v6, uint32_t x) {
    asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32 %7, %11, %12, %7;"
        : "+&r"(acc[0]), "+&r"(acc[1]), "+&r"(acc[2]), "+&r"(acc[3]), "+&r"(acc[4]), "+&r"(acc[5]), "+&r"(acc[6]), "+&r"(acc[7])
        : "r"(v0), "r"(v2), "r"(v4), "r"(v6), "r"(x));
}
// merge + chain: e0 += m1 (carry c); then acc[0..6] += (v1,v3,v5,v7)·x + c with acc[7] = hi(v7·x) + carry (acc[7] was 0)
DEV void chain_merge_mad8(uint32_t& e0, uint32_t m1, uint32_t* acc, uint32_t v1, uint32_t v3, uint32_t v5, uint32_t v7, uint32_t x) {
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "madc.lo.cc.u32 %1, %10, %14, %1;\n\t"
        "madc.hi.cc.u32 %2, %10, %14, %2;\n\t"
        "madc.lo.cc.u32 %3, %11, %14, %3;\n\t"
        "madc.hi.cc.u32 %4, %11, %14, %4;\n\t"
        "madc.lo.cc.u32 %5, %12, %14, %5;\n\t"
        "madc.hi.cc.u32 %6, %12, %14, %6;\n\t"
        "madc.lo.cc.u32 %7, %13, %14, %7;\n\t"
        "madc.hi.u32 %8, %13, %14, 0;"
        : "+&r"(e0), "+&r"(acc[0]), "+&r"(acc[1]), "+&r"(acc[2]), "+&r"(acc[3]), "+&r"(acc[4]), "+&r"(acc[5]), "+&r"(acc[6]), "=r"(acc[7])
        : "r"(m1), "r"(v1), "r"(v3), "r"(v5), "r"(v7), "r"(x));
}
// r[0..7] = a[0..7] + b[0..7], returns carry. Outputs are tied to the a-operands ("+r") so that no output
// register can alias a later-read input.
DEV uint32_t add8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t c;
    uint32_t t0 = a[0], t1 = a[1], t2 = a[2], t3 = a[3], t4 = a[4], t5 = a[5], t6 = a[6], t7 = a[7];
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, %10;\n\t"
        "addc.cc.u32 %2, %2, %11;\n\t"
        "addc.cc.u32 %3, %3, %12;\n\t"
        "addc.cc.u32 %4, %4, %13;\n\t"
        "addc.cc.u32 %5, %5, %14;\n\t"
        "addc.cc.u32 %6, %6, %15;\n\t"
        "addc.cc.u32 %7, %7, %16;\n\t"
        "addc.u32 %8, 0, 0;"
        : "+&r"(t0), "+&r"(t1), "+&r"(t2), "+&r"(t3), "+&r"(t4), "+&r"(t5), "+&r"(t6), "+&r"(t7), "=r"(c)
        : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    r[0] = t0; r[1] = t1; r[2] = t2; r[3] = t3; r[4] = t4; r[5] = t5; r[6] = t6; r[7] = t7;
    return c;
}
// r = a - b, returns borrow (1 when a < b)
DEV uint32_t sub8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t c;
    uint32_t t0 = a[0], t1 = a[1], t2 = a[2], t3 = a[3], t4 = a[4], t5 = a[5], t6 = a[6], t7 = a[7];
    asm("sub.cc.u32 %0, %0, %9;\n\t"
        "subc.cc.u32 %1, %1, %10;\n\t"
        "subc.cc.u32 %2, %2, %11;\n\t"
        "subc.cc.u32 %3, %3, %12;\n\t"
        "subc.cc.u32 %4, %4, %13;\n\t"
        "subc.cc.u32 %5, %5, %14;\n\t"
        "subc.cc.u32 %6, %6, %15;\n\t"
        "subc.cc.u32 %7, %7, %16;\n\t"
        "subc.u32 %8, 0, 0;"
        : "+&r"(t0), "+&r"(t1), "+&r"(t2), "+&r"(t3), "+&r"(t4), "+&r"(t5), "+&r"(t6), "+&r"(t7), "=r"(c)
        : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    r[0] = t0; r[1] = t1; r[2] = t2; r[3] = t3; r[4] = t4; r[5] = t5; r[6] = t6; r[7] = t7;
    return c & 1u;
}
#else
// ---- host bodies of the same primitives ----
inline void chain_mad8(uint32_t* acc, uint32_t& top, uint32_t top_in, uint32_t v0, uint32_t v2, uint32_t v4, uint32_t v6, uint32_t x) {
    const uint32_t v[4] = {v0, v2, v4, v6};
    uint64_t carry = 0;
    for (int t = 0; t < 4; ++t) {
        uint64_t prod = (uint64_t)v[t] * x;
        uint64_t lo = (uint64_t)acc[2 * t] + (uint32_t)prod + carry;
        acc[2 * t] = (uint32_t)lo;
        uint64_t hi = (uint64_t)acc[2 * t + 1] + (prod >> 32) + (lo >> 32);
        acc[2 * t + 1] = (uint32_t)hi;
        carry = hi >> 32;
    }
    top = top_in + (uint32_t)carry;
}
inline void chain_mad8_nocarry(uint32_t* acc, uint32_t v0, uint32_t v2, uint32_t v4, uint32_t v6, uint32_t x) {
    uint32_t top;
    chain_mad8(acc, top, 0, v0, v2, v4, v6, x);
}
inline void chain_merge_mad8(uint32_t& e0, uint32_t m1, uint32_t* acc, uint32_t v1, uint32_t v3, uint32_t v5, uint32_t v7, uint32_t x) {
    uint64_t s = (uint64_t)e0 + m1;
    e0 = (uint32_t)s;
    uint64_t carry = s >> 32;
    const uint32_t v[4] = {v1, v3, v5, v7};
    acc[7] = 0;
    for (int t = 0; t < 4; ++t) {
        uint64_t prod = (uint64_t)v[t] * x;
        uint64_t lo = (uint64_t)acc[2 * t] + (uint32_t)prod + carry;
        acc[2 * t] = (uint32_t)lo;
        uint64_t hi = (uint64_t)acc[2 * t + 1] + (prod >> 32) + (lo >> 32);
        acc[2 * t + 1] = (uint32_t)hi;
        carry = hi >> 32;
    }
}
inline uint32_t add8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint64_t c = 0;
    for (int i = 0; i < 8; ++i) {
        c += (uint64_t)a[i] + b[i];
        r[i] = (uint32_t)c;
        c >>= 32;
    }
    return (uint32_t)c;
}
inline uint32_t sub8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint64_t br = 0;
    for (int i = 0; i < 8; ++i) {
        uint64_t d = (uint64_t)a[i] - b[i] - br;
        r[i] = (uint32_t)d;
        br = (d >> 32) & 1;
    }
    return (uint32_t)br;
}
#endif

// ------------------------------------------------------------------------------------------------
template <class C>
HD Field<C> f_zero() {
    Field<C> r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = 0;
    return r;
}
template <class C>
HD Field<C> f_one() {
    Field<C> r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = C::R1(i);
    return r;
}
template <class C>
HD bool f_is_zero(const Field<C>& a) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) o |= a.l[i];
    return o == 0;
}
template <class C>
HD bool f_eq(const Field<C>& a, const Field<C>& b) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) o |= a.l[i] ^ b.l[i];
    return o == 0;
}
// r = (t >= P) ? t - P : t ; `carry` = a 2^256 overflow bit of t
template <class C>
HD Field<C> f_reduce_once(const Field<C>& t, uint32_t carry) {
    uint32_t p[8], d[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = C::P(i);
    uint32_t borrow = sub8(d, t.l, p);
    bool use_d = carry | (borrow ^ 1u);
    Field<C> r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = use_d ? d[i] : t.l[i];
    return r;
}
template <class C>
HD Field<C> f_add(const Field<C>& a, const Field<C>& b) {
    Field<C> t;
    uint32_t c = add8(t.l, a.l, b.l);
    return f_reduce_once<C>(t, c);
}
template <class C>
HD Field<C> f_sub(const Field<C>& a, const Field<C>& b) {
    Field<C> t, u;
    uint32_t borrow = sub8(t.l, a.l, b.l);
    uint32_t p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = C::P(i);
    add8(u.l, t.l, p);
#pragma unroll
    for (int i = 0; i < 8; ++i) t.l[i] = borrow ? u.l[i] : t.l[i];
    return t;
}
template <class C>
HD Field<C> f_neg(const Field<C>& a) {
    if (f_is_zero(a)) return a;
    Field<C> r;
    uint32_t p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = C::P(i);
    sub8(r.l, p, a.l);
    return r;
}
template <class C>
HD Field<C> f_dbl(const Field<C>& a) {
    return f_add<C>(a, a);
}

// Montgomery product a·b·2^-256 mod P, canonical result. Precondition: a < P (the operand that is multiplied
// whole in every row — it bounds the running sum below 2P); b may be any value < 2^256.
template <class C>
HD Field<C> f_mul_chains(const Field<C>& a, const Field<C>& b) {
    // buf[0] and buf[1] hold the two carry-save halves; the live window slides up as limbs retire.
    uint32_t X[20], Y[20];
#pragma unroll
    for (int i = 0; i < 20; ++i) X[i] = Y[i] = 0;
    uint32_t* E = X;  // even-aligned half, 9 limbs live: E[0..8]
    uint32_t* O = Y;  // odd-aligned half (one limb higher), 8 limbs live: O[0..7]
    uint32_t p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = C::P(i);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t bi = b.l[i];
        if (i == 0) {
            chain_mad8_nocarry(O, a.l[1], a.l[3], a.l[5], a.l[7], bi);
            chain_mad8(E, E[8], 0, a.l[0], a.l[2], a.l[4], a.l[6], bi);
        } else {
            // retire limb 0 of the previous row: new E := old O (+ old E[1] at limb 0), new O := old E >> 64
            uint32_t* nE = O;
            uint32_t* nO = E + 2;
            chain_merge_mad8(nE[0], E[1], nO, a.l[1], a.l[3], a.l[5], a.l[7], bi);
            chain_mad8(nE, nE[8], 0, a.l[0], a.l[2], a.l[4], a.l[6], bi);
            E = nE;
            O = nO;
        }
        const uint32_t m = E[0] * C::INV;
        chain_mad8(E, E[8], E[8], p[0], p[2], p[4], p[6], m);
        chain_mad8_nocarry(O, p[1], p[3], p[5], p[7], m);
    }
    // E[0] == 0 now; result = (E >> 32) + O
    Field<C> t;
    uint32_t c = add8(t.l, E + 1, O);
    return f_reduce_once<C>(t, c);
}
// ---- product scanning (Comba) with predicate carries ------------------------------------------------------------------------
// Every partial product is added into a three-word column accumulator (t0,t1,t2) with an IMAD.WIDE.U32 that writes its
// carry-OUT to a predicate and takes no carry-IN; the carries are counted into t2 by IADD3.X on the ALU pipe (ptxas folds
// two carry predicates into one IADD3.X) and Montgomery reduction is interleaved column by column (FIPS): 122
// IMAD.WIDE.U32 + 8 IMAD.HI + 8 IMAD, no .X multiply-adds at all, 30 registers. MEASURED on B200 (tools/microbench.cu,
// profiles/microbench_r02.json): an IMAD.WIDE.U32 that writes a carry predicate issues at 8.6 T/s — the same half rate as
// the carry-in form IMAD.WIDE.U32.X (9.2 T/s; plain IMAD.WIDE.U32: 17.2 T/s). So any 32-bit-limb multiplier whose
// multiply-adds touch the carry flag is bound by ≈ 4 cycles per multiply-add per SM sub-partition: f_mul_comba reaches
// 65 G products/s against 67 G/s for the carry chains, and the chains stay the default product. The SQUARE is different:
// product scanning lets it use 36 instead of 64 partial products (108 instead of 138 multiply-adds) and f_sqr_comba reaches
// 79 G/s, so f_sqr uses it.
#if defined(__CUDA_ARCH__)
DEV void mac3(uint32_t& t0, uint32_t& t1, uint32_t& t2, uint32_t x, uint32_t y) {
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
        "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32 %2, %2, 0;"
        : "+r"(t0), "+r"(t1), "+r"(t2)
        : "r"(x), "r"(y));
}
#else
inline void mac3(uint32_t& t0, uint32_t& t1, uint32_t& t2, uint32_t x, uint32_t y) {
    const uint64_t prod = (uint64_t)x * y, lo = (uint64_t)t0 + (uint32_t)prod, hi = (uint64_t)t1 + (prod >> 32) + (lo >> 32);
    t0 = (uint32_t)lo;
    t1 = (uint32_t)hi;
    t2 += (uint32_t)(hi >> 32);
}
#endif
// Montgomery product a·b·2^-256 mod P, canonical. Precondition: a < P; b any value < 2^256.
template <class C>
HD Field<C> f_mul_comba(const Field<C>& a, const Field<C>& b) {
    uint32_t t0 = 0, t1 = 0, t2 = 0, m[8];
    Field<C> r;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int j = 0; j < i; ++j) {
            mac3(t0, t1, t2, a.l[j], b.l[i - j]);
            mac3(t0, t1, t2, m[j], C::P(i - j));
        }
        mac3(t0, t1, t2, a.l[i], b.l[0]);
        m[i] = t0 * C::INV;
        mac3(t0, t1, t2, m[i], C::P(0));  // t0 becomes 0
        t0 = t1;
        t1 = t2;
        t2 = 0;
    }
#pragma unroll
    for (int i = 8; i < 16; ++i) {
#pragma unroll
        for (int j = i - 7; j < 8; ++j) {
            mac3(t0, t1, t2, a.l[j], b.l[i - j]);
            mac3(t0, t1, t2, m[j], C::P(i - j));
        }
        r.l[i - 8] = t0;
        t0 = t1;
        t1 = t2;
        t2 = 0;
    }
    return f_reduce_once<C>(r, t0);
}
// (a1·b1 + a2·b2)·2^-256 mod P with ONE Montgomery reduction (all four operands < P, so the double-width sum is < 2P² and
// the reduced value < 1.4 P: one conditional subtraction): 128 + 72 multiply-adds instead of 2 × 136. The point-addition
// formulas have one such pair each (Y3 = R·(Q − X3) − Y1·PPP).
template <class C>
HD Field<C> f_mul2_add(const Field<C>& a1, const Field<C>& b1, const Field<C>& a2, const Field<C>& b2) {
    uint32_t t0 = 0, t1 = 0, t2 = 0, m[8];
    Field<C> r;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int l = i - j;
            if (l >= 0 && l < 8) {
                mac3(t0, t1, t2, a1.l[j], b1.l[l]);
                mac3(t0, t1, t2, a2.l[j], b2.l[l]);
            }
        }
        if (i < 8) {
#pragma unroll
            for (int j = 0; j < i; ++j) mac3(t0, t1, t2, m[j], C::P(i - j));
            m[i] = t0 * C::INV;
            mac3(t0, t1, t2, m[i], C::P(0));
        } else {
#pragma unroll
            for (int j = i - 7; j < 8; ++j) mac3(t0, t1, t2, m[j], C::P(i - j));
            r.l[i - 8] = t0;
        }
        t0 = t1;
        t1 = t2;
        t2 = 0;
    }
    return f_reduce_once<C>(r, t0);
}
// Montgomery square, a < P (< 2^254). The cross products 2·a_j·a_l (j < l) are taken against the limbs of the doubled
// upper part of a — d_l = limb l of 2a for l > j+1, and a_l << 1 (no incoming bit) for l = j+1 — so a column needs one
// multiply-add per unordered pair: 36 instead of 64 products for a·a, 108 instead of 136 in all.
template <class C>
HD Field<C> f_sqr_comba(const Field<C>& a) {
    uint32_t d[8], e[8];
#pragma unroll
    for (int l = 1; l < 8; ++l) {
        d[l] = (a.l[l] << 1) | (a.l[l - 1] >> 31);
        e[l] = a.l[l] << 1;
    }
    uint32_t t0 = 0, t1 = 0, t2 = 0, m[8];
    Field<C> r;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int l = i - j;
            if (l > j && l < 8) mac3(t0, t1, t2, a.l[j], l == j + 1 ? e[l] : d[l]);
        }
        if ((i & 1) == 0 && i / 2 < 8) mac3(t0, t1, t2, a.l[i / 2], a.l[i / 2]);
        if (i < 8) {
#pragma unroll
            for (int j = 0; j < i; ++j) mac3(t0, t1, t2, m[j], C::P(i - j));
            m[i] = t0 * C::INV;
            mac3(t0, t1, t2, m[i], C::P(0));
        } else {
#pragma unroll
            for (int j = i - 7; j < 8; ++j) mac3(t0, t1, t2, m[j], C::P(i - j));
            r.l[i - 8] = t0;
        }
        t0 = t1;
        t1 = t2;
        t2 = 0;
    }
    return f_reduce_once<C>(r, t0);
}
// ---- unsaturated multiplier: 29-bit limbs, carry-free column accumulation ------------------------------------------------
// On sm_100 the carry-chained IMAD.WIDE.U32.X issues at half the rate of plain IMAD.WIDE.U32 (tools/microbench:
// 9.1 vs 17.2 T/s) and 104 of the 139 multiply-adds of f_mul_chains are .X forms. Here both operands are re-sliced into
// nine 29-bit limbs, every partial product a_j·b_i (< 2^58) is added into a 64-bit column with a plain
// `mad.wide.u32` (18 products per column at most: < 2^63, no carries), and Montgomery reduction retires one 29-bit limb
// per row — eight rows of 29 bits and a last row of 24 bits, 8·29 + 24 = 256, so the result is still a·b·2^-256 mod P in
// the halo2curves representation. 162 IMAD.WIDE + 12 IMAD instead of 243 issue-slot equivalents on the multiplier pipe —
// but the limb re-slicing and carry propagation cost ≈210 ALU-pipe instructions per product, and MEASURED on B200 this
// variant reaches 43–45 G products/s against 67 G/s for the carry chains, so it is kept only as a tested alternative
// (-DB200ZK_MUL_U29) until values can stay in 29-bit form across whole kernels.
template <class C>
HD void f_unpack29(const uint32_t* w, uint32_t* L) {
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const int off = 29 * k, wi = off >> 5, sh = off & 31;
        uint32_t v = w[wi] >> sh;
        if (sh > 3 && wi + 1 < 8) v |= w[wi + 1] << (32 - sh);
        L[k] = v & 0x1fffffffu;
    }
}
template <class C>
HD Field<C> f_mul_u29(const Field<C>& a, const Field<C>& b) {
    constexpr uint32_t M29 = 0x1fffffffu;
    uint32_t A[9], B[9], P[9];
    {
        uint32_t pw[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pw[i] = C::P(i);
        f_unpack29<C>(pw, P);
    }
    f_unpack29<C>(a.l, A);
    f_unpack29<C>(b.l, B);
    const uint32_t pinv = C::INV;  // -P^-1 mod 2^32; its low 29 (24) bits are -P^-1 mod 2^29 (2^24)
    uint64_t T[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) T[j] = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
#pragma unroll
        for (int j = 0; j < 9; ++j) T[j] += (uint64_t)A[j] * B[i];
        const uint32_t m = ((uint32_t)T[0] * pinv) & (i < 8 ? M29 : 0x00ffffffu);
#pragma unroll
        for (int j = 0; j < 9; ++j) T[j] += (uint64_t)m * P[j];
        if (i < 8) {  // T[0] is a multiple of 2^29: retire the limb
            T[1] += T[0] >> 29;
#pragma unroll
            for (int j = 0; j < 9; ++j) T[j] = T[j + 1];
            T[9] = 0;
        }
    }
    // T[0] is a multiple of 2^24 and the value is sum T[j]·2^(29j) / 2^24 < 2P: normalise to 29-bit limbs, re-slice to words
    uint32_t L[10];
    uint64_t c = 0;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        const uint64_t t = T[j] + c;
        L[j] = (uint32_t)t & M29;
        c = t >> 29;
    }
    L[9] = (uint32_t)c;
    Field<C> r;
#pragma unroll
    for (int k = 0; k < 8; ++k) r.l[k] = 0;
    uint32_t top = 0;
#pragma unroll
    for (int j = 0; j < 10; ++j) {
        const int pos = 29 * j - 24;  // bit position of limb j in the result
        if (pos < 0) {
            r.l[0] |= L[j] >> (-pos);
        } else {
            const int wi = pos >> 5, sh = pos & 31;
            if (wi < 8) r.l[wi] |= L[j] << sh;
            else top |= L[j] << sh;
            if (sh > 3) {
                if (wi + 1 < 8) r.l[wi + 1] |= L[j] >> (32 - sh);
                else top |= L[j] >> (32 - sh);
            }
        }
    }
    return f_reduce_once<C>(r, top != 0 ? 1u : 0u);
}

// host fast path: 4×64-bit limbs, same value as f_mul_chains (checked by b200zk_host_selftest)
template <class C>
inline Field<C> f_mul_host64(const Field<C>& a, const Field<C>& b) {
    typedef unsigned __int128 u128;
    uint64_t x[4], y[4], p[4], t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
        x[i] = (uint64_t)a.l[2 * i] | ((uint64_t)a.l[2 * i + 1] << 32);
        y[i] = (uint64_t)b.l[2 * i] | ((uint64_t)b.l[2 * i + 1] << 32);
        p[i] = (uint64_t)C::P(2 * i) | ((uint64_t)C::P(2 * i + 1) << 32);
    }
    // -P^-1 mod 2^64 from the 32-bit constant by one Newton step
    uint64_t inv = (uint64_t)C::INV;  // inv ≡ -P^-1 (mod 2^32)
    inv = inv * (2 + p[0] * inv);     // (-x)·(2 - P·x) with x = -inv  ->  -P^-1 (mod 2^64)
    for (int i = 0; i < 4; ++i) {
        u128 c = 0;
        for (int j = 0; j < 4; ++j) {
            c += (u128)x[j] * y[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * inv;
        c = (u128)m * p[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; ++j) {
            c += (u128)m * p[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    Field<C> r;
    for (int i = 0; i < 4; ++i) {
        r.l[2 * i] = (uint32_t)t[i];
        r.l[2 * i + 1] = (uint32_t)(t[i] >> 32);
    }
    return f_reduce_once<C>(r, (uint32_t)t[4]);
}
#if defined(B200ZK_NOINLINE_MUL) && defined(__CUDACC__)
// experiment: out-of-line multiplier bodies (smaller kernels, friendlier to the instruction caches) — see DESIGN.md §3.3
template <class C>
__device__ __noinline__ Field<C> f_mul_chains_ool(Field<C> a, Field<C> b) {
    return f_mul_chains<C>(a, b);
}
template <class C>
__device__ __noinline__ Field<C> f_sqr_comba_ool(Field<C> a) {
    return f_sqr_comba<C>(a);
}
#endif
template <class C>
HD Field<C> f_mul(const Field<C>& a, const Field<C>& b) {
#if defined(__CUDA_ARCH__) && defined(B200ZK_NOINLINE_MUL)
    return f_mul_chains_ool<C>(a, b);
#elif defined(__CUDA_ARCH__)
#if defined(B200ZK_MUL_U29)  // measured alternative: 43–45 G mul/s (profiles/README.md)
    return f_mul_u29<C>(a, b);
#elif defined(B200ZK_MUL_COMBA)  // measured alternative: 65 G mul/s (profiles/microbench_r02.json)
    return f_mul_comba<C>(a, b);
#else
    return f_mul_chains<C>(a, b);  // 67 G mul/s
#endif
#else
    return f_mul_host64<C>(a, b);
#endif
}
// squaring: the product-scanning form needs 108 instead of 138 multiply-adds — 79 G/s against 67 G/s for f_mul(a, a)
template <class C>
HD Field<C> f_sqr(const Field<C>& a) {
#if defined(__CUDA_ARCH__) && defined(B200ZK_NOINLINE_MUL)
    return f_sqr_comba_ool<C>(a);
#elif defined(__CUDA_ARCH__) && !defined(B200ZK_MUL_U29) && !defined(B200ZK_SQR_BY_MUL)
    return f_sqr_comba<C>(a);
#else
    return f_mul<C>(a, a);
#endif
}
// a1·b1 − a2·b2
template <class C>
HD Field<C> f_mul2_sub(const Field<C>& a1, const Field<C>& b1, const Field<C>& a2, const Field<C>& b2) {
#if defined(__CUDA_ARCH__) && !defined(B200ZK_NO_FUSED_MUL2)
    return f_mul2_add<C>(a1, b1, f_neg<C>(a2), b2);
#else
    return f_sub<C>(f_mul<C>(a1, b1), f_mul<C>(a2, b2));
#endif
}
// out of Montgomery form: a·2^-256 mod P (canonical integer limbs)
template <class C>
HD Field<C> f_from_mont(const Field<C>& a) {
    Field<C> one;
#pragma unroll
    for (int i = 0; i < 8; ++i) one.l[i] = i == 0 ? 1u : 0u;
    return f_mul<C>(a, one);
}
template <class C>
HD Field<C> f_to_mont(const Field<C>& a) {
    Field<C> r2;
#pragma unroll
    for (int i = 0; i < 8; ++i) r2.l[i] = C::R2(i);
    return f_mul<C>(r2, a);
}
template <class C>
HD Field<C> f_pow(const Field<C>& a, const uint32_t* e, int nlimbs) {
    Field<C> r = f_one<C>();
    for (int i = nlimbs - 1; i >= 0; --i)
        for (int b = 31; b >= 0; --b) {
            r = f_sqr<C>(r);
            if ((e[i] >> b) & 1) r = f_mul<C>(r, a);
        }
    return r;
}
template <class C>
HD Field<C> f_pow_u64(const Field<C>& a, uint64_t e) {
    uint32_t w[2] = {(uint32_t)e, (uint32_t)(e >> 32)};
    return f_pow<C>(a, w, 2);
}
// Fermat inverse (0 -> 0)
template <class C>
HD Field<C> f_inv(const Field<C>& a) {
    uint32_t e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) e[i] = C::P(i);
    e[0] -= 2;  // P[0] >= 2 for both moduli
    return f_pow<C>(a, e, 8);
}
// halo2curves from_u512: (lo + 2^256·hi) mod P in Montgomery form
template <class C>
HD Field<C> f_from_u512(const uint32_t* w) {
    Field<C> d0, d1, r2, r3;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        d0.l[i] = w[i];
        d1.l[i] = w[8 + i];
        r2.l[i] = C::R2(i);
        r3.l[i] = C::R3(i);
    }
    return f_add<C>(f_mul<C>(r2, d0), f_mul<C>(r3, d1));
}

}  // namespace b200zk
