// BN254 G1 (y^2 = x^3 + 3 over Fq) point arithmetic for the MSM kernels and the host-side folds
// (replaces halo2curves::bn256::{G1Affine, G1}).
// Affine points use the halo2curves layout (x‖y, 64 B, identity = (0,0)). Accumulators use extended
// Jacobian "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2): mixed add 8M+2S, no inversions; the
// representation never leaves the library — results are normalised to canonical affine at the boundary.
// Squares use the 108-multiply-add square of field.cuh and each formula's Y3 = A·B − C·D is one fused dual product with a
// single Montgomery reduction (f_mul2_sub): a mixed addition costs 6 products + 2 squares + 1 dual product.
#pragma once
#include "field.cuh"

namespace b200zk {

struct alignas(16) G1Affine {
    Fq x, y;
};
struct alignas(16) G1X {
    Fq x, y, zz, zzz;  // identity <=> zz == 0
};

HD bool g1_is_identity(const G1Affine& p) { return f_is_zero(p.x) && f_is_zero(p.y); }
HD bool g1_is_identity(const G1X& p) { return f_is_zero(p.zz); }
HD G1X g1x_identity() {
    G1X r;
    r.x = f_zero<FqCfg>();
    r.y = f_zero<FqCfg>();
    r.zz = f_zero<FqCfg>();
    r.zzz = f_zero<FqCfg>();
    return r;
}
HD G1X g1x_from_affine(const G1Affine& p) {
    if (g1_is_identity(p)) return g1x_identity();
    G1X r;
    r.x = p.x;
    r.y = p.y;
    r.zz = f_one<FqCfg>();
    r.zzz = f_one<FqCfg>();
    return r;
}
HD G1Affine g1_neg(const G1Affine& p) {
    G1Affine r;
    r.x = p.x;
    r.y = f_neg(p.y);
    return r;
}
HD G1X g1x_neg(const G1X& p) {
    G1X r = p;
    r.y = f_neg(p.y);
    return r;
}
// dbl-2008-s-1 (a = 0)
HD G1X g1x_dbl(const G1X& p) {
    if (g1_is_identity(p)) return p;
    Fq u = f_dbl(p.y), v = f_sqr(u), w = f_mul(u, v), s = f_mul(p.x, v);
    Fq xx = f_sqr(p.x), m = f_add(f_dbl(xx), xx);
    G1X r;
    r.x = f_sub(f_sqr(m), f_dbl(s));
    r.y = f_mul2_sub(m, f_sub(s, r.x), w, p.y);
    r.zz = f_mul(v, p.zz);
    r.zzz = f_mul(w, p.zzz);
    return r;
}
HD G1X g1x_dbl_affine(const G1Affine& p) {
    Fq u = f_dbl(p.y), v = f_sqr(u), w = f_mul(u, v), s = f_mul(p.x, v);
    Fq xx = f_sqr(p.x), m = f_add(f_dbl(xx), xx);
    G1X r;
    r.x = f_sub(f_sqr(m), f_dbl(s));
    r.y = f_mul2_sub(m, f_sub(s, r.x), w, p.y);
    r.zz = v;
    r.zzz = w;
    return r;
}
// The doubling branches of the addition formulas are taken only when both operands are the same point — never for
// distinct SRS points — so on the device they are calls to out-of-line copies: the hot kernels stay a third shorter
// (instruction-cache pressure is a measured stall of msm_accumulate_kernel, profiles/ncu_summary_r02.md).
#if defined(__CUDACC__) && !defined(B200ZK_INLINE_COLD_PATHS)
static __device__ __noinline__ G1X g1x_dbl_affine_cold(G1Affine b) { return g1x_dbl_affine(b); }
static __device__ __noinline__ G1X g1x_dbl_cold(G1X a) { return g1x_dbl(a); }
#endif
HD G1X g1x_dbl_affine_rare(const G1Affine& b) {
#if defined(__CUDA_ARCH__) && !defined(B200ZK_INLINE_COLD_PATHS)
    return g1x_dbl_affine_cold(b);
#else
    return g1x_dbl_affine(b);
#endif
}
HD G1X g1x_dbl_rare(const G1X& a) {
#if defined(__CUDA_ARCH__) && !defined(B200ZK_INLINE_COLD_PATHS)
    return g1x_dbl_cold(a);
#else
    return g1x_dbl(a);
#endif
}
// madd-2008-s with the exceptional cases handled (identity operands, equal or opposite points)
HD G1X g1x_add_affine(const G1X& a, const G1Affine& b) {
    if (g1_is_identity(b)) return a;
    if (g1_is_identity(a)) return g1x_from_affine(b);
    Fq u2 = f_mul(b.x, a.zz), s2 = f_mul(b.y, a.zzz);
    Fq p = f_sub(u2, a.x), r = f_sub(s2, a.y);
    if (f_is_zero(p)) {
        if (f_is_zero(r)) return g1x_dbl_affine_rare(b);
        return g1x_identity();
    }
    Fq pp = f_sqr(p), ppp = f_mul(p, pp), q = f_mul(a.x, pp);
    G1X o;
    o.x = f_sub(f_sub(f_sqr(r), ppp), f_dbl(q));
    o.y = f_mul2_sub(r, f_sub(q, o.x), a.y, ppp);  // two products, one Montgomery reduction
    o.zz = f_mul(a.zz, pp);
    o.zzz = f_mul(a.zzz, ppp);
    return o;
}
// add-2008-s
HD G1X g1x_add(const G1X& a, const G1X& b) {
    if (g1_is_identity(b)) return a;
    if (g1_is_identity(a)) return b;
    Fq u1 = f_mul(a.x, b.zz), u2 = f_mul(b.x, a.zz);
    Fq s1 = f_mul(a.y, b.zzz), s2 = f_mul(b.y, a.zzz);
    Fq p = f_sub(u2, u1), r = f_sub(s2, s1);
    if (f_is_zero(p)) {
        if (f_is_zero(r)) return g1x_dbl_rare(a);
        return g1x_identity();
    }
    Fq pp = f_sqr(p), ppp = f_mul(p, pp), q = f_mul(u1, pp);
    G1X o;
    o.x = f_sub(f_sub(f_sqr(r), ppp), f_dbl(q));
    o.y = f_mul2_sub(r, f_sub(q, o.x), s1, ppp);
    o.zz = f_mul(f_mul(a.zz, b.zz), pp);
    o.zzz = f_mul(f_mul(a.zzz, b.zzz), ppp);
    return o;
}
// canonical affine (one inversion)
HD G1Affine g1x_to_affine(const G1X& p) {
    G1Affine r;
    if (g1_is_identity(p)) {
        r.x = f_zero<FqCfg>();
        r.y = f_zero<FqCfg>();
        return r;
    }
    Fq i = f_inv(f_mul(p.zz, p.zzz));
    r.x = f_mul(p.x, f_mul(i, p.zzz));  // X / ZZ
    r.y = f_mul(p.y, f_mul(i, p.zz));   // Y / ZZZ
    return r;
}
// k·P by double-and-add over `nbits` low bits of a canonical (non-Montgomery) scalar
HD G1X g1x_mul_bits(const G1X& p, const uint32_t* e, int nbits) {
    G1X r = g1x_identity();
    for (int i = nbits - 1; i >= 0; --i) {
        r = g1x_dbl(r);
        if ((e[i >> 5] >> (i & 31)) & 1) r = g1x_add(r, p);
    }
    return r;
}

}  // namespace b200zk
