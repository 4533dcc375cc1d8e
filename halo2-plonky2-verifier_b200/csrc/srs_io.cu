// ParamsKZG<Bn256>::{write, read} byte layout (halo2-base caches it as
// ./params/kzg_bn254_{k}.srs — cf. the reference's .gitignore:5):
//     k as u32 LE | n × G1 (g) | n × G1 (g_lagrange) | G2 g2 | G2 s_g2
// with halo2-axiom's two encodings of a curve point [EXT, unverifiable here — isolated in this file]:
//     RawBytes  : uncompressed coordinates as raw Montgomery limbs — G1 = 64 B (byte-identical to this library's device
//                 layout), G2 = 128 B (x.c0, x.c1, y.c0, y.c1)
//     Processed : compressed canonical form — G1 = 32 B (x little-endian, bit 6 of byte 31 = y odd, bit 7 = identity),
//                 G2 = 64 B (kept as opaque bytes: only the G1 side matters to the prover).
// G1 points are validated (on-curve) / decompressed / compressed on the device; the two G2 points of a freshly set-up
// SRS (g2 generator and s·g2) come from a small host-side Fq2 implementation.
#include "poly.cuh"

namespace b200zk {

#define LAUNCHED(k) do { g_launch_count += (k); CUDA_CHECK(cudaGetLastError()); } while (0)

// ---- host Fq2 / G2 (only for s·g2 at setup time) ---------------------------------------------------------------------
struct Fq2 {
    Fq c0, c1;
};
static Fq2 fq2_add(const Fq2& a, const Fq2& b) { return {f_add(a.c0, b.c0), f_add(a.c1, b.c1)}; }
static Fq2 fq2_sub(const Fq2& a, const Fq2& b) { return {f_sub(a.c0, b.c0), f_sub(a.c1, b.c1)}; }
static Fq2 fq2_mul(const Fq2& a, const Fq2& b) {  // u^2 = -1
    return {f_sub(f_mul(a.c0, b.c0), f_mul(a.c1, b.c1)), f_add(f_mul(a.c0, b.c1), f_mul(a.c1, b.c0))};
}
static Fq2 fq2_sqr(const Fq2& a) { return fq2_mul(a, a); }
static Fq2 fq2_dbl(const Fq2& a) { return fq2_add(a, a); }
static Fq2 fq2_inv(const Fq2& a) {
    const Fq d = f_inv(f_add(f_sqr(a.c0), f_sqr(a.c1)));
    return {f_mul(a.c0, d), f_neg(f_mul(a.c1, d))};
}
static bool fq2_is_zero(const Fq2& a) { return f_is_zero(a.c0) && f_is_zero(a.c1); }
struct G2J {
    Fq2 x, y, z;  // Jacobian, identity: z = 0
};
static G2J g2_dbl(const G2J& p) {
    if (fq2_is_zero(p.z)) return p;
    const Fq2 a = fq2_sqr(p.x), b = fq2_sqr(p.y), c = fq2_sqr(b);
    const Fq2 d = fq2_dbl(fq2_sub(fq2_sub(fq2_sqr(fq2_add(p.x, b)), a), c));
    const Fq2 e = fq2_add(fq2_dbl(a), a), f = fq2_sqr(e);
    G2J r;
    r.x = fq2_sub(f, fq2_dbl(d));
    r.z = fq2_dbl(fq2_mul(p.y, p.z));
    r.y = fq2_sub(fq2_mul(e, fq2_sub(d, r.x)), fq2_dbl(fq2_dbl(fq2_dbl(c))));
    return r;
}
static G2J g2_add(const G2J& p, const G2J& q) {
    if (fq2_is_zero(p.z)) return q;
    if (fq2_is_zero(q.z)) return p;
    const Fq2 z1z1 = fq2_sqr(p.z), z2z2 = fq2_sqr(q.z);
    const Fq2 u1 = fq2_mul(p.x, z2z2), u2 = fq2_mul(q.x, z1z1);
    const Fq2 s1 = fq2_mul(fq2_mul(p.y, z2z2), q.z), s2 = fq2_mul(fq2_mul(q.y, z1z1), p.z);
    const Fq2 h = fq2_sub(u2, u1), rr = fq2_dbl(fq2_sub(s2, s1));
    if (fq2_is_zero(h)) {
        if (fq2_is_zero(rr)) return g2_dbl(p);
        return G2J{p.x, p.y, Fq2{f_zero<FqCfg>(), f_zero<FqCfg>()}};
    }
    const Fq2 i = fq2_sqr(fq2_dbl(h)), j = fq2_mul(h, i), v = fq2_mul(u1, i);
    G2J r;
    r.x = fq2_sub(fq2_sub(fq2_sqr(rr), j), fq2_dbl(v));
    r.y = fq2_sub(fq2_mul(rr, fq2_sub(v, r.x)), fq2_dbl(fq2_mul(s1, j)));
    r.z = fq2_mul(fq2_sub(fq2_sub(fq2_sqr(fq2_add(p.z, q.z)), z1z1), z2z2), h);
    return r;
}
static Fq fq_from_hex(const char* s) {
    Fq c = f_zero<FqCfg>();
    const size_t n = strlen(s);
    for (size_t i = 0; i < n; ++i) {
        const char ch = s[n - 1 - i];
        const uint32_t v = (ch >= '0' && ch <= '9') ? ch - '0' : (ch >= 'a' && ch <= 'f') ? ch - 'a' + 10 : ch - 'A' + 10;
        c.l[i / 8] |= v << (4 * (i % 8));
    }
    return f_to_mont(c);
}
// RawBytes of (g2, s·g2): 2 × {x.c0, x.c1, y.c0, y.c1} raw Montgomery limbs. g2 = the EIP-197 generator.
void srs_g2_raw(const Fr& s_trapdoor, uint8_t out[256]) {
    G2J g;
    g.x = {fq_from_hex("1800deef121f1e76426a00665e5c4479674322d4f75edadd46debd5cd992f6ed"), fq_from_hex("198e9393920d483a7260bfb731fb5d25f1aa493335a9e71297e485b7aef312c2")};
    g.y = {fq_from_hex("12c85ea5db8c6deb4aab71808dcb408fe3d1e7690c43d37b4ce6cc0166fa7daa"), fq_from_hex("090689d0585ff075ec9e99ad690c3395bc4b313370b38ef355acdadcd122975b")};
    g.z = {f_one<FqCfg>(), f_zero<FqCfg>()};
    const Fr e = f_from_mont(s_trapdoor);
    G2J acc{g.x, g.y, Fq2{f_zero<FqCfg>(), f_zero<FqCfg>()}};
    for (int i = 255; i >= 0; --i) {
        acc = g2_dbl(acc);
        if ((e.l[i >> 5] >> (i & 31)) & 1) acc = g2_add(acc, g);
    }
    auto put = [&](const G2J& p, uint8_t* o) {
        const Fq2 zi = fq2_inv(p.z), zi2 = fq2_sqr(zi), zi3 = fq2_mul(zi2, zi);
        const Fq2 x = fq2_mul(p.x, zi2), y = fq2_mul(p.y, zi3);
        memcpy(o, x.c0.l, 32);
        memcpy(o + 32, x.c1.l, 32);
        memcpy(o + 64, y.c0.l, 32);
        memcpy(o + 96, y.c1.l, 32);
    };
    put(g, out);
    put(acc, out + 128);
}

// ---- device: G1 validation / (de)compression ----------------------------------------------------------------------------
DEV bool g1_on_curve_dev(const G1Affine& p) {
    if (g1_is_identity(p)) return true;
    Fq three = f_zero<FqCfg>();
    three.l[0] = 3;
    three = f_to_mont(three);
    return f_eq(f_sqr(p.y), f_add(f_mul(f_sqr(p.x), p.x), three));
}
DEV bool fq_is_canonical(const Fq& a) {  // a < q
    for (int i = 7; i >= 0; --i) {
        const uint32_t pi = FqCfg::P(i);
        if (a.l[i] != pi) return a.l[i] < pi;
    }
    return false;
}
__global__ void g1_check_kernel(const G1Affine* pts, size_t n, uint32_t* bad) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1Affine p;
    p.x = f_load(&pts[i].x);
    p.y = f_load(&pts[i].y);
    if (!fq_is_canonical(p.x) || !fq_is_canonical(p.y) || !g1_on_curve_dev(p)) atomicExch(bad, 1u);
}
__global__ void g1_compress_kernel(const G1Affine* pts, size_t n, uint8_t* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1Affine p;
    p.x = f_load(&pts[i].x);
    p.y = f_load(&pts[i].y);
    uint32_t w[8];
    if (g1_is_identity(p)) {
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = 0;
        w[7] = 0x80000000u;
    } else {
        const Fq x = f_from_mont(p.x), y = f_from_mont(p.y);
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = x.l[j];
        w[7] |= (y.l[0] & 1u) << 30;  // bit 6 of byte 31
    }
    uint4* o = reinterpret_cast<uint4*>(out + 32 * i);
    o[0] = make_uint4(w[0], w[1], w[2], w[3]);
    o[1] = make_uint4(w[4], w[5], w[6], w[7]);
}
__global__ void __launch_bounds__(128) g1_decompress_kernel(const uint8_t* in, size_t n, G1Affine* out, uint32_t* bad) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4* q = reinterpret_cast<const uint4*>(in + 32 * i);
    const uint4 a = q[0], b = q[1];
    Fq x;
    x.l[0] = a.x; x.l[1] = a.y; x.l[2] = a.z; x.l[3] = a.w;
    x.l[4] = b.x; x.l[5] = b.y; x.l[6] = b.z; x.l[7] = b.w;
    const bool is_inf = x.l[7] >> 31;
    const uint32_t ysign = (x.l[7] >> 30) & 1u;
    x.l[7] &= 0x3fffffffu;
    G1Affine p;
    if (is_inf) {
        if (!f_is_zero(x) || ysign) atomicExch(bad, 1u);
        p.x = f_zero<FqCfg>();
        p.y = f_zero<FqCfg>();
    } else {
        if (!fq_is_canonical(x)) atomicExch(bad, 1u);
        p.x = f_to_mont(x);
        Fq three = f_zero<FqCfg>();
        three.l[0] = 3;
        const Fq rhs = f_add(f_mul(f_sqr(p.x), p.x), f_to_mont(three));
        // q ≡ 3 (mod 4): sqrt = rhs^((q+1)/4)
        uint32_t e[8];
        uint32_t carry = 1;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint64_t t = (uint64_t)FqCfg::P(j) + carry;
            e[j] = (uint32_t)t;
            carry = (uint32_t)(t >> 32);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) e[j] = (e[j] >> 2) | (j < 7 ? e[j + 1] << 30 : 0);
        Fq y = f_pow(rhs, e, 8);
        if (!f_eq(f_sqr(y), rhs)) atomicExch(bad, 1u);
        if ((f_from_mont(y).l[0] & 1u) != ysign) y = f_neg(y);
        p.y = y;
    }
    f_store(&out[i].x, p.x);
    f_store(&out[i].y, p.y);
}

size_t srs_file_size(uint32_t k, int format) {
    const size_t n = (size_t)1 << k;
    return format == 0 ? 4 + 2 * n * 64 + 256 : 4 + 2 * n * 32 + 128;
}

void srs_build_tables(Context& ctx);

// format: 0 RawBytes, 1 Processed
void srs_read(Context& ctx, const uint8_t* data, size_t len, int format) {
    if (len < 4) throw std::invalid_argument("srs_read: truncated header");
    uint32_t k;
    memcpy(&k, data, 4);
    if (k < 1 || k > 26) throw std::invalid_argument("srs_read: k out of range");
    if (format != 0 && format != 1) throw std::invalid_argument("srs_read: unknown format");
    if (len != srs_file_size(k, format)) throw std::invalid_argument("srs_read: file size does not match k and format");
    cudaStream_t s = ctx.stream;
    const size_t n = (size_t)1 << k, psz = format == 0 ? 64 : 32, gsz = format == 0 ? 128 : 64;
    auto srs = std::make_unique<Srs>();
    srs->k = k;
    srs->n = n;
    srs->g.alloc_persistent(n, s);
    srs->g_lagrange.alloc_persistent(n, s);
    DevBuf<uint32_t> bad(1, s);
    CUDA_CHECK(cudaMemsetAsync(bad.get(), 0, 4, s));
    const unsigned blocks = (unsigned)((n + 127) / 128);
    if (format == 0) {
        CUDA_CHECK(cudaMemcpyAsync(srs->g.get(), data + 4, n * 64, cudaMemcpyHostToDevice, s));
        CUDA_CHECK(cudaMemcpyAsync(srs->g_lagrange.get(), data + 4 + n * 64, n * 64, cudaMemcpyHostToDevice, s));
    } else {
        DevBuf<uint8_t> raw(2 * n * 32, s);
        CUDA_CHECK(cudaMemcpyAsync(raw.get(), data + 4, 2 * n * 32, cudaMemcpyHostToDevice, s));
        g1_decompress_kernel<<<blocks, 128, 0, s>>>(raw.get(), n, srs->g.get(), bad.get());
        g1_decompress_kernel<<<blocks, 128, 0, s>>>(raw.get() + n * 32, n, srs->g_lagrange.get(), bad.get());
        LAUNCHED(2);
    }
    g1_check_kernel<<<blocks, 128, 0, s>>>(srs->g.get(), n, bad.get());
    g1_check_kernel<<<blocks, 128, 0, s>>>(srs->g_lagrange.get(), n, bad.get());
    LAUNCHED(2);
    uint32_t h = 0;
    CUDA_CHECK(cudaMemcpyAsync(&h, bad.get(), 4, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    if (h) throw std::invalid_argument("srs_read: a G1 point is not canonical or not on the curve");
    srs->g2_format = format;
    srs->g2_bytes.assign(data + 4 + 2 * n * psz, data + 4 + 2 * n * psz + 2 * gsz);
    ctx.srs = std::move(srs);
    srs_build_tables(ctx);
}

size_t srs_write(Context& ctx, int format, uint8_t* out, size_t cap) {
    if (!ctx.srs) throw std::runtime_error("no SRS loaded");
    if (format != 0 && format != 1) throw std::invalid_argument("srs_write: unknown format");
    Srs& srs = *ctx.srs;
    const size_t need = srs_file_size(srs.k, format);
    if (cap < need) throw std::invalid_argument("srs_write: output buffer too small");
    if (srs.g2_bytes.empty() || srs.g2_format != format)
        throw std::runtime_error("srs_write: the G2 points are not available in this format (set up or read the SRS in the same format)");
    cudaStream_t s = ctx.stream;
    const size_t n = srs.n;
    memcpy(out, &srs.k, 4);
    if (format == 0) {
        CUDA_CHECK(cudaMemcpyAsync(out + 4, srs.g.get(), n * 64, cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaMemcpyAsync(out + 4 + n * 64, srs.g_lagrange.get(), n * 64, cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaStreamSynchronize(s));
        memcpy(out + 4 + 2 * n * 64, srs.g2_bytes.data(), 256);
    } else {
        DevBuf<uint8_t> raw(2 * n * 32, s);
        const unsigned blocks = (unsigned)((n + 127) / 128);
        g1_compress_kernel<<<blocks, 128, 0, s>>>(srs.g.get(), n, raw.get());
        g1_compress_kernel<<<blocks, 128, 0, s>>>(srs.g_lagrange.get(), n, raw.get() + n * 32);
        LAUNCHED(2);
        CUDA_CHECK(cudaMemcpyAsync(out + 4, raw.get(), 2 * n * 32, cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaStreamSynchronize(s));
        memcpy(out + 4 + 2 * n * 32, srs.g2_bytes.data(), 128);
    }
    return need;
}

}  // namespace b200zk
