// Integer-pipe microbenchmarks on sm_100a: issue rates of the IMAD forms the two multipliers of field.cuh are made of,
// the achieved Montgomery product / square rates of both multipliers, and a bit-exactness cross-check between them —
// the denominators DESIGN.md uses for "fraction of INT32 roof". Prints one JSON object.
// build (cudart linked dynamically; the binary is not tracked):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -cudart shared -I halo2-plonky2-verifier_b200/csrc tools/microbench.cu -o tools/microbench
#include <cstdio>
#include <cuda_runtime.h>
#include "field.cuh"
using namespace b200zk;

template <int ILP>
__global__ void imad_kernel(uint32_t* out, uint32_t a, uint32_t b, int iters) {
    uint32_t x[ILP];
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
    }
    uint32_t s = 0;
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void imad_wide_kernel(unsigned long long* out, uint32_t a, int iters) {
    unsigned long long x[ILP];
    uint32_t m[ILP];
    for (int i = 0; i < ILP; ++i) { x[i] = threadIdx.x + i; m[i] = threadIdx.x * 7 + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[i]) : "r"(m[i]), "r"(a));
    }
    unsigned long long s = 0;
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// carry-chained wide multiply-adds as f_mul_chains issues them (mad.lo.cc / madc.hi.cc pairs -> IMAD.WIDE.U32.X)
template <int CH>
__global__ void imad_wide_x_kernel(uint32_t* out, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t x, int iters) {
    uint32_t acc[CH][8];
    for (int c = 0; c < CH; ++c)
        for (int i = 0; i < 8; ++i) acc[c][i] = threadIdx.x + c * 8 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            asm volatile(
                "mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
                "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
                "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
                "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
                "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
                "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
                "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
                "madc.hi.u32 %7, %11, %12, %7;"
                : "+r"(acc[c][0]), "+r"(acc[c][1]), "+r"(acc[c][2]), "+r"(acc[c][3]), "+r"(acc[c][4]), "+r"(acc[c][5]), "+r"(acc[c][6]), "+r"(acc[c][7])
                : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(x));
        }
    }
    uint32_t s = 0;
    for (int c = 0; c < CH; ++c)
        for (int i = 0; i < 8; ++i) s += acc[c][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// the column step of f_mul_comba: plain IMAD.WIDE.U32 with a carry-out predicate + IADD3.X on the ALU pipe
template <int ILP>
__global__ void mac3_kernel(uint32_t* out, uint32_t a, uint32_t b, int iters) {
    uint32_t t0[ILP], t1[ILP], t2[ILP];
    for (int i = 0; i < ILP; ++i) { t0[i] = threadIdx.x + i; t1[i] = threadIdx.x * 3 + i; t2[i] = 0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) mac3(t0[i], t1[i], t2[i], a + i, b);
    }
    uint32_t s = 0;
    for (int i = 0; i < ILP; ++i) s += t0[i] + t1[i] + t2[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// FP64 pipe: dependent-free DFMA rate, and whether DFMA co-issues with carry-chained IMAD.WIDE (the question behind a
// dual-pipe multiplier: 24-bit limbs in doubles accumulate exactly in the FMA, on a pipe the integer multiplier leaves idle).
// mode bit 0: carry chains of f_mul_chains' row primitive, bit 1: DFMA stream, both in the same thread when mode == 3.
template <int MODE>
__global__ void coissue_kernel(double* out, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t x, double m, int iters) {
    uint32_t acc[2][8];
    double d[8];
    for (int c = 0; c < 2; ++c)
        for (int i = 0; i < 8; ++i) acc[c][i] = threadIdx.x + c * 8 + i;
    for (int i = 0; i < 8; ++i) d[i] = (double)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
        if (MODE & 1) {
#pragma unroll
            for (int c = 0; c < 2; ++c)
                asm volatile(
                    "mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
                    "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
                    "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
                    "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
                    "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
                    "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
                    "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
                    "madc.hi.u32 %7, %11, %12, %7;"
                    : "+r"(acc[c][0]), "+r"(acc[c][1]), "+r"(acc[c][2]), "+r"(acc[c][3]), "+r"(acc[c][4]), "+r"(acc[c][5]), "+r"(acc[c][6]), "+r"(acc[c][7])
                    : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(x));
        }
        if (MODE & 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rz.f64 %0, %0, %1, %0;" : "+d"(d[i]) : "d"(m));
        }
    }
    double s = 0;
    for (int c = 0; c < 2; ++c)
        for (int i = 0; i < 8; ++i) s += acc[c][i];
    for (int i = 0; i < 8; ++i) s += d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE, int ILP, class C>
__global__ void fmul_kernel(Field<C>* out, const Field<C>* in, int iters) {
    Field<C> x[ILP];
    Field<C> w = in[1];
    for (int i = 0; i < ILP; ++i) x[i] = in[(threadIdx.x + i) & 7];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) x[i] = f_mul_chains<C>(x[i], w);
            if (MODE == 1) x[i] = f_mul_comba<C>(x[i], w);
            if (MODE == 2) x[i] = f_sqr_comba<C>(x[i]);
            if (MODE == 3) x[i] = f_mul2_add<C>(x[i], w, w, x[i]);
        }
    }
    Field<C> s = x[0];
    for (int i = 1; i < ILP; ++i) s = f_add(s, x[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// bit-exactness: comba product / square against the carry-chain product on pseudo-random and edge operands
template <class C>
__global__ void check_kernel(unsigned long long* bad, unsigned long long seed, int per_thread) {
    unsigned long long s = seed + 0x9e3779b97f4a7c15ull * (blockIdx.x * blockDim.x + threadIdx.x + 1);
    auto next = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 16); };
    unsigned long long n_bad = 0;
    for (int it = 0; it < per_thread; ++it) {
        Field<C> a, b;
        for (int i = 0; i < 8; ++i) { a.l[i] = next(); b.l[i] = next(); }
        if (it % 7 == 1) for (int i = 0; i < 8; ++i) a.l[i] = C::P(i) - (i == 0 ? 1 + it % 3 : 0);
        if (it % 11 == 2) for (int i = 0; i < 8; ++i) b.l[i] = 0xffffffffu;
        if (it % 13 == 3) for (int i = 0; i < 8; ++i) a.l[i] = i == 0 ? it % 5 : 0;
        a.l[7] &= 0x3fffffffu;
        a = f_reduce_once<C>(a, 0);
        a = f_reduce_once<C>(a, 0);
        const Field<C> r0 = f_mul_chains<C>(a, b), r1 = f_mul_comba<C>(a, b);
        const Field<C> s0 = f_mul_chains<C>(a, a), s1 = f_sqr_comba<C>(a);
        Field<C> c = b;
        c.l[7] &= 0x3fffffffu;
        c = f_reduce_once<C>(c, 0);
        c = f_reduce_once<C>(c, 0);
        const Field<C> d0 = f_sub<C>(f_mul_chains<C>(a, c), f_mul_chains<C>(c, s0)), d1 = f_mul2_add<C>(a, c, f_neg<C>(c), s0);
        if (!f_eq(r0, r1)) ++n_bad;
        if (!f_eq(s0, s1)) ++n_bad;
        if (!f_eq(d0, d1)) ++n_bad;
    }
    if (n_bad) atomicAdd(bad, n_bad);
}
template <class F>
float time_ms(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
    void* buf; cudaMalloc(&buf, (size_t)sms * 8 * 1024 * 64);
    Fr h[8]; for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) h[i].l[j] = FrCfg::R2(j) ^ (i == 0 ? 0 : (i * 0x01010101u & 0x0fffffffu) * (j < 7));
    Fr* din; cudaMalloc(&din, sizeof(h)); cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice);
    const int blocks = sms * 8, threads = 256, iters = 4096;
    {
        float ms = time_ms([&] { imad_kernel<8><<<blocks, threads>>>((uint32_t*)buf, 3, 5, iters); });
        printf(", \"imad_Tops\": %.3f", (double)blocks * threads * iters * 8 / ms / 1e9);
    }
    {
        float ms = time_ms([&] { imad_wide_kernel<8><<<blocks, threads>>>((unsigned long long*)buf, 3, iters); });
        printf(", \"imad_wide_Tops\": %.3f", (double)blocks * threads * iters * 8 / ms / 1e9);
    }
    {
        float ms = time_ms([&] { imad_wide_x_kernel<4><<<blocks, threads>>>((uint32_t*)buf, 3, 5, 7, 11, 13, iters / 4); });
        printf(", \"imad_wide_carry_chain_Tops\": %.3f", (double)blocks * threads * (iters / 4) * 4 * 4 / ms / 1e9);
    }
    {
        float ms = time_ms([&] { mac3_kernel<8><<<blocks, threads>>>((uint32_t*)buf, 3, 5, iters); });
        printf(", \"imad_wide_carry_out_plus_iadd3x_Tops\": %.3f", (double)blocks * threads * iters * 8 / ms / 1e9);
    }
    {
        // per iteration and thread: 8 wide multiply-adds (mode bit 0) and / or 8 DFMA (mode bit 1)
        const int it2 = 2048;
        float t1 = time_ms([&] { coissue_kernel<1><<<blocks, threads>>>((double*)buf, 3, 5, 7, 11, 13, 1e-9, it2); });
        float t2 = time_ms([&] { coissue_kernel<2><<<blocks, threads>>>((double*)buf, 3, 5, 7, 11, 13, 1e-9, it2); });
        float t3 = time_ms([&] { coissue_kernel<3><<<blocks, threads>>>((double*)buf, 3, 5, 7, 11, 13, 1e-9, it2); });
        const double ops = (double)blocks * threads * it2 * 8;
        printf(", \"dfma_Tops\": %.3f, \"coissue_ms\": {\"imad_wide_x_only\": %.3f, \"dfma_only\": %.3f, \"both\": %.3f}", ops / t2 / 1e9, t1, t2, t3);
    }
    const char* names[4] = {"chains_mul", "comba_mul", "comba_sqr", "fused_dual_product"};
    for (int mode = 0; mode < 4; ++mode) {
        float ms1 = 0, ms2 = 0;
        if (mode == 0) { ms1 = time_ms([&] { fmul_kernel<0, 1, FrCfg><<<blocks, threads>>>((Fr*)buf, din, 512); }); ms2 = time_ms([&] { fmul_kernel<0, 2, FrCfg><<<blocks, threads>>>((Fr*)buf, din, 512); }); }
        if (mode == 1) { ms1 = time_ms([&] { fmul_kernel<1, 1, FrCfg><<<blocks, threads>>>((Fr*)buf, din, 512); }); ms2 = time_ms([&] { fmul_kernel<1, 2, FrCfg><<<blocks, threads>>>((Fr*)buf, din, 512); }); }
        if (mode == 2) { ms1 = time_ms([&] { fmul_kernel<2, 1, FrCfg><<<blocks, threads>>>((Fr*)buf, din, 512); }); ms2 = time_ms([&] { fmul_kernel<2, 2, FrCfg><<<blocks, threads>>>((Fr*)buf, din, 512); }); }
        if (mode == 3) { ms1 = time_ms([&] { fmul_kernel<3, 1, FrCfg><<<blocks, threads>>>((Fr*)buf, din, 512); }); ms2 = time_ms([&] { fmul_kernel<3, 2, FrCfg><<<blocks, threads>>>((Fr*)buf, din, 512); }); }
        printf(", \"fr_%s_Gops_ilp1\": %.2f, \"fr_%s_Gops_ilp2\": %.2f", names[mode], (double)blocks * threads * 512 / ms1 / 1e6, names[mode],
               (double)blocks * threads * 512 * 2 / ms2 / 1e6);
    }
    {
        unsigned long long* bad; cudaMalloc(&bad, 8); cudaMemset(bad, 0, 8);
        check_kernel<FrCfg><<<sms * 4, 128>>>(bad, 1, 256);
        check_kernel<FqCfg><<<sms * 4, 128>>>(bad, 2, 256);
        unsigned long long hb = ~0ull; cudaMemcpy(&hb, bad, 8, cudaMemcpyDeviceToHost);
        printf(", \"comba_vs_chains_checked\": %llu, \"comba_vs_chains_mismatches\": %llu", (unsigned long long)sms * 4 * 128 * 256 * 2 * 3, hb);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf(", \"cuda\": \"%s\"}\n", cudaGetErrorString(e));
    return e == cudaSuccess ? 0 : 1;
}
