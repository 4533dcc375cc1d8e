// Integer-pipe microbenchmarks on sm_100a: dependent-free IMAD / IMAD.WIDE issue rates and the achieved
// Montgomery-product rate of field.cuh — the denominators DESIGN.md uses for "fraction of INT32 roof".
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I halo2-plonky2-verifier_b200/csrc tools/microbench.cu -o tools/microbench
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
// carry-chained wide multiply-adds exactly as field.cuh issues them (mad.lo.cc / madc.hi.cc pairs -> IMAD.WIDE.U32.X)
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
template <int ILP>
__global__ void fmul_kernel(Fr* out, const Fr* in, int iters) {
    Fr x[ILP];
    Fr w = in[1];
    for (int i = 0; i < ILP; ++i) x[i] = in[(threadIdx.x + i) & 7];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = f_mul(x[i], w);
    }
    Fr s = x[0];
    for (int i = 1; i < ILP; ++i) s = f_add(s, x[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F>
float time_ms(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
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
        double ops = (double)blocks * threads * iters * 8;
        printf(", \"imad_Tops\": %.3f", ops / ms / 1e9);
    }
    {
        float ms = time_ms([&] { imad_wide_kernel<8><<<blocks, threads>>>((unsigned long long*)buf, 3, iters); });
        double ops = (double)blocks * threads * iters * 8;
        printf(", \"imad_wide_Tops\": %.3f", ops / ms / 1e9);
    }
    {
        float ms = time_ms([&] { imad_wide_x_kernel<4><<<blocks, threads>>>((uint32_t*)buf, 3, 5, 7, 11, 13, iters / 4); });
        double ops = (double)blocks * threads * (iters / 4) * 4 * 4;  // wide multiply-adds (lo+hi pair = 1)
        printf(", \"imad_wide_carry_chain_Tops\": %.3f", ops / ms / 1e9);
    }
    {
        float ms = time_ms([&] { fmul_kernel<2><<<blocks, threads>>>((Fr*)buf, din, 512); });
        double ops = (double)blocks * threads * 512 * 2;
        printf(", \"fr_mul_Gops_ilp2\": %.2f", ops / ms / 1e6);
    }
    {
        float ms = time_ms([&] { fmul_kernel<1><<<blocks, threads>>>((Fr*)buf, din, 512); });
        double ops = (double)blocks * threads * 512;
        printf(", \"fr_mul_Gops_ilp1\": %.2f", ops / ms / 1e6);
    }
    printf("}\n");
    return 0;
}
