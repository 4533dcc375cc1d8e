// ParamsKZG<Bn256>::setup on the device: g[i] = s^i·G, g_lagrange[i] = L_i(s)·G.
// Upstream derives g_lagrange with a group-FFT of g; with the trapdoor in hand the same group elements are
// L_i(s)·G with L_i(s) = (s^n − 1)·ω^i / (n·(s − ω^i)) (one batch inversion), so both bases come from one fixed-base
// kernel: 8-bit windows over a host-built table of d·256^w·G, 32 mixed additions per point, then one inversion per
// point to normalise.
#include "poly.cuh"

namespace b200zk {

#define LAUNCHED(k) do { g_launch_count += (k); CUDA_CHECK(cudaGetLastError()); } while (0)

__global__ void lagrange_denominators_kernel(Fr* out, Fr s, const Fr* table, uint32_t table_log, uint32_t k) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ((size_t)1 << k)) return;
    f_store(out + i, f_sub(s, omega_pow_from_table(table, table_log, k, (uint32_t)i)));
}
__global__ void lagrange_scalars_kernel(Fr* inv_den, Fr common, const Fr* table, uint32_t table_log, uint32_t k) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ((size_t)1 << k)) return;
    const Fr w = omega_pow_from_table(table, table_log, k, (uint32_t)i);
    f_store(inv_den + i, f_mul(f_mul(f_load(inv_den + i), common), w));
}
// table: [32][255] affine points, entry (w, d-1) = d·256^w·G
__global__ void __launch_bounds__(128) fixed_base_kernel(const Fr* scalars, size_t n, const G1Affine* table, G1Affine* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Fr s = f_from_mont(f_load(scalars + i));
    G1X acc = g1x_identity();
    for (int w = 0; w < 32; ++w) {
        const uint32_t d = (s.l[w >> 2] >> (8 * (w & 3))) & 0xff;
        if (!d) continue;
        G1Affine p;
        p.x = f_load_ro(&table[w * 255 + d - 1].x);
        p.y = f_load_ro(&table[w * 255 + d - 1].y);
        acc = g1x_add_affine(acc, p);
    }
    const G1Affine r = g1x_to_affine(acc);
    f_store(&out[i].x, r.x);
    f_store(&out[i].y, r.y);
}

void msm_build_table(Context& ctx, const G1Affine* bases, size_t n, uint32_t c, DevBuf<G1Affine>& table);
void srs_g2_raw(const Fr& s_trapdoor, uint8_t out[256]);
uint32_t msm_table_window_bits(uint32_t k, int world);

// window tables for both bases of the loaded SRS (msm.cu merged-bucket mode)
void srs_build_tables(Context& ctx) {
    if (!ctx.srs) return;
    Srs& srs = *ctx.srs;
    CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
    srs.g_tab.release();
    srs.gl_tab.release();
    srs.tab_n = 0;
    srs.tab_agreed[0] = srs.tab_agreed[1] = -1;
    if (!ctx.msm_tables_enabled) return;
    // full-range tables on every rank: column-dealt batches need all points, point-range shards index into the same table
    size_t lo = 0, len = srs.n;
    uint32_t lg = 0;
    while (((size_t)1 << (lg + 1)) <= len) ++lg;
    srs.tab_c = msm_table_window_bits(lg, ctx.world);
    const uint32_t W = (255 + srs.tab_c - 1) / srs.tab_c;
    if ((size_t)W * len >= ((size_t)1 << 31)) return;  // entry indices are 31-bit: fall back to the generic path
    // a table is W× the basis (56 GiB at k = 26): build it only while at least 35 % of the device (and 16 GiB) stays free
    // for the prover's columns; a basis without a table takes the generic windowed path
    auto fits = [&]() {
        size_t free_b = 0, total_b = 0;
        CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
        const size_t need = (size_t)W * len * sizeof(G1Affine), keep = std::max<size_t>((size_t)16 << 30, total_b / 100 * 35);
        return free_b >= need + keep;
    };
    if (fits()) msm_build_table(ctx, srs.g.get() + lo, len, srs.tab_c, srs.g_tab);
    if (fits()) msm_build_table(ctx, srs.g_lagrange.get() + lo, len, srs.tab_c, srs.gl_tab);
    srs.tab_lo = lo;
    srs.tab_n = len;
    CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
}

static std::vector<G1Affine> build_generator_table() {
    std::vector<G1Affine> t(32 * 255);
    G1Affine gen;
    gen.x = f_to_mont(Fq{{1, 0, 0, 0, 0, 0, 0, 0}});
    gen.y = f_to_mont(Fq{{2, 0, 0, 0, 0, 0, 0, 0}});
    G1X base = g1x_from_affine(gen);
    for (int w = 0; w < 32; ++w) {
        G1X acc = base;
        for (int d = 1; d <= 255; ++d) {
            t[w * 255 + d - 1] = g1x_to_affine(acc);
            acc = g1x_add(acc, base);
        }
        base = acc;  // 256 × previous base
    }
    return t;
}

void srs_setup(Context& ctx, uint32_t k, const Fr& s_trapdoor) {
    if (k < 1 || k > 26) throw std::invalid_argument("srs_setup: k out of range");
    cudaStream_t st = ctx.stream;
    const size_t n = (size_t)1 << k;
    static const std::vector<G1Affine> host_table = build_generator_table();
    DevBuf<G1Affine> table(host_table.size(), st);
    CUDA_CHECK(cudaMemcpyAsync(table.get(), host_table.data(), host_table.size() * sizeof(G1Affine), cudaMemcpyHostToDevice, st));
    auto srs = std::make_unique<Srs>();
    srs->k = k;
    srs->n = n;
    srs->g.alloc_persistent(n, st);
    srs->g_lagrange.alloc_persistent(n, st);
    // monomial basis: scalars s^i
    DevBuf<Fr> sc(n, st);
    build_twiddle_table(sc.get(), s_trapdoor, k + 1, st);  // T[i] = s^i, i < 2^k
    fixed_base_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(sc.get(), n, table.get(), srs->g.get());
    LAUNCHED(1);
    // Lagrange basis
    const TwiddleTable& tw = ctx.std_table(k);
    const Fr n_inv = f_inv(fr_from_u64(n));
    const Fr common = f_mul(f_sub(f_pow_u64(s_trapdoor, n), f_one<FrCfg>()), n_inv);
    const unsigned blocks = (unsigned)((n + 255) / 256);
    lagrange_denominators_kernel<<<blocks, 256, 0, st>>>(sc.get(), s_trapdoor, tw.t.get(), tw.log_n, k);
    LAUNCHED(1);
    fr_batch_invert(sc.get(), n, st);
    lagrange_scalars_kernel<<<blocks, 256, 0, st>>>(sc.get(), common, tw.t.get(), tw.log_n, k);
    fixed_base_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(sc.get(), n, table.get(), srs->g_lagrange.get());
    LAUNCHED(2);
    CUDA_CHECK(cudaStreamSynchronize(st));
    srs->g2_bytes.resize(256);
    srs->g2_format = 0;
    srs_g2_raw(s_trapdoor, srs->g2_bytes.data());
    ctx.srs = std::move(srs);
    srs_build_tables(ctx);
}

}  // namespace b200zk
