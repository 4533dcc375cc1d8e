// Quotient numerator h(X) over the extended domain, fused with divide_by_vanishing_poly (SURVEY.md §8a row H,
// Appendix D.8/D.9; replaces halo2_proofs::plonk::evaluation::Evaluator::evaluate_h + the t_inv scaling of
// vanishing::Argument::construct). ConstraintSystem = halo2-base's BaseConfig: A vertical
// gates q·(a + b·c − d) on rotations 0..3 of one advice column each, a permutation argument over P columns in sets
// of 2, L single-column lookups against one table column.
//
// Rows [row_begin, row_end) are evaluated (the whole extended domain on one GPU, a contiguous slice per rank when sharded).
// One thread per extended row i (4n rows): all column reads are coalesced 32-byte elements; rotations are row
// offsets of ±4·r inside the same column (served by L1/L2). The running value is y-Horner-accumulated in
// registers across ALL terms of a part, so h is written once per kernel: gates → permutation → one kernel per
// lookup (its three cosets are transient, as upstream).
#include "quotient.cuh"

namespace b200zk {

#define LAUNCHED(k) do { g_launch_count += (k); CUDA_CHECK(cudaGetLastError()); } while (0)

DEV size_t rot_idx(size_t i, int r, size_t en) { return (i + en + (size_t)((long long)r * 4)) & (en - 1); }

__global__ void __launch_bounds__(256) h_gates_kernel(QuotientArgs Q, Fr* h) {
    const size_t en = (size_t)4 << Q.k;
    const size_t i = Q.row_begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Q.row_end) return;
    const size_t i1 = rot_idx(i, 1, en), i2 = rot_idx(i, 2, en), i3 = rot_idx(i, 3, en);
    Fr v = f_zero<FrCfg>();
    for (uint32_t c = 0; c < Q.A; ++c) {
        const Fr* a = Q.advice[c];
        const Fr g = f_sub(f_add(f_load(a + i), f_mul(f_load(a + i1), f_load(a + i2))), f_load(a + i3));
        v = f_mul2_add(v, Q.y, f_load(Q.fixed[Q.F + 1 + c] + i), g);  // v·y + q·g with one Montgomery reduction
    }
    f_store(h + i, v);
}

// the interpreter: one thread per extended row, intermediates in local memory (L1-resident), program and constants read
// through the read-only path (every thread reads the same words)
__global__ void __launch_bounds__(128) h_gates_program_kernel(QuotientArgs Q, GateProgramDev P, Fr* h) {
    const size_t en = (size_t)4 << Q.k;
    const size_t i = Q.row_begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Q.row_end) return;
    Fr vals[GATE_MAX_CALCS];
    auto fetch = [&](const GateSrc& src, const Fr* done) -> Fr {
        switch (src.kind) {
            case GATE_SRC_CONSTANT: return f_load_ro(P.constants + src.index);
            case GATE_SRC_INTERMEDIATE: return done[src.index];
            case GATE_SRC_FIXED: return f_load(Q.fixed[src.index] + rot_idx(i, src.rotation, en));
            default: return f_load(Q.advice[src.index] + rot_idx(i, src.rotation, en));
        }
    };
    for (uint32_t j = 0; j < P.ncalcs; ++j) {
        const GateCalc c = P.calcs[j];
        const Fr a = fetch(c.a, vals);
        Fr r;
        switch (c.op) {
            case GATE_ADD: r = f_add(a, fetch(c.b, vals)); break;
            case GATE_SUB: r = f_sub(a, fetch(c.b, vals)); break;
            case GATE_MUL: r = f_mul(a, fetch(c.b, vals)); break;
            case GATE_SQUARE: r = f_sqr(a); break;
            case GATE_DOUBLE: r = f_dbl(a); break;
            case GATE_NEGATE: r = f_neg(a); break;
            default: r = a; break;  // GATE_STORE
        }
        vals[j] = r;
    }
    Fr v = f_zero<FrCfg>();
    for (uint32_t g = 0; g < P.nresults; ++g) v = f_add(f_mul(v, Q.y), vals[__ldg(P.results + g)]);
    f_store(h + i, v);
}

__global__ void __launch_bounds__(256) h_permutation_kernel(QuotientArgs Q, Fr* h, int final_scale) {
    const size_t en = (size_t)4 << Q.k;
    const size_t i = Q.row_begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Q.row_end) return;
    const size_t r_next = rot_idx(i, 1, en), r_last = rot_idx(i, -(int)(Q.blinding_factors + 1), en);
    const Fr one = f_one<FrCfg>();
    const Fr l0 = f_load(Q.l0 + i), l_last = f_load(Q.l_last + i), l_active = f_load(Q.l_active + i);
    Fr v = f_load(h + i);
    const uint32_t ns = Q.num_sets;
    {
        const Fr z0 = f_load(Q.z[0] + i);
        v = f_mul2_add(v, Q.y, f_sub(one, z0), l0);
        const Fr zl = f_load(Q.z[ns - 1] + i);
        v = f_mul2_add(v, Q.y, f_sub(f_sqr(zl), zl), l_last);
    }
    for (uint32_t s = 1; s < ns; ++s) v = f_mul2_add(v, Q.y, f_sub(f_load(Q.z[s] + i), f_load(Q.z[s - 1] + r_last)), l0);
    // beta·zeta·omega_ext^i, then ·delta per column
    Fr current_delta = f_mul(Q.beta_zeta, omega_pow_from_table(Q.table, Q.table_log, Q.k + 2, (uint32_t)i));
    for (uint32_t s = 0; s < ns; ++s) {
        const uint32_t j0 = s * 2, j1 = j0 + 2 < Q.P ? j0 + 2 : Q.P;
        Fr left = f_load(Q.z[s] + r_next), right = f_load(Q.z[s] + i);
        for (uint32_t j = j0; j < j1; ++j) {
            const Fr val = f_load(Q.perm_cols[j] + i);
            left = f_mul(left, f_add(f_add(val, f_mul(Q.beta, f_load(Q.sigma[j] + i))), Q.gamma));
            right = f_mul(right, f_add(f_add(val, current_delta), Q.gamma));
            current_delta = f_mul(current_delta, Q.delta);
        }
        v = f_mul2_add(v, Q.y, f_sub(left, right), l_active);
    }
    if (final_scale) v = f_mul(v, f_load_ro(Q.t_inv + (i & 3)));
    f_store(h + i, v);
}

__global__ void __launch_bounds__(256) h_lookup_kernel(QuotientArgs Q, LookupCosets Lk, Fr* h, int final_scale) {
    const size_t en = (size_t)4 << Q.k;
    const size_t i = Q.row_begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Q.row_end) return;
    const size_t r_next = rot_idx(i, 1, en), r_prev = rot_idx(i, -1, en);
    const Fr one = f_one<FrCfg>();
    const Fr l0 = f_load(Q.l0 + i), l_last = f_load(Q.l_last + i), l_active = f_load(Q.l_active + i);
    const Fr z = f_load(Lk.z + i), a = f_load(Lk.a + i), sp = f_load(Lk.s + i);
    const Fr table_value = f_mul(f_add(f_load(Lk.input + i), Q.beta), f_add(f_load(Lk.table + i), Q.gamma));
    const Fr a_minus_s = f_sub(a, sp);
    Fr v = f_load(h + i);
    v = f_mul2_add(v, Q.y, f_sub(one, z), l0);
    v = f_mul2_add(v, Q.y, f_sub(f_sqr(z), z), l_last);
    {
        const Fr lhs = f_mul(f_mul(f_load(Lk.z + r_next), f_add(a, Q.beta)), f_add(sp, Q.gamma));
        v = f_mul2_add(v, Q.y, f_sub(lhs, f_mul(z, table_value)), l_active);
    }
    v = f_mul2_add(v, Q.y, a_minus_s, l0);
    v = f_mul2_add(v, Q.y, f_mul(a_minus_s, f_sub(a, f_load(Lk.a + r_prev))), l_active);
    if (final_scale) v = f_mul(v, f_load_ro(Q.t_inv + (i & 3)));
    f_store(h + i, v);
}

void h_gates(const QuotientArgs& Q, Fr* h, cudaStream_t s) {
    const size_t en = Q.row_end - Q.row_begin;
    const int prof_h = prof_begin(PROF_QUOTIENT, s, (double)en);
    h_gates_kernel<<<(unsigned)((en + 255) / 256), 256, 0, s>>>(Q, h);
    prof_end(prof_h, s);
    LAUNCHED(1);
}
void h_gates_program(const QuotientArgs& Q, const GateProgramDev& P, Fr* h, cudaStream_t s) {
    const size_t en = Q.row_end - Q.row_begin;
    const int prof_h = prof_begin(PROF_QUOTIENT, s, (double)en);
    h_gates_program_kernel<<<(unsigned)((en + 127) / 128), 128, 0, s>>>(Q, P, h);
    prof_end(prof_h, s);
    LAUNCHED(1);
}
void h_permutation(const QuotientArgs& Q, Fr* h, bool final_scale, cudaStream_t s) {
    const size_t en = Q.row_end - Q.row_begin;
    const int prof_h = prof_begin(PROF_QUOTIENT, s, (double)en);
    h_permutation_kernel<<<(unsigned)((en + 255) / 256), 256, 0, s>>>(Q, h, final_scale);
    prof_end(prof_h, s);
    LAUNCHED(1);
}
void h_lookup(const QuotientArgs& Q, const LookupCosets& Lk, Fr* h, bool final_scale, cudaStream_t s) {
    const size_t en = Q.row_end - Q.row_begin;
    const int prof_h = prof_begin(PROF_QUOTIENT, s, (double)en);
    h_lookup_kernel<<<(unsigned)((en + 255) / 256), 256, 0, s>>>(Q, Lk, h, final_scale);
    prof_end(prof_h, s);
    LAUNCHED(1);
}

}  // namespace b200zk
