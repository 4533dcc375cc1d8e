// Arguments of the h(X) kernels (quotient.cu).
#pragma once
#include "poly.cuh"

namespace b200zk {

constexpr int Q_MAX_ADVICE = 48, Q_MAX_FIXED = 48, Q_MAX_PERM = 64, Q_MAX_SETS = 32;

struct QuotientArgs {
    uint32_t k, A, L, F, P, num_sets, blinding_factors;
    uint32_t table_log;
    unsigned long long row_begin, row_end;  // extended rows evaluated by this launch
    const Fr* table;   // twiddle table (w^j of the 2^table_log-th root), table_log >= k+2
    const Fr* t_inv;   // 4 inverted vanishing evaluations (device)
    const Fr* advice[Q_MAX_ADVICE];    // extended cosets of the A + L advice columns
    const Fr* fixed[Q_MAX_FIXED];      // extended cosets of the F + 1 + A fixed columns
    const Fr* perm_cols[Q_MAX_PERM];   // extended cosets of the P permutation columns (aliases into advice / fixed)
    const Fr* sigma[Q_MAX_PERM];       // extended cosets of the sigma polynomials
    const Fr* z[Q_MAX_SETS];           // extended cosets of the permutation products
    const Fr *l0, *l_last, *l_active;
    Fr y, beta, gamma, delta, beta_zeta;
};
struct LookupCosets {
    const Fr *z, *a, *s, *input, *table;
};

void h_gates(const QuotientArgs& Q, Fr* h, cudaStream_t s);
void h_permutation(const QuotientArgs& Q, Fr* h, bool final_scale, cudaStream_t s);
void h_lookup(const QuotientArgs& Q, const LookupCosets& Lk, Fr* h, bool final_scale, cudaStream_t s);

}  // namespace b200zk
