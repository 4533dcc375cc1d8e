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
// Generic gate evaluator (upstream's Evaluator is a DAG interpreter — `Calculation::{Add, Sub, Mul,
// Square, Double, Negate, Store}` over `ValueSource::{Constant, Intermediate, Fixed, Advice}`): calculation j produces
// intermediate j; the intermediates listed in `results` are the gate polynomials, folded into h by Horner in y in list
// order. Installed per proving key through b200zk_pk_set_gates; without a program the specialised halo2-base gate kernel
// (q·(a + b·c − d) per gate column) runs.
enum GateSrcKind : uint32_t { GATE_SRC_CONSTANT = 0, GATE_SRC_INTERMEDIATE = 1, GATE_SRC_FIXED = 2, GATE_SRC_ADVICE = 3 };
enum GateOp : uint32_t { GATE_ADD = 0, GATE_SUB = 1, GATE_MUL = 2, GATE_SQUARE = 3, GATE_DOUBLE = 4, GATE_NEGATE = 5, GATE_STORE = 6 };
struct GateSrc {
    uint32_t kind, index;
    int32_t rotation;
};
struct GateCalc {
    uint32_t op;
    GateSrc a, b;
};
constexpr uint32_t GATE_MAX_CALCS = 128, GATE_MAX_CONSTANTS = 64, GATE_MAX_RESULTS = 64;
struct GateProgramDev {
    const GateCalc* calcs = nullptr;   // device
    const Fr* constants = nullptr;     // device
    const uint32_t* results = nullptr; // device
    uint32_t ncalcs = 0, nresults = 0;
};
void h_gates_program(const QuotientArgs& Q, const GateProgramDev& P, Fr* h, cudaStream_t s);

struct LookupCosets {
    const Fr *z, *a, *s, *input, *table;
};

void h_gates(const QuotientArgs& Q, Fr* h, cudaStream_t s);
void h_permutation(const QuotientArgs& Q, Fr* h, bool final_scale, cudaStream_t s);
void h_lookup(const QuotientArgs& Q, const LookupCosets& Lk, Fr* h, bool final_scale, cudaStream_t s);

}  // namespace b200zk
