// Device-side polynomial / column helpers used between the MSM, NTT and h(X) kernels inside create_proof
// (SURVEY.md §8a rows E, F, G, I): batch inversion, prefix products, Horner evaluation, synthetic division,
// linear combinations, the lookup permutation and the ChaCha-based Fr::random stream.
#pragma once
#include "context.cuh"

namespace b200zk {

// ---- elementwise / utility -----------------------------------------------------------------------------------
void fr_fill(Fr* a, const Fr& v, size_t n, cudaStream_t s);
// out[i] = sum_j coeff[j] * polys[j][i]   (polys, coeffs: host arrays of device pointers / values, m <= 128 per call)
void fr_lincomb(Fr* out, const std::vector<const Fr*>& polys, const std::vector<Fr>& coeffs, size_t n, bool accumulate, cudaStream_t s);
// a[i] *= c
void fr_scale(Fr* a, const Fr& c, size_t n, cudaStream_t s);
// a[i] -= small[i] for i < m (m <= 8)
void fr_sub_low(Fr* a, const Fr* small_host, uint32_t m, cudaStream_t s);

// ---- halo2 batch_invert: a[i] <- a[i]^-1, zeros stay zero ------------------------------------------------------
void fr_batch_invert(Fr* a, size_t n, cudaStream_t s);
// z[0] = first; z[i] = z[i-1] * m[i-1]  for i < n (exclusive running product scaled by `first`); z may not alias m
void fr_prefix_product(Fr* z, const Fr* m, const Fr& first, size_t n, cudaStream_t s);

// ---- evaluation: out_host[j] = polys[j](point) for one point, many polynomials of n coefficients -----------------
void fr_eval_many(Context& ctx, const std::vector<const Fr*>& polys, size_t n, const Fr& point, Fr* out_host);
// q = a / (X - b) dropping the remainder (kate_division): a has n coefficients, q gets n-1 and q[n-1] = 0. q may alias a.
void fr_kate_division(Context& ctx, const Fr* a, Fr* q, size_t n, const Fr& b);

// ---- permutation / lookup argument helpers ------------------------------------------------------------------------
// m[i] (*)= beta*sigma[i] + gamma + v[i]     (first: assign instead of multiply)
void perm_denominator(Fr* m, const Fr* v, const Fr* sigma, const Fr& beta, const Fr& gamma, size_t n, bool first, cudaStream_t s);
// m[i] *= delta_pow * omega^i * beta + gamma + v[i]; omega^i read from the twiddle table of the 2^table_log root
void perm_numerator(Fr* m, const Fr* v, const Fr& delta_pow_beta, const Fr& gamma, const Fr* table, uint32_t table_log, uint32_t k, cudaStream_t s);
// p[i] = (beta + a[i]) * (gamma + s[i])
void lookup_denominator(Fr* p, const Fr* a, const Fr* sp, const Fr& beta, const Fr& gamma, size_t n, cudaStream_t s);
// p[i] *= (in[i] + beta) * (tab[i] + gamma)
void lookup_numerator(Fr* p, const Fr* in, const Fr* tab, const Fr& beta, const Fr& gamma, size_t n, cudaStream_t s);
// permute_expression_pair for values < n: a counting sort over the value domain [0, n), which covers the
// halo2-base range table. Writes rows [0, usable) of a_out / s_out. Returns (after synchronising) 0 on success or a bit
// mask: LOOKUP_UNSUPPORTED — a TABLE value is >= n (not a range-style table: outside what this sort handles, NOT a
// statement about the witness); LOOKUP_NOT_IN_TABLE — an input value is missing from the table (halo2's
// ConstraintSystemFailure; an input >= n is missing from any supported table).
constexpr int LOOKUP_UNSUPPORTED = 1, LOOKUP_NOT_IN_TABLE = 2;
// fill_from_end: ascending leftovers go to the repeated rows popped from the end (classic rule) or, false, in ascending order
int lookup_permute(Context& ctx, const Fr* input, const Fr* table, Fr* a_out, Fr* s_out, size_t n, size_t usable, bool fill_from_end = true);

// radix-4 combine of four plain size-n inverse transforms Y[j][k] (of the stride-4 subsequences of a 4n-point vector) into
// extended_to_coeff's 3n outputs, for k in [k_lo, k_hi); I = ω_ext^(−n); post3 = the domain's extended_to_coeff factors
void fr_e2c_combine(const Fr* Y, Fr* out, size_t n, size_t k_lo, size_t k_hi, const Fr* table, uint32_t table_log, uint32_t ext_k, const Fr& I, const Fr* post3,
                    cudaStream_t s);

// ---- Fr::random stream: element j of a rand_chacha BlockRng stream = from_u512 of ChaCha block (counter0 + j) ---------
void fr_random_stream(Fr* out, size_t n, const uint32_t key[8], uint64_t counter0, int rounds, cudaStream_t s);
// out[i] = Fr::from_u512 of the i-th 64-byte group of host-supplied random words (device buffer of 16·n words)
void fr_from_u512(Fr* out, const uint32_t* words_dev, size_t n, cudaStream_t s);

// ---- sigma columns from the permutation mapping: sigma[i] = delta^col(i) * omega^row(i) ------------------------------
void sigma_from_mapping(Fr* sigma, const uint32_t* map_col, const uint32_t* map_row, const Fr* delta_pows_dev, const Fr* table,
                        uint32_t table_log, uint32_t k, cudaStream_t s);

// omega^i from a twiddle table holding w^j (j < 2^(table_log-1)) of the 2^table_log-th root, for the 2^k-th root
DEV Fr omega_pow_from_table(const Fr* table, uint32_t table_log, uint32_t k, uint32_t i) {
    const uint32_t half = 1u << (k - 1);
    const uint32_t shift = table_log - k;
    if (i < half) return f_load_ro(table + ((size_t)i << shift));
    return f_neg(f_load_ro(table + ((size_t)(i - half) << shift)));
}

}  // namespace b200zk
