// Types of the device-resident prover (prover.cu).
#pragma once
#include "host/transcript.hpp"
#include "quotient.cuh"

namespace b200zk {

// ConstraintSystem of halo2-base's BaseConfig: column order and counts.
struct Shape {
    uint32_t k, A, L, F;
    size_t n() const { return (size_t)1 << k; }
    uint32_t num_advice() const { return A + L; }
    uint32_t num_fixed() const { return F + 1 + A; }  // constants, lookup table, gate selectors
    uint32_t table_col() const { return F; }
    uint32_t selector_col(uint32_t c) const { return F + 1 + c; }
    uint32_t num_perm() const { return F + A + L; }   // constants, gate advice, lookup advice
    static constexpr uint32_t blinding_factors = 6;
    static constexpr uint32_t degree = 4;
    static constexpr uint32_t chunk_len = degree - 2;
    size_t usable_rows() const { return n() - (blinding_factors + 1); }
    uint32_t num_sets() const { return (num_perm() + chunk_len - 1) / chunk_len; }
    bool perm_is_fixed(uint32_t j) const { return j < F; }
    uint32_t perm_col_index(uint32_t j) const { return j < F ? j : j - F; }
    size_t proof_size() const {
        size_t points = num_advice() + 2 * L + num_sets() + L + 1 + 3 + 2;
        size_t evals = 4 * A + L + num_fixed() + 1 + num_perm() + (3 * num_sets() - 1) + 5 * L;
        return 32 * (points + evals);
    }
};

struct SynthesisError : std::runtime_error {
    explicit SynthesisError(const std::string& s) : std::runtime_error(s) {}
};

// plonk::ProvingKey (+ VerifyingKey commitments), all columns device-resident
struct ProvingKeyDev {
    Shape shape;
    DevBuf<Fr> fixed_values, fixed_polys, fixed_cosets;  // [num_fixed][n], [num_fixed][n], [num_fixed][4n]
    DevBuf<Fr> sigma_values, sigma_polys, sigma_cosets;  // [num_perm][...]
    DevBuf<Fr> l_polys;                                  // l0, l_last, l_active_row on the extended domain [3][4n]
    std::vector<G1Affine> fixed_commitments, perm_commitments;
    Fr transcript_repr;
    // optional generic gate program (b200zk_pk_set_gates); empty = the specialised halo2-base gate kernel
    DevBuf<GateCalc> gate_calcs;
    DevBuf<Fr> gate_constants;
    DevBuf<uint32_t> gate_results;
    GateProgramDev gate_program() const {
        GateProgramDev p;
        p.calcs = gate_calcs.get();
        p.constants = gate_constants.get();
        p.results = gate_results.get();
        p.ncalcs = (uint32_t)gate_calcs.size();
        p.nresults = (uint32_t)gate_results.size();
        return p;
    }
};
// validates and installs a gate program (throws std::invalid_argument): indices in range, advice rotations within the
// proof's query set (gate columns 0..3, lookup columns 0; fixed columns 0), total degree <= Shape::degree
void pk_set_gates(Context& ctx, ProvingKeyDev& pk, const GateCalc* calcs, size_t ncalcs, const Fr* constants, size_t nconstants,
                  const uint32_t* results, size_t nresults);

// wall-clock split of one create_proof call (seconds, stream synchronised at each boundary)
struct ProofTimings {
    double upload = 0, msm = 0, ntt = 0, lookup = 0, products = 0, quotient = 0, evals = 0, shplonk = 0, other = 0;
    // multi-GPU: wall time inside collectives (transfer + waiting for the slowest peer); already contained in the stages above
    double comm = 0;
};

std::unique_ptr<ProvingKeyDev> keygen(Context& ctx, const Shape& sh, const Fr* fixed_host, const uint32_t* copies, size_t ncopies);
void evaluate_h(Context& ctx, const ProvingKeyDev& pk, const Fr* advice_coeff, const Fr* z_coeff, const Fr* lookup_coeff, const Fr& y, const Fr& beta,
                const Fr& gamma, Fr* h_out);
std::vector<uint8_t> create_proof(Context& ctx, const ProvingKeyDev& pk, const Fr* advice_in, bool advice_on_device, host::FrRandomStream& rng,
                                  ProofTimings* tm);

}  // namespace b200zk
