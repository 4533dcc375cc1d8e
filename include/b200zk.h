/* libb200zk — C ABI of the B200-native KZG-BN254 proving backend.
 *
 * This is the drop-in boundary for the hot path of halo2_proofs::plonk::{keygen_pk, create_proof} that the
 * reference reaches through halo2-base `bench_builder` (reference: verifier/src/stark/mod.rs:543 and :593).
 * The reference has no FFI for this path (it is plain Rust generics in the un-vendored crates halo2-axiom /
 * halo2curves-axiom); each entry point below names the upstream function whose body a patched halo2_proofs
 * replaces with this call (INTEGRATION.md shows the Rust `-sys` binding).
 *
 * Conventions
 *   - `fr` / `fq` = uint64_t[4], little-endian limbs, Montgomery form (R = 2^256), canonical: byte-for-byte
 *     the in-memory layout of halo2curves::bn256::{Fr,Fq}. `g1_affine` = fq x, fq y (64 bytes); identity = (0,0).
 *   - Every function returns 0 on success and a negative B200ZK_E* code otherwise; nothing unwinds across the
 *     boundary. `b200zk_last_error(ctx)` returns a human-readable description of the last failure on ctx.
 *   - Host buffers are borrowed for the duration of the call only. Device memory lives behind the context.
 *   - Entry points ending in `_dev` take DEVICE pointers (same layouts) and run on the context's stream
 *     (`b200zk_stream`); they return after enqueueing unless stated otherwise.
 *   - Any OS thread may call; calls on one context are serialised by an internal mutex.
 *   - There is no CPU fallback: without a CUDA device `b200zk_create` fails with B200ZK_ENODEV.
 */
#ifndef B200ZK_H
#define B200ZK_H
#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define B200ZK_API __attribute__((visibility("default")))
#else
#define B200ZK_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200zk_ctx b200zk_ctx;
typedef struct { uint64_t l[4]; } b200zk_fr;
typedef struct { uint64_t l[4]; } b200zk_fq;
typedef struct { b200zk_fq x, y; } b200zk_g1_affine;

enum {
    B200ZK_OK = 0,
    B200ZK_ENODEV = -1,   /* no usable CUDA device */
    B200ZK_EINVAL = -2,   /* bad argument */
    B200ZK_ECUDA = -3,    /* CUDA runtime failure (see b200zk_last_error) */
    B200ZK_ESTATE = -4,   /* missing prerequisite (e.g. SRS not loaded) */
    B200ZK_ESYNTH = -5    /* constraint system failure (halo2 Error::ConstraintSystemFailure) */
};

/* ---- context ------------------------------------------------------------------------------------------- */
B200ZK_API int b200zk_create(int device, b200zk_ctx** out);
/* One process, several GPUs (`b200zk_create(const int* devices, int ndev)`): the returned context owns one
 * device context and one worker thread per listed device (distinct CUDA ordinals) and an NCCL communicator among them
 * (NCCL is taken from the process, libnccl.so.2). b200zk_srs_*, b200zk_keygen, b200zk_create_proof(_dev), b200zk_msm,
 * b200zk_msm_batch and b200zk_msm_bases called on it run on every device in lockstep — commit batches dealt by column with
 * the remainder split by point range, NTTs by column, h(X) by row slice, partial sums over NVLink — and return what a
 * single-GPU context returns, byte for byte. This is what a single test process such as the reference's
 * `bench_builder` caller (verifier/src/stark/mod.rs:543, :593) binds to use all GPUs of a box. Entry points that take device
 * pointers (`*_dev` MSM / NTT forms, dev_alloc, h2d …) and the column primitives address the first device alone.
 * ndev = 1 is the same as b200zk_create(devices[0]). */
B200ZK_API int b200zk_create_multi(const int* devices, int ndev, b200zk_ctx** out);
/* number of devices behind the context (1 for b200zk_create) */
B200ZK_API int b200zk_group_size(b200zk_ctx* ctx);
B200ZK_API int b200zk_destroy(b200zk_ctx* ctx);
B200ZK_API const char* b200zk_last_error(b200zk_ctx* ctx);
/* cudaStream_t of the context, for event timing by the caller */
B200ZK_API void* b200zk_stream(b200zk_ctx* ctx);
B200ZK_API int b200zk_sync(b200zk_ctx* ctx);
/* kernels launched by this library since process start (bench.py "gpu_launches") */
B200ZK_API unsigned long long b200zk_launch_count(void);
/* Multi-GPU with one process per GPU (for one process driving all GPUs see b200zk_create_multi): this
 * context is rank `rank` of `world`; all ranks must issue the same calls in the same order. Commit batches are dealt by
 * column with the remainder split by point range, create_proof shards its other stages too. `fn` must
 * all-gather `bytes` bytes from every rank into recv[world][bytes] (host buffers; bind it to an MPI / NCCL all-gather)
 * and return 0: the library uses it ONCE, to carry the 128-byte id of its own NCCL communicator (b200zk_comm_init, or the
 * first create_proof) — after that every exchange is an NCCL call on the library's stream. Without NCCL in the process the
 * communicator stays down and the MSM entry points exchange their partial sums (128 B each) through `fn` itself.
 * world = 1 disables sharding. */
typedef int (*b200zk_allgather_fn)(void* user, const void* send, size_t bytes, void* recv);
B200ZK_API int b200zk_set_allgather(b200zk_ctx* ctx, int rank, int world, b200zk_allgather_fn fn, void* user);
/* Brings up the library's own NCCL communicator now (collective: every rank must call it; the 128-byte id travels through
 * the callback above — its only use from then on). create_proof does this on first use; a host that only calls the MSM
 * entry points calls it once so that the partial sums go over NVLink inside the library instead of through the callback. */
B200ZK_API int b200zk_comm_init(b200zk_ctx* ctx);
/* Upstream details that change proof BYTES and that could not be checked against halo2-axiom's source (no Rust sources in
 * the build image; SURVEY.md §8c items 1–4). Each is one switch, mirrored by the CPU oracle; 0 / 0 = the defaults (classic
 * PSE-halo2 behaviour). If a proof ever differs from the Rust prover's, flipping a bit here is the fix:
 *   NO_UNUSED_BLIND_DRAWS : create_proof does not draw the `Blind` scalars that KZG commitments ignore (default: it does)
 *   LOOKUP_FILL_ASCENDING : permute_expression_pair assigns the ascending leftover table values to the repeated rows in
 *                           ascending order (default: to the repeated rows popped from the end)
 *   POINT_SIGN_BIT7       : compressed G1 points carry the y-sign in bit 7 and the identity flag in bit 6 (default: 6 / 7)
 *   random_poly_chunks    : vanishing::commit's random polynomial: 0 = n sequential Fr::random draws (default); T > 0 = T
 *                           worker chunks of n / T (plus one for a remainder), each filled from its own ChaCha20Rng whose
 *                           32-byte seed is drawn from the caller's stream (upstream's thread-count-dependent variant) */
enum { B200ZK_COMPAT_NO_UNUSED_BLIND_DRAWS = 1, B200ZK_COMPAT_LOOKUP_FILL_ASCENDING = 2, B200ZK_COMPAT_POINT_SIGN_BIT7 = 4 };
B200ZK_API int b200zk_set_compat(b200zk_ctx* ctx, uint32_t flags, uint32_t random_poly_chunks);
/* Precomputed window tables 2^(c·w)·P_i for the SRS bases (one bucket set per MSM, no host fold; costs W× SRS memory).
 * On by default; switch off before loading a large SRS to save memory. */
B200ZK_API int b200zk_set_msm_tables(b200zk_ctx* ctx, int on);
/* Batched-affine pre-reduction of dense MSM columns: `rounds` (0..6) rounds of pairwise affine additions with column-wide
 * batched inversions shrink every bucket's run before the XYZZ accumulation. Bit-exact; measured SLOWER than the XYZZ path
 * alone on B200 (≈ +8 ms per round and proof at k=20), so the default is 0. Kept as a tested alternative. */
B200ZK_API int b200zk_set_msm_affine_rounds(b200zk_ctx* ctx, int rounds);
/* per-kernel-family CUDA-event timing (off by default). ids: 0 msm_accumulate, 1 (reserved), 2 ntt_pass, 3 quotient.
 * profile_get synchronises, sums the spans recorded since the last reset and clears them. */
B200ZK_API int b200zk_profile_enable(b200zk_ctx* ctx, int on);
B200ZK_API int b200zk_profile_get(b200zk_ctx* ctx, int id, double* total_ms, unsigned long long* launches);
/* algorithmic work of the spans recorded since the last reset (call before profile_get, which clears them): mixed point
 * additions for id 0, butterflies for id 2, extended rows for id 3 */
B200ZK_API int b200zk_profile_work(b200zk_ctx* ctx, int id, double* units);
/* raw device memory helpers for hosts without their own allocator */
B200ZK_API int b200zk_dev_alloc(b200zk_ctx* ctx, size_t bytes, void** out);
B200ZK_API int b200zk_dev_free(b200zk_ctx* ctx, void* p);
B200ZK_API int b200zk_h2d(b200zk_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
B200ZK_API int b200zk_d2h(b200zk_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);

/* ---- row A: field arithmetic (halo2curves::bn256::{Fr,Fq}) — vector forms, used by the parity tests --------
 * field: 0 = Fr, 1 = Fq. op: 0 add, 1 sub, 2 mul, 3 inverse (0 -> 0), 4 neg, 5 from Montgomery (to_repr limbs),
 * 6 to Montgomery. Host buffers of n elements; b is ignored by unary ops. */
B200ZK_API int b200zk_field_vec_op(b200zk_ctx* ctx, int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n);
/* G1: op 0 = a + b (affine in, affine out), 1 = scalar[i]·a[i] (b = n fr scalars), 2 = 2·a */
B200ZK_API int b200zk_g1_vec_op(b200zk_ctx* ctx, int op, const b200zk_g1_affine* a, const void* b, b200zk_g1_affine* out, size_t n);

/* ---- row C: halo2_proofs::arithmetic::best_fft(a, omega, log_n) — natural order in and out, in place ----- */
B200ZK_API int b200zk_ntt(b200zk_ctx* ctx, b200zk_fr* a, uint32_t log_n, const b200zk_fr* omega);
B200ZK_API int b200zk_ntt_dev(b200zk_ctx* ctx, b200zk_fr* a_dev, uint32_t log_n, const b200zk_fr* omega);
/* `batch` columns of 2^log_n elements, `stride` elements apart (device) */
B200ZK_API int b200zk_ntt_batch_dev(b200zk_ctx* ctx, b200zk_fr* a_dev, uint32_t log_n, const b200zk_fr* omega, uint32_t batch, size_t stride);

/* ---- row D: EvaluationDomain::new(4, k) and its transforms ------------------------------------------------
 * lagrange_to_coeff: n -> n in place. coeff_to_extended: n -> 4n. extended_to_coeff: 4n -> 3n. */
B200ZK_API int b200zk_lagrange_to_coeff(b200zk_ctx* ctx, uint32_t k, b200zk_fr* a);
B200ZK_API int b200zk_coeff_to_extended(b200zk_ctx* ctx, uint32_t k, const b200zk_fr* in, b200zk_fr* out);
B200ZK_API int b200zk_extended_to_coeff(b200zk_ctx* ctx, uint32_t k, const b200zk_fr* in, b200zk_fr* out);
B200ZK_API int b200zk_lagrange_to_coeff_dev(b200zk_ctx* ctx, uint32_t k, b200zk_fr* a_dev, uint32_t batch, size_t stride);
B200ZK_API int b200zk_coeff_to_extended_dev(b200zk_ctx* ctx, uint32_t k, const b200zk_fr* in_dev, b200zk_fr* out_dev, uint32_t batch,
                                 size_t stride_in, size_t stride_out);
B200ZK_API int b200zk_extended_to_coeff_dev(b200zk_ctx* ctx, uint32_t k, const b200zk_fr* in_dev, b200zk_fr* out_dev);

/* ---- rows B, K: ParamsKZG<Bn256> bases and halo2curves::msm::best_multiexp ------------------------------------
 * srs_load copies both bases (n = 2^k affine points each) to the device once; `basis` below: 0 = g (monomial,
 * ParamsKZG::commit), 1 = g_lagrange (ParamsKZG::commit_lagrange). Results are canonical affine points. */
B200ZK_API int b200zk_srs_load(b200zk_ctx* ctx, uint32_t k, const b200zk_g1_affine* g, const b200zk_g1_affine* g_lagrange);
/* ParamsKZG::setup(k, rng) with s = Fr::random(rng) drawn from ChaCha20Rng::from_seed(seed) (halo2-base gen_srs uses
 * seed = [0;32]), or from an explicit trapdoor; both bases are generated on the device. For tests and benches only —
 * a production SRS comes from a ceremony through b200zk_srs_load. `trapdoor_out` (optional) receives s. */
B200ZK_API int b200zk_srs_setup(b200zk_ctx* ctx, uint32_t k, const uint8_t seed[32], b200zk_fr* trapdoor_out);
B200ZK_API int b200zk_srs_setup_trapdoor(b200zk_ctx* ctx, uint32_t k, const b200zk_fr* s);
/* ParamsKZG::{write, read} byte layout: k (u32 LE) | n G1 (g) | n G1 (g_lagrange) | G2 g2 | G2 s_g2.
 * format 0 = SerdeFormat::RawBytes (64-byte G1 / 128-byte G2, raw Montgomery limbs), 1 = Processed (32-byte compressed
 * G1; the two 64-byte G2 encodings are carried as opaque bytes). read validates every G1 point on the device. write needs
 * the G2 bytes in the same format (available after srs_setup for RawBytes, or after a read of that format). */
B200ZK_API size_t b200zk_srs_file_size(uint32_t k, int format);
B200ZK_API int b200zk_srs_read(b200zk_ctx* ctx, const uint8_t* data, size_t len, int format);
B200ZK_API int b200zk_srs_write(b200zk_ctx* ctx, int format, uint8_t* out, size_t capacity, size_t* written);
/* copy the bases back (either may be NULL) */
B200ZK_API int b200zk_srs_download(b200zk_ctx* ctx, b200zk_g1_affine* g, b200zk_g1_affine* g_lagrange);
B200ZK_API int b200zk_msm(b200zk_ctx* ctx, int basis, const b200zk_fr* scalars, size_t n, b200zk_g1_affine* out);
B200ZK_API int b200zk_msm_dev(b200zk_ctx* ctx, int basis, const b200zk_fr* scalars_dev, size_t n, b200zk_g1_affine* out);
/* `ncols` commitments over the same basis: cols_dev[i] points at n device-resident scalars; out gets ncols affine points.
 * Buckets are accumulated per column and reduced once for the whole batch (the reduction tail is latency bound). */
B200ZK_API int b200zk_msm_batch_dev(b200zk_ctx* ctx, int basis, const b200zk_fr* const* cols_dev, size_t ncols, size_t n, b200zk_g1_affine* out);
/* the same with HOST columns (cols[i] points at n scalars in host memory): what a patched ParamsKZG::commit /
 * commit_lagrange loop over several polynomials binds (SURVEY.md §8b `b200zk_msm_batch`). Uploads are pipelined with
 * the accumulation of the previous column. */
B200ZK_API int b200zk_msm_batch(b200zk_ctx* ctx, int basis, const b200zk_fr* const* cols, size_t ncols, size_t n, b200zk_g1_affine* out);
/* best_multiexp(coeffs, bases) with caller-supplied bases (host buffers) */
B200ZK_API int b200zk_msm_bases(b200zk_ctx* ctx, const b200zk_g1_affine* bases, const b200zk_fr* scalars, size_t n, b200zk_g1_affine* out);
/* device bases + device scalars (bench / multi-GPU shards): point range [0, n) of bases_dev */
B200ZK_API int b200zk_msm_bases_dev(b200zk_ctx* ctx, const b200zk_g1_affine* bases_dev, const b200zk_fr* scalars_dev, size_t n,
                                    b200zk_g1_affine* out);

/* ---- rows E, F, I: column primitives used inside create_proof, exported for hosts that keep their own prover loop ------
 * (host buffers; each call stages through device memory)
 * batch_invert      : halo2 `batch_invert` — a[i] <- a[i]^-1, zeros stay zero.
 * prefix_product    : z[0] = first, z[i] = z[i-1]·m[i-1] (the running product of permutation::prover / lookup::prover).
 * eval_polynomial   : arithmetic::eval_polynomial(poly, point).
 * kate_division     : arithmetic::kate_division(a, b): quotient of a(X) by (X - b), n-1 coefficients.
 * permute_expression_pair : lookup::prover::permute_expression_pair on the first n-7 rows, for range-style tables: every
 *                     table value must be < n = 2^k (B200ZK_EINVAL otherwise — an unsupported table, not a verdict on
 *                     the witness); B200ZK_ESYNTH when an input is missing from the table. n-7 rows out. */
B200ZK_API int b200zk_batch_invert(b200zk_ctx* ctx, b200zk_fr* a, size_t n);
B200ZK_API int b200zk_prefix_product(b200zk_ctx* ctx, const b200zk_fr* m, const b200zk_fr* first, b200zk_fr* z, size_t n);
B200ZK_API int b200zk_eval_polynomial(b200zk_ctx* ctx, const b200zk_fr* poly, size_t n, const b200zk_fr* point, b200zk_fr* out);
B200ZK_API int b200zk_kate_division(b200zk_ctx* ctx, const b200zk_fr* a, size_t n, const b200zk_fr* b, b200zk_fr* q);
B200ZK_API int b200zk_permute_expression_pair(b200zk_ctx* ctx, uint32_t k, const b200zk_fr* input, const b200zk_fr* table, b200zk_fr* a_out,
                                              b200zk_fr* s_out);

/* ---- rows J, E–I: plonk::{keygen_vk, keygen_pk, create_proof} for the halo2-base ConstraintSystem -----------------
 * Shape: A gate advice columns (one `q·(a + b·c − d)` gate each, rotations 0..3), L lookup
 * advice columns against one table column, F constant columns. Fixed column order: F constants, table, A selectors.
 * Permutation column order (indices used by `copies`): F constants, A gate columns, L lookup columns.
 * `fixed`: num_fixed × 2^k elements, column-major; `copies`: ncopies × {col_a,row_a,col_b,row_b}.
 * Requires an SRS of the same k on the context. The key lives on the device until b200zk_pk_free. */
typedef struct b200zk_pk b200zk_pk;
B200ZK_API int b200zk_keygen(b200zk_ctx* ctx, uint32_t k, uint32_t A, uint32_t L, uint32_t F, const b200zk_fr* fixed, const uint32_t* copies,
                             size_t ncopies, b200zk_pk** out);
B200ZK_API int b200zk_pk_free(b200zk_ctx* ctx, b200zk_pk* pk);
/* Generic custom gates (halo2_proofs::plonk::evaluation::{GraphEvaluator, Calculation, ValueSource}):
 * calculation j produces intermediate j from its value sources; the intermediates listed in `results` are the gate
 * polynomials, folded into h(X) by Horner in the challenge y in list order. Installing a program replaces the built-in
 * halo2-base gates (one `q·(a + b·c − d)` per gate column) in evaluate_h / create_proof for this key; ncalcs = 0 restores
 * them. Column indices are those of b200zk_keygen (fixed: F constants, table, A selectors; advice: A gate, L lookup columns).
 * The proof's query set is unchanged, so advice sources may use rotations 0..3 on gate columns and 0 on lookup columns,
 * fixed sources rotation 0; the total degree of every gate must be <= 4 (the constraint system's degree). Limits: 128
 * calculations, 64 constants, 64 gates. The permutation and lookup arguments are unaffected. */
enum { B200ZK_SRC_CONSTANT = 0, B200ZK_SRC_INTERMEDIATE = 1, B200ZK_SRC_FIXED = 2, B200ZK_SRC_ADVICE = 3 };
enum { B200ZK_CALC_ADD = 0, B200ZK_CALC_SUB = 1, B200ZK_CALC_MUL = 2, B200ZK_CALC_SQUARE = 3, B200ZK_CALC_DOUBLE = 4,
       B200ZK_CALC_NEGATE = 5, B200ZK_CALC_STORE = 6 };
typedef struct { uint32_t kind; uint32_t index; int32_t rotation; } b200zk_value_source;
typedef struct { uint32_t op; b200zk_value_source a, b; } b200zk_calculation;  /* b is ignored by unary calculations */
B200ZK_API int b200zk_pk_set_gates(b200zk_ctx* ctx, b200zk_pk* pk, const b200zk_calculation* calcs, size_t ncalcs,
                                   const b200zk_fr* constants, size_t nconstants, const uint32_t* results, size_t nresults);
/* vk.fixed_commitments (num_fixed points) and vk.permutation.commitments (num_perm points); either may be NULL */
B200ZK_API int b200zk_pk_commitments(b200zk_ctx* ctx, const b200zk_pk* pk, b200zk_g1_affine* fixed_out, b200zk_g1_affine* perm_out);
/* vk.transcript_repr: read, or override with the value the Rust side computed (opaque 32-byte scalar) */
B200ZK_API int b200zk_pk_transcript_repr(b200zk_ctx* ctx, b200zk_pk* pk, b200zk_fr* get_out, const b200zk_fr* set_in);
/* test access to key columns: which 0 = sigma values j (n), 1 = fixed coset i (4n), 2 = sigma coset j (4n),
 * 3 = l0 / l_last / l_active_row (idx 0..2, 4n) */
B200ZK_API int b200zk_pk_get_column(b200zk_ctx* ctx, const b200zk_pk* pk, int which, uint32_t idx, b200zk_fr* out);
B200ZK_API size_t b200zk_proof_size(uint32_t k, uint32_t A, uint32_t L, uint32_t F);
/* create_proof::<KZGCommitmentScheme<Bn256>, ProverSHPLONK, Challenge255, StdRng, Blake2bWrite> for ONE circuit without
 * instances. `advice`: (A+L) × 2^k witness values, column-major (rows >= 2^k − 7 are overwritten by blinding).
 * rng = StdRng::seed_from_u64(rng_seed). `proof_out` must hold b200zk_proof_size bytes. `timings` (optional, 10 doubles):
 * seconds spent in upload, msm, ntt, lookup, products, quotient, evals, shplonk, other, and — multi-GPU, already contained in
 * the stages before it — inside collectives (transfer plus waiting for the slowest peer). */
B200ZK_API int b200zk_create_proof(b200zk_ctx* ctx, const b200zk_pk* pk, const b200zk_fr* advice, uint64_t rng_seed, uint8_t* proof_out,
                                   size_t* proof_len, double* timings);
/* The same for ANY `RngCore` the host holds (create_proof is generic in `R: RngCore`; halo2-base passes StdRng, the seeded
 * entry point above): `fill` is the generator's `fill_bytes` — it must write `nbytes` random bytes and return 0. Every draw,
 * including the 2^k Fr::random draws of the random polynomial (64 bytes each), is pulled through it in upstream's order, so
 * the proof equals the one the Rust prover makes with that generator. Single-GPU contexts only. */
typedef int (*b200zk_rng_fill_fn)(void* user, uint8_t* out, size_t nbytes);
B200ZK_API int b200zk_create_proof_rng(b200zk_ctx* ctx, const b200zk_pk* pk, const b200zk_fr* advice, b200zk_rng_fill_fn fill, void* user,
                                       uint8_t* proof_out, size_t* proof_len, double* timings);
/* same with the advice columns already resident in device memory (the witness upload excluded) */
B200ZK_API int b200zk_create_proof_dev(b200zk_ctx* ctx, const b200zk_pk* pk, const b200zk_fr* advice_dev, uint64_t rng_seed, uint8_t* proof_out,
                                       size_t* proof_len, double* timings);
/* Row H on its own — evaluation::Evaluator::evaluate_h + EvaluationDomain::divide_by_vanishing_poly — for a host that
 * keeps halo2's create_proof loop and replaces only this step (SURVEY.md §8b `b200zk_evaluate_h`). All inputs are host
 * buffers in COEFFICIENT form, as halo2 holds them at that point:
 *   advice_coeff : (A+L) × 2^k,   perm_z_coeff : num_sets × 2^k (permutation product polynomials),
 *   lookup_coeff : L × 3 × 2^k, per lookup Z, a' (permuted input), s' (permuted table); may be NULL when L = 0.
 * y, beta, gamma: the transcript challenges (theta is not needed: halo2-base lookups are single-column).
 * h_ext_out: 4·2^k values of h on the extended coset domain, already divided by the vanishing polynomial, i.e. the
 * input of extended_to_coeff. Runs on this context's GPU alone, also inside a multi-rank job. */
B200ZK_API int b200zk_evaluate_h(b200zk_ctx* ctx, const b200zk_pk* pk, const b200zk_fr* advice_coeff, const b200zk_fr* perm_z_coeff,
                                 const b200zk_fr* lookup_coeff, const b200zk_fr* y, const b200zk_fr* beta, const b200zk_fr* gamma,
                                 b200zk_fr* h_ext_out);
/* number of permutation product polynomials (sets of degree − 2 = 2 columns) for a shape */
B200ZK_API uint32_t b200zk_num_sets(uint32_t A, uint32_t L, uint32_t F);
/* ---- host-only helpers (no CUDA device needed) -------------------------------------------------------------------
 * Sum of n affine G1 points (canonical affine out): combines per-GPU partial MSM results. */
B200ZK_API int b200zk_g1_sum_host(const b200zk_g1_affine* points, size_t n, b200zk_g1_affine* out);
/* Cross-checks the host paths of the shared field/curve code (carry-chain emulation vs 64-bit limbs, group laws) on
 * `iters` pseudo-random inputs; returns 0 when every identity holds. */
B200ZK_API int b200zk_host_selftest(uint64_t seed, size_t iters);

#ifdef __cplusplus
}
#endif
#endif
