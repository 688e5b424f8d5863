/* b200_msm.h — C-ABI of the B200-native MSM / fixed-base batch_exp engine that
 * drops in behind libff's scalar_multiplication templates as used by LegoSNARK,
 * plus the neighbours of that path: the knowledge-commitment MSM,
 * the Fr vector work of CPPoly / sum-check tables, libfqfft's radix-2 transform and
 * the point compression of the stream operators.
 *
 * The reference has no FFI: its boundary is a set of C++ function templates in
 *   LFF/algebra/scalar_multiplication/multiexp.hpp
 * (LFF = depends/libsnark/depends/libff/libff) instantiated in the caller's TU.
 * legosnark_b200/shim/libff/algebra/scalar_multiplication/multiexp.hpp shadows
 * that header and forwards the four concrete groups (alt_bn128 / bn128 G1, G2)
 * to the entry points below; INTEGRATION.md shows the binding.
 *
 * Data layout (identical to the reference's in-memory objects, so callers pass
 * &vec[0] reinterpret-cast, no repacking):
 *   Fr / Fq element : 4 x u64 little-endian limbs, Montgomery form, R = 2^256
 *                     (fp.hpp:42 mont_repr; bn::Fp, ATE/include/zm2.h:266)
 *   G1 point        : X | Y | Z             = 12 limbs = 96 B, Jacobian, Z == 0 <=> zero
 *                     (alt_bn128_g1.hpp:35; bn128_g1.hpp:36 coord[3])
 *   G2 point        : X.c0 X.c1 | Y.c0 Y.c1 | Z.c0 Z.c1 = 24 limbs = 192 B
 *                     (alt_bn128_g2.hpp:36; bn128_g2.hpp coord[3] of Fp2T{a_,b_})
 * Bases may be zero, repeated, and non-normalised Jacobian (Z != 1).
 * Point outputs are normalised: (x, y, 1) in Montgomery form, or the zero
 * (0, 1, 0) of alt_bn128 (the shim rewrites it to bn128's (1,1,0)).  Results
 * equal the reference's as group elements and bit-for-bit after
 * to_affine_coordinates().
 *
 * Every function returns 0 on success; otherwise b200_last_error() describes the
 * failure.  There is NO CPU fallback: without a CUDA device b200_init fails and
 * every compute entry point returns an error.  Calls are serialised per process
 * by an internal mutex (every shipped caller is single-threaded).
 */
#ifndef B200_MSM_H
#define B200_MSM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_ERR_CUDA 1
#define B200_ERR_ARG 2
#define B200_ERR_NOT_INIT 3
#define B200_ERR_NO_DEVICE 4

/* ---- lifecycle ------------------------------------------------------------- */
/* Use devices 0..n_gpus-1 (n_gpus <= 0: all visible).  An MSM / batch_exp call is
 * sharded by index range across them (multiexp.tcc:417-438 does the same across
 * OpenMP threads) and the per-GPU partials are summed on the host. */
int b200_init(int n_gpus);
/* The same, started on a background thread: returns at once; every later call (b200_init included) waits for it to finish and
 * reports its error, if any.  The CUDA driver's own start-up is 0.6-2.3 s per process on these boxes; a prover that calls this
 * first thing in main() hides it under its circuit / witness construction.  Idempotent. */
int b200_init_async(int n_gpus);
/* Use exactly these CUDA device ordinals (one process per GPU under torchrun: {LOCAL_RANK}). */
int b200_init_devices(const int *device_ids, int n);
void b200_shutdown(void);
int b200_device_count(void); /* devices in use; 0 before init */
const char *b200_last_error(void);
const char *b200_version(void);

/* ---- multi_exp / multi_exp_with_mixed_addition  (multiexp.hpp:56-61, 70-75;
 *      multiexp.tcc:402-441, 443-496, inner BDLO12 :165-282) ---------------------
 * out = sum_i scalars[i] * bases[i].  `Method` and `chunks` of the reference are
 * hints that do not change the group element; every variant maps here. */
int b200_msm_g1(const uint64_t *bases /* n x 12 */, const uint64_t *scalars_mont /* n x 4 */, size_t n,
                uint64_t out[12]);
int b200_msm_g2(const uint64_t *bases /* n x 24 */, const uint64_t *scalars_mont /* n x 4 */, size_t n,
                uint64_t out[24]);

/* knowledge_commitment<G2, G1> MSM: Groth16's B query (SNK/knowledge_commitment/kc_multiexp.tcc:21-89
 * -> multi_exp<knowledge_commitment<T1,T2>>, called from r1cs_gg_ppzksnark.tcc:453-463; SNK =
 * depends/libsnark/libsnark).  out_g2 = sum_i s_i g_i and out_g1 = sum_i s_i h_i over ONE scalar
 * vector, which is uploaded once and stays on the devices for the second sum. */
int b200_msm_g2g1(const uint64_t *g2_bases /* n x 24 */, const uint64_t *g1_bases /* n x 12 */,
                  const uint64_t *scalars_mont /* n x 4 */, size_t n, uint64_t out_g2[24], uint64_t out_g1[12]);

/* Many small MSMs in ONE call: out[j] = sum_{i in [offsets[j], offsets[j+1])} scalars[i] * bases[i], j < count.
 * Replaces the loop of tiny multi_exp calls behind LegoSNARK's sparse-matrix keygen: mtxmultiexp
 * (LS/gadgets/subspace.cc:18-25) calls simplesparsemexp -> sparsemexpG -> multi_exp<BDLO12>
 * (LS/utils/sparsemexp.h:62-90, sparsemexp.cc:15-24) once per matrix column, 2 050 columns of one or two terms for
 * the shipped cplink example.  One thread per term runs a 4-bit windowed scalar multiplication, one thread per
 * MSM sums its terms, the outputs are normalised together (one shared inversion per thread block of outputs);
 * the host sees one upload, one kernel chain and one download instead of `count` round trips.
 * offsets: count + 1 non-decreasing indices, offsets[0] = 0; outputs are normalised like b200_msm_*. */
int b200_msm_batch_g1(const uint64_t *bases /* offsets[count] x 12 */, const uint64_t *scalars_mont /* offsets[count] x 4 */,
                      const uint64_t *offsets /* count + 1 */, size_t count, uint64_t *out /* count x 12 */);
int b200_msm_batch_g2(const uint64_t *bases /* offsets[count] x 24 */, const uint64_t *scalars_mont,
                      const uint64_t *offsets, size_t count, uint64_t *out /* count x 24 */);

/* Host-side sum of partial results (north star: "each GPU returns a partial group element,
 * and the partials are summed on the host"; the serial sum at multiexp.tcc:433-438).  Used by
 * the engine for its own per-device partials and by one-process-per-GPU launchers (bench.py
 * under torchrun) for the per-rank partials.  Pure host code: needs no device. */
int b200_sum_partials_g1(const uint64_t *pts /* n x 12, any Jacobian */, size_t n, uint64_t out[12]);
int b200_sum_partials_g2(const uint64_t *pts /* n x 24 */, size_t n, uint64_t out[24]);

/* ---- resident commitment keys ------------------------------------------------
 * LegoSNARK commits many vectors under one key (CommScheme::commit,
 * LS/prototools/commit.h:149-158; CPPoly::prove uses prefixes of the same g1s,
 * LS/gadgets/poly.h:77-88).  pin uploads + normalises the bases once (sharded
 * over the devices in use); msm_pinned then only moves scalars.  offset/n select
 * the sub-range [offset, offset+n) of the key. */
int b200_pin_bases_g1(const uint64_t *bases, size_t n, uint64_t *handle);
int b200_pin_bases_g2(const uint64_t *bases, size_t n, uint64_t *handle);
int b200_unpin_bases(uint64_t handle);
int b200_msm_pinned_g1(uint64_t handle, size_t offset, const uint64_t *scalars_mont, size_t n, uint64_t out[12]);
int b200_msm_pinned_g2(uint64_t handle, size_t offset, const uint64_t *scalars_mont, size_t n, uint64_t out[24]);

/* Precomputed key: extends a pinned key IN PLACE by the window multiples 2^(c k) P_i,
 * k = 1 .. ceil(255/c)-1, as affine points in HBM (W x 64 B per G1 base: 0.9 GB at 2^20 with
 * c = 20, 56 GB at 2^26 -- the B200's 180 GB make this affordable).  Later msm_pinned calls on
 * the key then sort the digits of all W windows into ONE set of 2^(c-1) buckets (the digit of
 * window k selects level k of base i), which removes the per-window bucket reduction and the
 * Horner doublings and lets c grow (fewer windows = fewer point additions).  One-off cost:
 * (W-1) c doublings per base; amortised over the many commitments LegoSNARK makes under one
 * key (commit.h:149-158).  window_bits 0 = chosen for the key length.  Sub-range MSMs too short
 * to fill the buckets keep using the plain path.  Results are the same group elements. */
int b200_key_precompute_g1(uint64_t handle, uint32_t window_bits);
int b200_key_precompute_g2(uint64_t handle, uint32_t window_bits);

/* Scalars already in device memory (cudaMalloc / torch) on the handle's first
 * device; all kernels are enqueued on `cuda_stream` (a cudaStream_t, NULL = the
 * engine's own stream) so the caller can bracket the call with its own events.
 * Single-device only.  Used by bench.py for the HBM-resident `value` number. */
int b200_msm_pinned_dev_g1(uint64_t handle, size_t offset, const void *d_scalars_mont, size_t n, void *cuda_stream,
                           uint64_t out[12]);
int b200_msm_pinned_dev_g2(uint64_t handle, size_t offset, const void *d_scalars_mont, size_t n, void *cuda_stream,
                           uint64_t out[24]);

/* ---- fixed-base: get_window_table + batch_exp / batch_exp_with_coeff
 *      (multiexp.hpp:96-124; multiexp.tcc:509-681) --------------------------------
 * out[i] = (coeff * scalars[i]) * base   (coeff_mont == NULL: no coefficient).
 * The reference's window_table layout is not observable by any caller
 * (SURVEY.md §8b); the engine builds its own affine table on the device with a
 * window chosen for the GPU.  Outputs are normalised (batch_to_special form). */
size_t b200_exp_window_size_g1(size_t num_scalars); /* libff's table: alt_bn128_init.cpp:157-201 */
size_t b200_exp_window_size_g2(size_t num_scalars); /* alt_bn128_init.cpp:220-264 */
int b200_batch_exp_g1(const uint64_t base[12], const uint64_t *scalars_mont, size_t n, const uint64_t *coeff_mont,
                      uint64_t *out /* n x 12 */);
int b200_batch_exp_g2(const uint64_t base[24], const uint64_t *scalars_mont, size_t n, const uint64_t *coeff_mont,
                      uint64_t *out /* n x 24 */);
/* table reuse across calls (r1cs_gg_ppzksnark.tcc:296-360 builds one G1 and one G2 table
 * and runs five batch_exps over them) */
int b200_window_table_create_g1(const uint64_t base[12], size_t expected_scalars, uint64_t *handle);
int b200_window_table_create_g2(const uint64_t base[24], size_t expected_scalars, uint64_t *handle);
int b200_window_table_destroy(uint64_t handle);
int b200_batch_exp_table_g1(uint64_t handle, const uint64_t *scalars_mont, size_t n, const uint64_t *coeff_mont,
                            uint64_t *out);
int b200_batch_exp_table_g2(uint64_t handle, const uint64_t *scalars_mont, size_t n, const uint64_t *coeff_mont,
                            uint64_t *out);
/* device-resident variant: scalars in, affine points (x|y, 8 resp. 16 limbs each, (0,0) = zero)
 * out, both in device memory; enqueued on cuda_stream.  Doubles as the on-device generator of
 * benchmark bases k_i * G. */
int b200_batch_exp_table_dev_g1(uint64_t handle, const void *d_scalars_mont, size_t n, void *d_out_affine,
                                void *cuda_stream);
int b200_batch_exp_table_dev_g2(uint64_t handle, const void *d_scalars_mont, size_t n, void *d_out_affine,
                                void *cuda_stream);
/* pin bases that already sit in device memory in the affine layout above (no copy is taken of
 * host data; the engine copies device-to-device) */
int b200_pin_affine_dev_g1(const void *d_affine, size_t n, uint64_t *handle);
int b200_pin_affine_dev_g2(const void *d_affine, size_t n, uint64_t *handle);

/* ---- batch_to_special (multiexp.hpp:126-127; multiexp.tcc:683-715;
 *      alt_bn128_g1.cpp:499-521; field_utils.tcc:171-194) ----------------------------
 * In place: every non-zero point becomes (X/Z^2, Y/Z^3, 1); zeros become (0,1,0). */
int b200_batch_to_affine_g1(uint64_t *pts, size_t n);
int b200_batch_to_affine_g2(uint64_t *pts, size_t n);

/* ---- Fr vector work on either side of the MSMs -----------
 * All vectors hold Montgomery-form Fr elements, 4 limbs each, as the reference's
 * std::vector<Fr> does.  LS = src/, FQFFT = depends/libsnark/depends/libfqfft/libfqfft/.
 *
 * Witness folding of CPPoly::prove (LS/gadgets/poly.h:45-67): v has 2^d values, r has d.
 * Level i pairs (2p, 2p+1): w_coeffs[start_i + p] = -t[2p] + t[2p+1] and
 * t[p] <- -t[2p] (r_i - 1) + t[2p+1] r_i, start_i = 2^d - 2^(d-i).  w_coeffs gets the
 * reference's 2^d-entry vector (last entry zero); eval gets the last t[0], which is
 * MultiVPolyT::evalMLE(v, r).  Either output may be NULL. */
int b200_fr_fold_witness(const uint64_t *v, const uint64_t *r, size_t d, uint64_t *w_coeffs /* 2^d x 4 */, uint64_t eval[4]);
/* MultiVPolyT::evalMLE (LS/prototools/polytools.h:207-234): sum_p v[p] prod_i (bit i of p ? r_i : 1 - r_i). */
int b200_fr_eval_mle(const uint64_t *v /* 2^d x 4 */, const uint64_t *r /* d x 4 */, size_t d, uint64_t out[4]);
/* DPMle::pushRandomness (LS/prototools/mle.h:199-210): out[p] = table[p] (1 - r) + table[p + half] r. */
int b200_fr_mle_bind(const uint64_t *table /* 2 half x 4 */, size_t half, const uint64_t r[4], uint64_t *out /* half x 4 */);
/* step_radix2_domain<Fr> transforms over 2^log_big + 2^log_small points (log_small < log_big), in place on a host vector:
 * mode 0 FFT, 1 iFFT, 2 cosetFFT(g), 3 icosetFFT(g) (FQFFT/evaluation_domain/domains/step_radix2_domain.tcc:38-152).  The two
 * radix-2 transforms inside and the O(m) loops the reference wraps around them (serial omega_i *= omega chains on the
 * host there) run on the device; one upload and one download per call. */
int b200_fr_step_fft(uint64_t *a /* (2^log_big + 2^log_small) x 4 */, size_t log_big, size_t log_small, int mode, const uint64_t *coset_g);

/* The vector part of libsnark's r1cs_to_qap_witness_map (SNK/reductions/r1cs_to_qap/r1cs_to_qap.tcc:232-311, SNK =
 * depends/libsnark/libsnark) for d1 = d2 = d3 = 0, as every Groth16 prover calls it (r1cs_gg_ppzksnark.tcc:402-415): from the
 * evaluations aA, aB, aC of the constraint polynomials on the domain to the coefficients of H = (A B - C) / Z --
 * three iFFTs, three coset FFTs, the pointwise product and difference, divide_by_Z_on_coset and the inverse coset FFT --
 * with ONE upload of the three vectors and one download; nothing returns to the host between the seven transforms.
 * Domain: log_small == B200_QAP_BASIC: basic_radix2_domain of m = 2^log_big points, div_consts = {Z(g)^-1} (4 limbs);
 * otherwise step_radix2_domain of m = 2^log_big + 2^log_small points, div_consts = {c1, ratio, c0, Z1^-1} (16 limbs) as in
 * b200_fr_scale_inv_geometric.  H receives m elements (the caller appends coefficients_for_H[m] = 0). */
#define B200_QAP_BASIC ((size_t)-1)
int b200_qap_h_coefficients(const uint64_t *aA /* m x 4 */, const uint64_t *aB /* m x 4 */, const uint64_t *aC /* m x 4 */, size_t log_big,
                            size_t log_small, const uint64_t coset_g[4], const uint64_t *div_consts, uint64_t *H /* m x 4 */);

/* step_radix2_domain::divide_by_Z_on_coset (FQFFT/evaluation_domain/domains/step_radix2_domain.tcc:213-241; FQFFT =
 * depends/libsnark/depends/libfqfft/libfqfft): the domain libfqfft picks for 2^k + 2^r constraints (the 128 x 128 matrix
 * product of BASELINE.json configs[3]: 2^21 + 1) divides by Z with ONE FIELD INVERSION PER POINT on the host, 95 % of the
 * Groth16 prover once the MSMs and FFTs are on the device.  In place: P[i] *= (c1 * ratio^i - c0)^-1 for i < n_geo, and
 * P[n_geo + i] *= tail for i < n_tail (tail may be NULL when n_tail == 0); the caller forms the four constants
 * (shim/libfqfft/.../step_radix2_domain.hpp).  Batched inversion on the device; inverses are unique, so the limbs agree. */
int b200_fr_scale_inv_geometric(uint64_t *P /* (n_geo + n_tail) x 4 */, size_t n_geo, const uint64_t c1[4], const uint64_t ratio[4],
                                const uint64_t c0[4], size_t n_tail, const uint64_t *tail /* 4 */);

/* The Lagrange-coefficient vectors of libfqfft's radix-2 and step domains: _basic_radix2_evaluate_all_lagrange_polynomials
 * (FQFFT/evaluation_domain/domains/basic_radix2_domain_aux.tcc:183-236: u[i] = l * (t - r).inverse(); l *= omega; r *= omega) and
 * step_radix2_domain::evaluate_all_lagrange_polynomials (step_radix2_domain.tcc:161-186), which libsnark's generators call through
 * r1cs_to_qap_instance_map_with_evaluation (SNK/reductions/r1cs_to_qap/r1cs_to_qap.tcc:127-190).  Both spend ONE FIELD INVERSION
 * PER POINT on the host (5.4 of the 5.7 host seconds of the Groth16 generator for the 128 x 128 matrix product of BASELINE.json
 * configs[3]).  out[i] = in[i] * a0 * a_ratio^i / prod_f (c1_f * ratio_f^i - c0_f) for i < n and f < n_factors (1 or 2);
 * in == NULL means in[i] = 1, in == out is allowed.  consts = a0, a_ratio, then (c1, ratio, c0) per factor: (2 + 3 n_factors) x 4
 * limbs, formed by the caller (shim/libfqfft/.../basic_radix2_domain_aux.hpp, step_radix2_domain.hpp).  No denominator may be
 * zero (the reference tests t^m == 1 first and so do the shadows).  Batched inversion on the device; same limbs. */
int b200_fr_geometric_quotients(uint64_t *out /* n x 4 */, const uint64_t *in /* n x 4 or NULL */, size_t n, const uint64_t *consts,
                                size_t n_factors);

/* Sum-check dynamic-programming tables (LS/prototools/mle.h, LS/gadgets/sumcheck.h; all values Montgomery-form Fr).
 * b200_fr_eq_table: DPBeta::compute_eq_tbl (mle.h:93-105), level by level exactly as written there:
 *   T_0 = {1};  T_{j+1}[p] = eqbit(p >= 2^j, r[j]) * T_j[p >> 1], p < 2^(j+1);  out = T_d.
 * b200_fr_matrix_mle: DPMatrixMle's constructor (mle.h:241-259): v[r] = sum_l A[(l << d) + r] * eq[l] with
 *   eq = the table above for rho; A is the vectorised 2^d x 2^d matrix.
 * b200_fr_sumcheck_round: the sum over p inside CPSumcheck::make_new_h_poly (sumcheck.h:85-106) for two tables:
 *   out = the coefficients (c0, c1, c2) of  sum_{p < half} w[p] (a[p](1-x) + a[p+half] x)(b[p](1-x) + b[p+half] x);
 *   w = the round's beta suffix values (DPBeta::getBetaSuff) or NULL for DPBetaDummy (w = 1); the caller
 *   multiplies by eqbit_poly(rho[j]) * beta_pre (getBetaPoly, mle.h:78-84) when there is a beta.
 * b200_fr_sumcheck_rounds: all d rounds of the beta-less sum-check (CPSumcheckMatrix, sumcheck.h:118-131; the
 *   loop of CPSumcheck::prove, sumcheck.cc:56-70): h[3 i .. 3 i + 2] = round i's coefficients; between rounds both
 *   tables are bound to r[i] on the device (DPMle::pushRandomness, mle.h:199-210). */
int b200_fr_eq_table(const uint64_t *r /* d x 4 */, size_t d, uint64_t *out /* 2^d x 4 */);
int b200_fr_matrix_mle(const uint64_t *A /* 4^d x 4 */, const uint64_t *rho /* d x 4 */, size_t d, uint64_t *v /* 2^d x 4 */);
int b200_fr_sumcheck_round(const uint64_t *a /* 2 half x 4 */, const uint64_t *b /* 2 half x 4 */, const uint64_t *w /* half x 4 or NULL */,
                           size_t half, uint64_t out[12]);
int b200_fr_sumcheck_rounds(const uint64_t *a /* 2^d x 4 */, const uint64_t *b /* 2^d x 4 */, const uint64_t *r /* d x 4 */, size_t d,
                            uint64_t *h /* d x 12 */);
/* CPPoly::prove (poly.h:45-91) against a resident G1 key (b200_pin_bases_g1, single device):
 * folding on the device, then witness[i] = multiExpMA(g1s, w_i) over the first 2^(d-i-1)
 * bases with the scalars never leaving the device.  witness: d normalised points; the
 * reference's witnessa[i] (i >= 1) repeats the same MSM over the same bases (poly.h:84-86) and
 * equals witness[i].  eval (may be NULL) = evalMLE(v, r). */
int b200_cppoly_prove_g1(uint64_t key_handle, const uint64_t *v, const uint64_t *r, size_t d, uint64_t *witness /* d x 12 */,
                         uint64_t eval[4]);
/* libfqfft basic_radix2_domain<Fr> (FQFFT/evaluation_domain/domains/basic_radix2_domain.tcc,
 * _aux.tcc:42-75) on m = 2^log_n values in place.  mode 0: FFT, 1: iFFT (times m^-1),
 * 2: cosetFFT(a, g), 3: icosetFFT(a, g), 4: _basic_radix2_FFT(a, omega^-1), the unscaled inverse the
 * extended / step radix-2 domains call directly.  omega = get_root_of_unity(m) (field_utils.tcc:38-51,
 * Fr::s = 28).  Twiddle and coset tables are built on the device and cached per size. */
int b200_fr_fft(uint64_t *a, size_t log_n, int mode, const uint64_t *coset_g /* modes 2, 3 */);
/* same, on a device-resident vector, enqueued on cuda_stream (bench.py / callers that chain transforms) */
int b200_fr_fft_dev(void *d_a, size_t log_n, int mode, const uint64_t *coset_g, void *cuda_stream);

/* ---- wire formats: point compression of the reference's stream operators ----
 * operator<< / operator>> of alt_bn128_G1/G2 (alt_bn128_g1.cpp:404-459, alt_bn128_g2.cpp:414-475) and
 * bn128_G1/G2 (bn128_g1.cpp:344-463, bn128_g2.cpp:374-470) with point compression (the default):
 * a point is written as  is_zero | X | lsb(Y)  after to_affine_coordinates(), and read back as
 * Y = +-sqrt(X^3 + b) with the sign picked by the stored bit.  The arithmetic (one shared inversion
 * for the normalisation; one 252-bit Fq resp. 503-bit Fq2 exponentiation per point for the root) runs on
 * the device; the byte framing ('0'/'1' characters, separators) is host work (shim: b200shim::write_points /
 * read_points).  flavour 0: alt_bn128 (X in standard form, bit = lsb of Y.as_bigint()); 1: alt_bn128 with
 * -DMONTGOMERY_OUTPUT (X Montgomery image); 2: bn128 -DBINARY_OUTPUT (raw Montgomery X, bit = lsb of Y's
 * Montgomery image).  x: n x 4 (G1) / n x 8 (G2) limbs; flags[i]: bit 0 = the Y bit, bit 1 = is_zero.
 * decompress writes (X, Y, 1) Montgomery Jacobian or (0, 1, 0); bad (may be NULL) gets 1 where X^3 + b is
 * not a square -- with bad == NULL such input makes the call fail (the reference's sqrt would not return). */
int b200_compress_g1(const uint64_t *pts /* n x 12 */, size_t n, int flavour, uint64_t *x_out, uint8_t *flags);
int b200_compress_g2(const uint64_t *pts /* n x 24 */, size_t n, int flavour, uint64_t *x_out, uint8_t *flags);
int b200_decompress_g1(const uint64_t *x, const uint8_t *flags, size_t n, int flavour, uint64_t *pts_out /* n x 12 */, uint8_t *bad);
int b200_decompress_g2(const uint64_t *x, const uint8_t *flags, size_t n, int flavour, uint64_t *pts_out /* n x 24 */, uint8_t *bad);

/* ---- parity hooks: element-wise kernels over the device arithmetic ------------
 * field: 0 Fq, 1 Fr, 2 Fq2.  op: 0 mul, 1 sqr, 2 add, 3 sub, 4 inverse, 5 neg,
 * 6 as_bigint (Fq/Fr), 7 from bigint (Fq/Fr).  b may be NULL for unary ops. */
int b200_test_field_op(int field, int op, const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out);
/* group: 0 G1, 1 G2; Jacobian in / Jacobian out (not normalised).
 * op: 0 a+b, 1 a + affine b (mixed), 2 2a, 6 a - affine b, 7 k*a, 8 Jacobian->XYZZ->Jacobian */
int b200_test_group_op(int group, int op, const uint64_t *a, const uint64_t *b, size_t n, uint32_t k,
                       uint64_t *out);

/* ---- introspection for bench.py / profiling ------------------------------------ */
typedef struct {
    uint64_t n;                /* points in the last MSM (per device 0) */
    uint32_t window_bits;      /* c */
    uint32_t num_windows;      /* W */
    uint32_t chunk_len;        /* max entries per accumulation task */
    uint32_t kernel_launches;  /* kernels launched by the last call on device 0 */
    uint64_t num_tasks;        /* accumulation tasks */
    double host_finalize_us;   /* Horner + normalise on the host */
    double h2d_bytes;          /* bytes copied host->device by the last call */
    double d2h_bytes;
    uint64_t num_entries;      /* bucket entries = mixed additions done by k_accumulate (device 0) */
    double accumulate_ms;      /* CUDA-event time of k_accumulate on device 0 */
    double device_ms;          /* CUDA-event time of the whole enqueued pipeline on device 0 */
    double sort_ms;            /* CUDA-event time from the start of the pipeline to the start of k_accumulate: the digit sort */
} b200_stats_t;
int b200_last_stats(b200_stats_t *out);
/* Overrides for tuning / tests: window bits c (0 = auto), chunk length L (0 = auto). */
int b200_set_tuning(int window_bits, int chunk_len);
/* More knobs for sweeps: "reduce_log_segment" (-1 = model), "reduce_split" (0 = auto),
 * "host_horner" (0: a precomputed key's one-window reduction is weighted and summed on the device too),
 * "ones_filter" (0: scalars equal to one go through the bucket sort instead of the direct sum),
 * "use_precomputed" (0: ignore a key's precomputed levels; 1: where the cost model prefers them; 2: always).
 * Measured alternatives that stay in the tree (defaults first): "partition_sort" (1 / 0: round-1 global-atomics counting sort),
 * "reduce_quads" (1 / 0: one thread per partial sum in stage 2 of the window reduction), "reduce_marginals" (0 / 1),
 * "reduce_block" (128 / 32..96 threads per block in stage 1), "dense_direct" (1 / 0: pipelined MSMs fold every bucket after
 * every chunk), "batch_affine" (0 / 1, 2 tree levels of affine pair additions), "g2_lane_pairs" (0 / 1), "g2_blocks" (1 / 2, 3:
 * register budgets of the G2 accumulation), "g1_paired" (0 / 1), "even_chunks" (1 / 0), "overlap_sort" (1 / 0: chunked calls
 * sort and accumulate on one stream), "pinned_chunks" (0 = auto / n: upload chunks of a resident-key MSM's host scalars).
 * DESIGN.md sections 4 and 6 have the numbers.
 * The same keys can be set for a whole process with B200_TUNE="key=value,key=value"; B200_TRACE=1 prints one stderr line per
 * call (wall time, sizes, transfers) and every host stall above 1 ms inside it. */
int b200_set_tuning_ex(const char *key, int value);
/* Host-buffer MSMs (b200_msm_g1/g2) upload their inputs in `chunks` index chunks whose H2D copy
 * overlaps the sort + accumulation of the previous chunk (0 = auto: 1 below 2^17 points, 2 below
 * 2^19, 4 below 2^22, 8 below 2^24, else 16 = the maximum).  1 disables the pipeline.  The ingest + sort of a chunk run on a
 * second stream under the accumulation of the previous one (tuning key "overlap_sort"). */
int b200_set_pipeline_chunks(int chunks);
/* IMAD roofline microbenchmark on every SM of device 0; returns multiply-adds (lane-ops)
 * per second.  kind 0: the 32x32+64 multiply-add stream of the Montgomery product
 * (IMAD.WIDE.U32 / IMAD.WIDE.U32.X carry rows, 16 wide + 1 narrow per step; `iters` rounds
 * of 4 steps per thread) — the denominator of the IMAD roofline; kind 1: IMAD (32x32+32
 * low half), 8 chains; kind 2: Montgomery multiplications per second (Fq::mul chains). */
int b200_imad_peak(int kind, int iters, double *ops_per_sec, double *elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif /* B200_MSM_H */
