/* bn254_oracle.h — C-ABI of the plain-C CPU restatement of the reference's MSM /
 * batch_exp hot path.  TEST INFRASTRUCTURE ONLY: may be imported by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs,
 * never by the product library.
 *
 * Parity status: PINNED — tests/test_oracle_vs_reference.py checks every entry
 * point bit-for-bit against oracle/_ref/libffref.so (the unmodified reference
 * sources compiled in place) and tests/test_oracle_golden.py against the
 * fixtures in tests/golden/ that tools/make_golden.py produced from that same
 * reference build.  The Fr vector routines (fold / evalMLE / mle_bind / fft) are pinned
 * the same way by tests/test_fr_vectors.py against oracle/_ref/liblsref.so (the reference's
 * own classes behind oracle/ref_wrap_ls.cpp) and tests/golden/fr_vectors.npz
 * (tools/make_golden_fr.py).
 *
 * All buffers: little-endian u64 limbs, Montgomery form with R = 2^256
 * (LFF/algebra/fields/fp.hpp:42).  G1 point = X|Y|Z = 12 limbs (Jacobian,
 * Z == 0 <=> infinity).  G2 point = X.c0|X.c1|Y.c0|Y.c1|Z.c0|Z.c1 = 24 limbs.
 * The same ABI is exported by oracle/ref_wrap.cpp with the prefix ref_.
 */
#ifndef BN254_ORACLE_H
#define BN254_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

int orc_max_threads(void);

/* variant: 0 = multi_exp<BDLO12>, 1 = multi_exp_with_mixed_addition<BDLO12>,
 *          4 = multi_exp<naive_plain>.  normalise != 0 => to_affine_coordinates() */
int orc_msm_g1(const uint64_t *bases, const uint64_t *scalars, size_t n, size_t chunks, int variant,
               int normalise, uint64_t *out);
int orc_msm_g2(const uint64_t *bases, const uint64_t *scalars, size_t n, size_t chunks, int variant,
               int normalise, uint64_t *out);

/* get_exp_window_size + get_window_table + batch_exp[_with_coeff] (+ batch_to_special) */
int orc_batch_exp_g1(const uint64_t *base, const uint64_t *scalars, size_t n, const uint64_t *coeff,
                     size_t window, int normalise, uint64_t *out);
int orc_batch_exp_g2(const uint64_t *base, const uint64_t *scalars, size_t n, const uint64_t *coeff,
                     size_t window, int normalise, uint64_t *out);
size_t orc_exp_window_size_g1(size_t n);
size_t orc_exp_window_size_g2(size_t n);

int orc_batch_to_special_g1(uint64_t *pts, size_t n);
int orc_batch_to_special_g2(uint64_t *pts, size_t n);

/* op: 0 operator+, 1 mixed_add, 2 dbl, 3 to_affine_coordinates, 4 negate, 5 add() */
int orc_g1_op(int op, const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out);
int orc_g2_op(int op, const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out);

int orc_scalar_mul_g1(const uint64_t *base, const uint64_t *scalars, size_t n, int stride_base,
                      int normalise, uint64_t *out);
int orc_scalar_mul_g2(const uint64_t *base, const uint64_t *scalars, size_t n, int stride_base,
                      int normalise, uint64_t *out);

/* op: 0 mul, 1 squared, 2 add, 3 sub, 4 inverse, 5 neg */
int orc_fq_op(int op, const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out);
int orc_fr_op(int op, const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out);
int orc_fq2_op(int op, const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out);

int orc_fr_from_bigint(const uint64_t *a, size_t n, uint64_t *out);
int orc_fr_as_bigint(const uint64_t *a, size_t n, uint64_t *out);
int orc_fq_from_bigint(const uint64_t *a, size_t n, uint64_t *out);

int orc_sha512_rng_fr(uint64_t idx0, size_t n, uint64_t *out);

/* Fr vector work next to the MSMs; the same functions are exported by
 * oracle/ref_wrap_ls.cpp (prefix ref_) from the reference's own headers.
 * fold: CPPoly::prove's w_coeffs + last tmp_v[0] (LS/gadgets/poly.h:45-67); eval_mle:
 * MultiVPolyT::evalMLE (LS/prototools/polytools.h:207-234); mle_bind: DPMle::pushRandomness
 * (LS/prototools/mle.h:199-210); fft: basic_radix2_domain<Fr> mode 0 FFT, 1 iFFT, 2 cosetFFT,
 * 3 icosetFFT (FQFFT/evaluation_domain/domains/basic_radix2_domain.tcc). */
int orc_fr_fold_witness(const uint64_t *v, const uint64_t *r, size_t d, uint64_t *w_coeffs, uint64_t *eval);
int orc_fr_eval_mle(const uint64_t *v, const uint64_t *r, size_t d, uint64_t *out);
int orc_fr_mle_bind(const uint64_t *table, size_t half, const uint64_t *r, uint64_t *out);
int orc_fr_step_fft(uint64_t *a, size_t log_big, size_t log_small, int mode, const uint64_t *g);
int orc_fr_eq_table(const uint64_t *r, size_t d, uint64_t *out);
int orc_fr_matrix_mle(const uint64_t *A, const uint64_t *rho, size_t d, uint64_t *v);
int orc_fr_sumcheck_round(const uint64_t *a, const uint64_t *b, const uint64_t *w, size_t half, uint64_t *out);
int orc_fr_sumcheck_rounds(const uint64_t *a, const uint64_t *b, const uint64_t *r, size_t d, uint64_t *h);
int orc_fr_fft(uint64_t *a, size_t log_n, int mode, const uint64_t *g);

/* wire format: the arithmetic of the reference's compressed operator<< / operator>>
 * (alt_bn128_g1.cpp:404-459, alt_bn128_g2.cpp:414-475, bn128_g1.cpp:344-463, bn128_g2.cpp:374-470).
 * flavour 0 alt_bn128, 1 alt_bn128 -DMONTGOMERY_OUTPUT, 2 bn128; flags[i]: bit 0 = Y bit, bit 1 = is_zero.
 * ref_* (oracle/ref_wrap.cpp) runs the reference's own stream operators: curve 0 = flavour 0, curve 1 = flavour 2. */
int orc_compress_g1(const uint64_t *pts, size_t n, int flavour, uint64_t *x_out, uint8_t *flags);
int orc_compress_g2(const uint64_t *pts, size_t n, int flavour, uint64_t *x_out, uint8_t *flags);
int orc_decompress_g1(const uint64_t *x, const uint8_t *flags, size_t n, int flavour, uint64_t *pts_out);
int orc_decompress_g2(const uint64_t *x, const uint8_t *flags, size_t n, int flavour, uint64_t *pts_out);

int orc_g1_one(uint64_t *out);
int orc_g2_one(uint64_t *out);

#ifdef __cplusplus
}
#endif
#endif
