// Shadow of libfqfft's step_radix2_domain.hpp: the reference's class as it is (included next on the path), plus an
// explicit specialisation of divide_by_Z_on_coset for BN254's Fr.  The reference divides by the vanishing polynomial
// with one Fp inversion PER POINT of the big half (step_radix2_domain.tcc:213-241: `P[i] *= (... * elt - ...).inverse()`
// for i < big_m, serial, ~3 us each through mpn_gcdext): 6 s of the 6.9 s Groth16 prover of the 128 x 128 matrix product
// (2^21 + 1 constraints -> this domain), BASELINE.json configs[3].  Here the host forms the four constants exactly as
// the reference does and the device inverts the denominators in batches (b200_fr_scale_inv_geometric).
// FFT / iFFT / cosetFFT / icosetFFT are specialised as well: the reference runs the two radix-2 transforms inside them
// through _basic_radix2_FFT (which the engine already serves) but wraps them in serial O(m) host loops
// (omega_i *= omega chains, :43-50, :95-101, :118-124; _multiply_by_coset, aux.tcc:172-180) that cost more than the
// transforms; b200_fr_step_fft runs the whole member on the device with one upload and one download.
#ifndef B200_SHIM_STEP_RADIX2_DOMAIN_HPP_
#define B200_SHIM_STEP_RADIX2_DOMAIN_HPP_

#include_next <libfqfft/evaluation_domain/domains/step_radix2_domain.hpp>

#if defined(CURVE_BN128) || defined(CURVE_ALT_BN128)
#include <libff/common/default_types/ec_pp.hpp>
#include <libfqfft/evaluation_domain/domains/basic_radix2_domain_aux.hpp>

namespace libfqfft {

namespace b200_detail {
inline void step_fft(std::vector<libff::Fr<libff::default_ec_pp>> &a, size_t m, size_t big_m, size_t small_m, int mode,
                     const libff::Fr<libff::default_ec_pp> *g)
{
    static_assert(sizeof(libff::Fr<libff::default_ec_pp>) == 32, "unexpected scalar layout");
    if (a.size() != m) throw DomainSizeException("step_radix2: expected a.size() == this->m");
    ensure_engine();
    if (b200_fr_step_fft(reinterpret_cast<uint64_t *>(a.data()), libff::log2(big_m), libff::log2(small_m), mode,
                         reinterpret_cast<const uint64_t *>(g)) != B200_OK)
        throw std::runtime_error(std::string("b200_fr_step_fft failed: ") + b200_last_error());
}
}  // namespace b200_detail

template <>
inline void step_radix2_domain<libff::Fr<libff::default_ec_pp>>::FFT(std::vector<libff::Fr<libff::default_ec_pp>> &a)
{
    b200_detail::step_fft(a, this->m, big_m, small_m, 0, nullptr);
}
template <>
inline void step_radix2_domain<libff::Fr<libff::default_ec_pp>>::iFFT(std::vector<libff::Fr<libff::default_ec_pp>> &a)
{
    b200_detail::step_fft(a, this->m, big_m, small_m, 1, nullptr);
}
template <>
inline void step_radix2_domain<libff::Fr<libff::default_ec_pp>>::cosetFFT(std::vector<libff::Fr<libff::default_ec_pp>> &a,
                                                                          const libff::Fr<libff::default_ec_pp> &g)
{
    b200_detail::step_fft(a, this->m, big_m, small_m, 2, &g);
}
template <>
inline void step_radix2_domain<libff::Fr<libff::default_ec_pp>>::icosetFFT(std::vector<libff::Fr<libff::default_ec_pp>> &a,
                                                                           const libff::Fr<libff::default_ec_pp> &g)
{
    b200_detail::step_fft(a, this->m, big_m, small_m, 3, &g);
}

template <>
inline void step_radix2_domain<libff::Fr<libff::default_ec_pp>>::divide_by_Z_on_coset(std::vector<libff::Fr<libff::default_ec_pp>> &P)
{
    typedef libff::Fr<libff::default_ec_pp> FieldT;
    static_assert(sizeof(FieldT) == 32, "unexpected scalar layout");
    if (P.size() < big_m + small_m) throw DomainSizeException("step_radix2: divide_by_Z_on_coset expects the whole coset");
    // the constants of step_radix2_domain.tcc:216-222 and :236-237
    const FieldT coset = FieldT::multiplicative_generator;
    const FieldT Z0 = (coset ^ big_m) - FieldT::one();
    const FieldT c1 = (coset ^ small_m) * Z0;
    const FieldT c0 = (omega ^ small_m) * Z0;
    const FieldT ratio = omega ^ (2 * small_m);
    const FieldT Z1 = ((((coset * omega) ^ big_m) - FieldT::one()) * (((coset * omega) ^ small_m) - (omega ^ small_m)));
    const FieldT Z1_inverse = Z1.inverse();
    b200_detail::ensure_engine();
    if (b200_fr_scale_inv_geometric(reinterpret_cast<uint64_t *>(P.data()), big_m, reinterpret_cast<const uint64_t *>(&c1),
                                    reinterpret_cast<const uint64_t *>(&ratio), reinterpret_cast<const uint64_t *>(&c0), small_m,
                                    reinterpret_cast<const uint64_t *>(&Z1_inverse)) != B200_OK)
        throw std::runtime_error(std::string("b200_fr_scale_inv_geometric failed: ") + b200_last_error());
}

// evaluate_all_lagrange_polynomials (step_radix2_domain.tcc:161-186): besides the two radix-2 Lagrange vectors (one inversion per
// point each in the reference, served by the shadow of basic_radix2_domain_aux.hpp) the big half is multiplied by
// L0 * (elt - omega^small_m).inverse() with one more inversion per point (:172-176).  Both denominators go to the device in one
// call: result[i] = (l_0 L0) big_omega^i / ((t - big_omega^i) (rho^i - omega^small_m)), rho = big_omega^small_m.
template <>
inline std::vector<libff::Fr<libff::default_ec_pp>> step_radix2_domain<libff::Fr<libff::default_ec_pp>>::evaluate_all_lagrange_polynomials(
    const libff::Fr<libff::default_ec_pp> &t)
{
    typedef libff::Fr<libff::default_ec_pp> FieldT;
    std::vector<FieldT> result(this->m, FieldT::zero());
    const FieldT L0 = (t ^ small_m) - (omega ^ small_m);
    const FieldT omega_to_small_m = omega ^ small_m;
    const FieldT big_omega_to_small_m = big_omega ^ small_m;
    const FieldT t_big = t ^ big_m;
    b200_detail::ensure_engine();
    if (t_big == FieldT::one()) {
        // t on the big domain: inner_big is the reference's unit vector (aux.tcc:201-214); the :172-176 loop as one device pass over it
        const std::vector<FieldT> inner_big = _basic_radix2_evaluate_all_lagrange_polynomials(big_m, t);
        const FieldT consts[5] = {L0, FieldT::one(), FieldT::one(), big_omega_to_small_m, omega_to_small_m};
        if (b200_fr_geometric_quotients(reinterpret_cast<uint64_t *>(result.data()), reinterpret_cast<const uint64_t *>(inner_big.data()), big_m,
                                        reinterpret_cast<const uint64_t *>(consts), 1) != B200_OK)
            throw std::runtime_error(std::string("b200_fr_geometric_quotients failed: ") + b200_last_error());
    } else {
        const FieldT consts[8] = {(t_big - FieldT::one()) * FieldT(big_m).inverse() * L0, big_omega,        // numerator l_0 L0 big_omega^i
                                  -FieldT::one(), big_omega, -t,                                                // t - big_omega^i
                                  FieldT::one(), big_omega_to_small_m, omega_to_small_m};                       // rho^i - omega^small_m
        if (b200_fr_geometric_quotients(reinterpret_cast<uint64_t *>(result.data()), nullptr, big_m, reinterpret_cast<const uint64_t *>(consts), 2) !=
            B200_OK)
            throw std::runtime_error(std::string("b200_fr_geometric_quotients failed: ") + b200_last_error());
    }
    const std::vector<FieldT> inner_small = _basic_radix2_evaluate_all_lagrange_polynomials(small_m, t * omega.inverse());  // :164
    const FieldT L1 = ((t ^ big_m) - FieldT::one()) * ((omega ^ big_m) - FieldT::one()).inverse();                           // :178
    for (size_t i = 0; i < small_m; ++i) result[big_m + i] = L1 * inner_small[i];
    return result;
}

}  // namespace libfqfft
#endif  // BN254 default curve

#endif  // B200_SHIM_STEP_RADIX2_DOMAIN_HPP_
