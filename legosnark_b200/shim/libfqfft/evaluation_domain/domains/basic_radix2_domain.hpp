// Shadow of libfqfft's basic_radix2_domain.hpp: the reference's class as it is (included next on the
// path), plus explicit specialisations of three members for BN254's Fr so that the O(n) host loops the
// reference wraps around the transform are fused into the device pass instead of running serially:
//   iFFT      : _basic_radix2_FFT(a, omega^-1) then a[i] *= 1/n          (basic_radix2_domain.tcc:49-60)
//   cosetFFT  : _multiply_by_coset(a, g) then FFT                         (:62-67; aux.tcc:163-171)
//   icosetFFT : iFFT then _multiply_by_coset(a, g^-1)                     (:69-74)
// -> b200_fr_fft modes 1 / 2 / 3 (the coset powers and 1/n are tables on the device, applied while the
// first pass loads / the last pass stores).  FFT itself already reaches the engine through the shadow of
// basic_radix2_domain_aux.hpp; the extended / step domains keep the reference's code around it.
// Only when the build's default curve is BN254 (CURVE_BN128 / CURVE_ALT_BN128, like LegoSNARK's own build).
#ifndef B200_SHIM_BASIC_RADIX2_DOMAIN_HPP_
#define B200_SHIM_BASIC_RADIX2_DOMAIN_HPP_

#include_next <libfqfft/evaluation_domain/domains/basic_radix2_domain.hpp>

#if defined(CURVE_BN128) || defined(CURVE_ALT_BN128)
#include <libff/common/default_types/ec_pp.hpp>

namespace libfqfft {
namespace b200_detail {
typedef libff::Fr<libff::default_ec_pp> FrT;

inline void fused_fft(std::vector<FrT> &a, size_t m, int mode, const FrT *g)
{
    static_assert(sizeof(FrT) == 32, "unexpected scalar layout");
    if (a.size() != m) throw DomainSizeException("basic_radix2: expected a.size() == this->m");
    ensure_engine();
    if (b200_fr_fft(reinterpret_cast<uint64_t *>(a.data()), libff::log2(m), mode, reinterpret_cast<const uint64_t *>(g)) != B200_OK)
        throw std::runtime_error(std::string("b200_fr_fft failed: ") + b200_last_error());
}
}  // namespace b200_detail

template <>
inline void basic_radix2_domain<b200_detail::FrT>::iFFT(std::vector<b200_detail::FrT> &a)
{
    b200_detail::fused_fft(a, this->m, 1, nullptr);
}
template <>
inline void basic_radix2_domain<b200_detail::FrT>::cosetFFT(std::vector<b200_detail::FrT> &a, const b200_detail::FrT &g)
{
    b200_detail::fused_fft(a, this->m, 2, &g);
}
template <>
inline void basic_radix2_domain<b200_detail::FrT>::icosetFFT(std::vector<b200_detail::FrT> &a, const b200_detail::FrT &g)
{
    b200_detail::fused_fft(a, this->m, 3, &g);
}

}  // namespace libfqfft
#endif  // BN254 default curve

#endif  // B200_SHIM_BASIC_RADIX2_DOMAIN_HPP_
