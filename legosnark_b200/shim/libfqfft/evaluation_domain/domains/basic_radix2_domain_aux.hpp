// Shadow of libfqfft's basic_radix2_domain_aux.hpp (FQFFT = depends/libsnark/depends/libfqfft/libfqfft):
// with `legosnark_b200/shim` ahead of libfqfft on the include path, every radix-2 transform of
// the reference runs on the B200 engine when the field is BN254's Fr — no reference file is edited.
//
// The reference funnels ALL its radix-2 work through one function,
//   _basic_radix2_FFT(std::vector<FieldT> &a, const FieldT &omega)
//     (FQFFT/evaluation_domain/domains/basic_radix2_domain_aux.hpp:23-24; the name is a macro for
//      the serial / OpenMP variant, basic_radix2_domain_aux.tcc:33-37, 42-75, 150-168),
// called by basic_radix2_domain::FFT/iFFT (basic_radix2_domain.tcc:41-60), extended_radix2_domain
// (extended_radix2_domain.tcc) and step_radix2_domain (step_radix2_domain.tcc), i.e. by everything
// get_evaluation_domain() hands to libsnark's r1cs_to_qap_witness_map and LegoSNARK's Interpolator
// (LS/prototools/interp.h:61-65) / lipmaa.cc:94-185.  This header includes the reference's own
// header (next on the path), then re-points that macro at b200_radix2_FFT below.  The unscaled
// transform with omega = get_root_of_unity(n) or its inverse goes to b200_fr_fft (modes 0 / 4);
// other fields (libff::Double, other curves) and other roots keep the reference's template.
// Scaling by 1/n, coset shifts and divide_by_Z_on_coset stay in the caller as the reference wrote them.
// _basic_radix2_evaluate_all_lagrange_polynomials (one inversion per point in the reference) is re-pointed the same way.
#ifndef B200_SHIM_BASIC_RADIX2_DOMAIN_AUX_HPP_
#define B200_SHIM_BASIC_RADIX2_DOMAIN_AUX_HPP_

#include_next <libfqfft/evaluation_domain/domains/basic_radix2_domain_aux.hpp>

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include <libff/algebra/fields/field_utils.hpp>

#include <cstdlib>

#include "b200_msm.h"

namespace libfqfft {
namespace b200_detail {

// Fp_model<4, r> with r = BN254's group order (alt_bn128_init.cpp:40, bn128_init.cpp:38): detected by
// shape (4 limbs, mont_repr, static mod) at compile time and by the modulus limbs at run time.
template <typename FieldT, typename = void>
struct looks_like_fp4 : std::false_type {};
template <typename FieldT>
struct looks_like_fp4<FieldT, typename std::enable_if<FieldT::num_limbs == 4 && sizeof(FieldT) == 32 &&
                                                        sizeof(decltype(FieldT::mod.data)) == 32>::type> : std::true_type {};

template <typename FieldT>
inline bool is_bn254_fr()
{
    static const uint64_t R[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
    return std::memcmp(FieldT::mod.data, R, 32) == 0;
}

inline void ensure_engine()
{
    static const bool once = [] {
        const char *e = std::getenv("B200_GPUS");  // same rule as b200shim::ensure_init: whichever shim runs first, B200_GPUS holds
        if (b200_device_count() == 0 && b200_init(e ? std::atoi(e) : 0) != B200_OK)
            throw std::runtime_error(std::string("b200_init failed: ") + b200_last_error());
        return true;
    }();
    (void)once;
}

template <typename FieldT, bool Fp4 = looks_like_fp4<FieldT>::value>
struct fft_dispatch {
    static void run(std::vector<FieldT> &a, const FieldT &omega) { _basic_radix2_FFT(a, omega); }  // the reference's macro
};

template <typename FieldT>
struct fft_dispatch<FieldT, true> {
    static void run(std::vector<FieldT> &a, const FieldT &omega)
    {
        const size_t n = a.size(), logn = libff::log2(n);
        if (n != ((size_t)1 << logn)) throw DomainSizeException("expected n == (1u << logn)");  // aux.tcc:45
        if (n < 2 || logn > FieldT::s || !is_bn254_fr<FieldT>()) {
            _basic_radix2_FFT(a, omega);
            return;
        }
        const FieldT w = libff::get_root_of_unity<FieldT>(n);
        int mode;
        if (omega == w) mode = 0;
        else if (omega * w == FieldT::one()) mode = 4;
        else {
            _basic_radix2_FFT(a, omega);  // some other primitive root: not a case the engine covers
            return;
        }
        ensure_engine();
        if (b200_fr_fft(reinterpret_cast<uint64_t *>(a.data()), logn, mode, nullptr) != B200_OK)
            throw std::runtime_error(std::string("b200_fr_fft failed: ") + b200_last_error());  // no CPU fallback behind a failure
    }
};

// _basic_radix2_evaluate_all_lagrange_polynomials (FQFFT/evaluation_domain/domains/basic_radix2_domain_aux.tcc:183-236): the
// reference computes u[i] = l * (t - r).inverse() with ONE Fp inversion per domain point (:228-233); libsnark's generators reach it
// through r1cs_to_qap_instance_map_with_evaluation (r1cs_to_qap.tcc:127-190).  Here the host forms l_0 = (t^m - 1) / m and the
// device computes u[i] = l_0 omega^i / (t - omega^i) with batched inversions (b200_fr_geometric_quotients).  t on the domain itself
// (t^m == 1) keeps the reference's search loop (:201-214), which divides nothing.
template <typename FieldT, bool Fp4 = looks_like_fp4<FieldT>::value>
struct lagrange_dispatch {
    static std::vector<FieldT> run(const size_t m, const FieldT &t) { return _basic_radix2_evaluate_all_lagrange_polynomials(m, t); }
};

template <typename FieldT>
struct lagrange_dispatch<FieldT, true> {
    static std::vector<FieldT> run(const size_t m, const FieldT &t)
    {
        if (m < 2 || !is_bn254_fr<FieldT>()) return _basic_radix2_evaluate_all_lagrange_polynomials(m, t);
        if (m != ((size_t)1 << libff::log2(m))) throw DomainSizeException("expected m == (1u << log2(m))");  // aux.tcc:190
        const FieldT tm = t ^ m;
        if (tm == FieldT::one()) return _basic_radix2_evaluate_all_lagrange_polynomials(m, t);
        const FieldT omega = libff::get_root_of_unity<FieldT>(m);
        // a0, a_ratio, then the factor (c1, ratio, c0): u[i] = a0 omega^i / (-omega^i + t)
        const FieldT consts[5] = {(tm - FieldT::one()) * FieldT(m).inverse(), omega, -FieldT::one(), omega, -t};
        std::vector<FieldT> u(m, FieldT::zero());
        ensure_engine();
        if (b200_fr_geometric_quotients(reinterpret_cast<uint64_t *>(u.data()), nullptr, m, reinterpret_cast<const uint64_t *>(consts), 1) != B200_OK)
            throw std::runtime_error(std::string("b200_fr_geometric_quotients failed: ") + b200_last_error());
        return u;
    }
};

}  // namespace b200_detail

template <typename FieldT>
void b200_radix2_FFT(std::vector<FieldT> &a, const FieldT &omega)
{
    b200_detail::fft_dispatch<FieldT>::run(a, omega);
}

template <typename FieldT>
std::vector<FieldT> b200_radix2_lagrange(const size_t m, const FieldT &t)
{
    return b200_detail::lagrange_dispatch<FieldT>::run(m, t);
}

}  // namespace libfqfft

// from here on the reference's call sites (basic / extended / step radix-2 domains) reach the engine
#undef _basic_radix2_FFT
#define _basic_radix2_FFT b200_radix2_FFT
// (a function template in the reference, not a macro: the definition above this line keeps its name, the callers below get ours)
#define _basic_radix2_evaluate_all_lagrange_polynomials b200_radix2_lagrange

#endif  // B200_SHIM_BASIC_RADIX2_DOMAIN_AUX_HPP_
