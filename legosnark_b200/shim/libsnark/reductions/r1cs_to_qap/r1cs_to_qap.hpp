// Shadow of libsnark's r1cs_to_qap.hpp (SNK = depends/libsnark/libsnark): the reference header is included as it is, with
// its witness map reachable as libsnark::libsnark_cpu_r1cs_to_qap_witness_map, and r1cs_to_qap_witness_map is re-declared
// with the reference's signature (SNK/reductions/r1cs_to_qap/r1cs_to_qap.hpp:57-63).
//
// For BN254's Fr, d1 = d2 = d3 = 0 (what r1cs_gg_ppzksnark_prover passes, r1cs_gg_ppzksnark.tcc:402-415) and a basic or
// step radix-2 domain, the map is split where the data changes hands:
//   host    the constraint polynomials' evaluations aA, aB, aC on the domain: libsnark's own linear_combination::evaluate
//           per constraint (r1cs_to_qap.tcc:232-246, 281-286), here under `omp parallel for` (the reference loops serially);
//           the product check aA[i] aB[i] == aC[i] of the same loop stands in for `assert(cs.is_satisfied(...))` (:220);
//   device  everything between those vectors and coefficients_for_H: b200_qap_h_coefficients (three iFFTs, three coset
//           FFTs, A B - C, divide_by_Z_on_coset, the inverse coset FFT; one upload, one download).
// Any other field, domain or non-zero d1 / d2 / d3 (the ZK patch needs the coefficient vectors on the host) runs the
// reference's function, whose transforms still reach the engine through the libfqfft shadows.
// 128 x 128 matrix product (2^21 + 1 constraints): witness map 6.55 s (round 1) -> 1.15 s (per-transform shims) -> see
// profiles/; the H vector equals the reference's element for element (integration/groth16matrix_driver.cc checks it).
#ifndef B200_SHIM_R1CS_TO_QAP_HPP_
#define B200_SHIM_R1CS_TO_QAP_HPP_

#define r1cs_to_qap_witness_map libsnark_cpu_r1cs_to_qap_witness_map
#include_next <libsnark/reductions/r1cs_to_qap/r1cs_to_qap.hpp>
#undef r1cs_to_qap_witness_map

#include <cassert>
#include <stdexcept>
#include <string>

#include <libfqfft/evaluation_domain/domains/basic_radix2_domain.hpp>
#include <libfqfft/evaluation_domain/domains/step_radix2_domain.hpp>
#include <libfqfft/evaluation_domain/get_evaluation_domain.hpp>

#include "b200_msm.h"

namespace libsnark {
namespace b200_detail {

template <typename FieldT, bool Fp4 = libfqfft::b200_detail::looks_like_fp4<FieldT>::value>
struct witness_map_dispatch {
    static qap_witness<FieldT> run(const r1cs_constraint_system<FieldT> &cs, const r1cs_primary_input<FieldT> &primary_input,
                                   const r1cs_auxiliary_input<FieldT> &auxiliary_input, const FieldT &d1, const FieldT &d2, const FieldT &d3)
    {
        return libsnark_cpu_r1cs_to_qap_witness_map(cs, primary_input, auxiliary_input, d1, d2, d3);
    }
};

template <typename FieldT>
struct witness_map_dispatch<FieldT, true> {
    static qap_witness<FieldT> run(const r1cs_constraint_system<FieldT> &cs, const r1cs_primary_input<FieldT> &primary_input,
                                   const r1cs_auxiliary_input<FieldT> &auxiliary_input, const FieldT &d1, const FieldT &d2, const FieldT &d3)
    {
        const std::shared_ptr<libfqfft::evaluation_domain<FieldT>> domain =
            libfqfft::get_evaluation_domain<FieldT>(cs.num_constraints() + cs.num_inputs() + 1);
        const auto *step = dynamic_cast<const libfqfft::step_radix2_domain<FieldT> *>(domain.get());
        const auto *basic = dynamic_cast<const libfqfft::basic_radix2_domain<FieldT> *>(domain.get());
        if (!libfqfft::b200_detail::is_bn254_fr<FieldT>() || !(d1.is_zero() && d2.is_zero() && d3.is_zero()) || (!step && !basic))
            return libsnark_cpu_r1cs_to_qap_witness_map(cs, primary_input, auxiliary_input, d1, d2, d3);

        const size_t m = domain->m, nc = cs.num_constraints();
        r1cs_variable_assignment<FieldT> full = primary_input;
        full.insert(full.end(), auxiliary_input.begin(), auxiliary_input.end());
        std::vector<FieldT> aA(m, FieldT::zero()), aB(m, FieldT::zero()), aC(m, FieldT::zero());
        for (size_t i = 0; i <= cs.num_inputs(); ++i) aA[i + nc] = (i > 0 ? full[i - 1] : FieldT::one());  // input_i * 0 = 0
        int satisfied = 1;
#ifdef MULTICORE
#pragma omp parallel for reduction(&& : satisfied)
#endif
        for (size_t i = 0; i < nc; ++i) {
            aA[i] = cs.constraints[i].a.evaluate(full);
            aB[i] = cs.constraints[i].b.evaluate(full);
            aC[i] = cs.constraints[i].c.evaluate(full);
            satisfied = satisfied && (aA[i] * aB[i] == aC[i]);
        }
        assert(satisfied);  // the reference: assert(cs.is_satisfied(primary_input, auxiliary_input))
        (void)satisfied;

        // the constants of divide_by_Z_on_coset, formed with the reference's own field arithmetic
        const FieldT coset = FieldT::multiplicative_generator;
        FieldT div[4];
        size_t log_big, log_small;
        if (basic) {
            log_big = libff::log2(m);
            log_small = B200_QAP_BASIC;
            div[0] = ((coset ^ m) - FieldT::one()).inverse();  // basic_radix2_domain::divide_by_Z_on_coset
        } else {
            const size_t big_m = step->big_m, small_m = step->small_m;
            const FieldT omega = step->omega;
            log_big = libff::log2(big_m);
            log_small = libff::log2(small_m);
            const FieldT Z0 = (coset ^ big_m) - FieldT::one();  // step_radix2_domain.tcc:216-222, 236-237
            div[0] = (coset ^ small_m) * Z0;
            div[1] = omega ^ (2 * small_m);
            div[2] = (omega ^ small_m) * Z0;
            div[3] = ((((coset * omega) ^ big_m) - FieldT::one()) * (((coset * omega) ^ small_m) - (omega ^ small_m))).inverse();
        }
        libfqfft::b200_detail::ensure_engine();
        std::vector<FieldT> H(m + 1, FieldT::zero());
        static_assert(sizeof(FieldT) == 32, "unexpected scalar layout");
        if (b200_qap_h_coefficients(reinterpret_cast<const uint64_t *>(aA.data()), reinterpret_cast<const uint64_t *>(aB.data()),
                                    reinterpret_cast<const uint64_t *>(aC.data()), log_big, log_small,
                                    reinterpret_cast<const uint64_t *>(&coset), reinterpret_cast<const uint64_t *>(div),
                                    reinterpret_cast<uint64_t *>(H.data())) != B200_OK)
            throw std::runtime_error(std::string("b200_qap_h_coefficients failed: ") + b200_last_error());
        return qap_witness<FieldT>(cs.num_variables(), m, cs.num_inputs(), d1, d2, d3, full, std::move(H));
    }
};

}  // namespace b200_detail

template <typename FieldT>
qap_witness<FieldT> r1cs_to_qap_witness_map(const r1cs_constraint_system<FieldT> &cs, const r1cs_primary_input<FieldT> &primary_input,
                                            const r1cs_auxiliary_input<FieldT> &auxiliary_input, const FieldT &d1, const FieldT &d2,
                                            const FieldT &d3)
{
    return b200_detail::witness_map_dispatch<FieldT>::run(cs, primary_input, auxiliary_input, d1, d2, d3);
}

}  // namespace libsnark
#endif  // B200_SHIM_R1CS_TO_QAP_HPP_
