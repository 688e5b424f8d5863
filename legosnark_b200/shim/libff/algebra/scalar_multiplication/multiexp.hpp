// Shadow of libff's multiexp.hpp: put `legosnark_b200/shim` (and `include/`) AHEAD of
// libff on the include path and every caller of
//   libff::multi_exp / multi_exp_with_mixed_addition            (LFF multiexp.hpp:56-75)
//   libff::get_window_table / windowed_exp / batch_exp /
//          batch_exp_with_coeff / batch_to_special              (LFF multiexp.hpp:96-127)
// compiles unchanged and runs on the B200 engine for G in {alt_bn128, bn128} x {G1, G2};
// no reference file is edited.  LegoSNARK reaches these through
// LS/utils/globl.h:47-78, LS/utils/util.h:119-134, LS/prototools/interp.h:36-65,
// LS/utils/sparsemexp.h:58,89 and libsnark's r1cs_gg_ppzksnark.tcc:300-360,442-484.
//
// Mechanism: the reference header (next on the include path) is included with its entry
// points renamed to libff::libff_cpu_*, so its templates stay available for the types the
// engine does not cover (knowledge_commitment<T1,T2>, other curves); the public names are
// then re-declared here with the reference's exact signatures.
#ifndef B200_SHIM_MULTIEXP_HPP_
#define B200_SHIM_MULTIEXP_HPP_

#define multi_exp libff_cpu_multi_exp
#define multi_exp_with_mixed_addition libff_cpu_multi_exp_with_mixed_addition
#define get_window_table libff_cpu_get_window_table
#define windowed_exp libff_cpu_windowed_exp
#define batch_exp libff_cpu_batch_exp
#define batch_exp_with_coeff libff_cpu_batch_exp_with_coeff
#define batch_to_special libff_cpu_batch_to_special
#include_next <libff/algebra/scalar_multiplication/multiexp.hpp>
#undef multi_exp
#undef multi_exp_with_mixed_addition
#undef get_window_table
#undef windowed_exp
#undef batch_exp
#undef batch_exp_with_coeff
#undef batch_to_special

#include "b200_libff.hpp"

namespace libff {
namespace b200_detail {

// the engine serves the four BN254 groups with their own scalar field; any other instantiation (another group, or a
// supported group with a foreign FieldT) keeps the reference's template
template <typename T, typename FieldT, bool Supported = b200shim::group_traits<T>::supported>
struct on_gpu : std::false_type {};
template <typename T, typename FieldT>
struct on_gpu<T, FieldT, true> : std::is_same<FieldT, typename T::scalar_field> {};

template <typename T, typename FieldT, multi_exp_method Method, bool OnGpu = on_gpu<T, FieldT>::value>
struct msm_dispatch {
    typedef typename std::vector<T>::const_iterator TI;
    typedef typename std::vector<FieldT>::const_iterator SI;
    static T plain(TI vs, TI ve, SI ss, SI se, const size_t chunks)
    {
        return libff_cpu_multi_exp<T, FieldT, Method>(vs, ve, ss, se, chunks);
    }
    static T mixed(TI vs, TI ve, SI ss, SI se, const size_t chunks)
    {
        return libff_cpu_multi_exp_with_mixed_addition<T, FieldT, Method>(vs, ve, ss, se, chunks);
    }
};

// Every method / chunk count denotes the same group element (multiexp.tcc:402-496); the zero
// and one pre-filter of multi_exp_with_mixed_addition (:455-487) is subsumed by the digit
// recoding on the device (zero digits are skipped).
template <typename T, typename FieldT, multi_exp_method Method>
struct msm_dispatch<T, FieldT, Method, true> {
    typedef typename std::vector<T>::const_iterator TI;
    typedef typename std::vector<FieldT>::const_iterator SI;
    static T plain(TI vs, TI ve, SI ss, SI se, const size_t)
    {
        const size_t n = (size_t)(ve - vs);
        if ((size_t)(se - ss) != n) throw std::runtime_error("multi_exp: bases and scalars differ in length");
        return b200shim::msm<T, FieldT>(n ? &*vs : nullptr, n ? &*ss : nullptr, n);
    }
    static T mixed(TI vs, TI ve, SI ss, SI se, const size_t chunks) { return plain(vs, ve, ss, se, chunks); }
};

// A lazy window table is one row {0, g}: the engine builds its own affine table on the
// device, and no caller indexes the reference's.
template <typename T>
inline bool is_lazy_table(const size_t, const size_t, const window_table<T> &t)
{
    return t.size() == 1 && t[0].size() == 2;  // what table_dispatch<T, true>::make returns, whatever (scalar_size, window) say
}
// row 0 of any window table is {0, g, 2g, ...} (multiexp.tcc:563-578): t[0][1] is the base
template <typename T>
inline const T &table_base(const window_table<T> &t)
{
    if (t.empty() || t[0].size() < 2) throw std::runtime_error("window_table has no base entry (expected row 0 = {0, g, ...})");
    return t[0][1];
}

template <typename T, bool OnGpu = b200shim::group_traits<T>::supported>
struct table_dispatch {
    static window_table<T> make(const size_t scalar_size, const size_t window, const T &g)
    {
        return libff_cpu_get_window_table<T>(scalar_size, window, g);
    }
    template <typename FieldT>
    static T one_exp(const size_t scalar_size, const size_t window, const window_table<T> &t, const FieldT &pow)
    {
        return libff_cpu_windowed_exp<T, FieldT>(scalar_size, window, t, pow);
    }
    template <typename FieldT>
    static std::vector<T> many(const size_t scalar_size, const size_t window, const window_table<T> &t, const FieldT *coeff,
                               const std::vector<FieldT> &v)
    {
        return coeff ? libff_cpu_batch_exp_with_coeff<T, FieldT>(scalar_size, window, t, *coeff, v)
                     : libff_cpu_batch_exp<T, FieldT>(scalar_size, window, t, v);
    }
    static void special(std::vector<T> &vec) { libff_cpu_batch_to_special<T>(vec); }
};

template <typename T>
struct table_dispatch<T, true> {
    static window_table<T> make(const size_t, const size_t, const T &g)
    {
        return window_table<T>(1, std::vector<T>{T::zero(), g});
    }
    // single exponentiations (kc_batch_exp_internal, SNK/knowledge_commitment/kc_multiexp.tcc:113-114)
    // are not a throughput path: Fr * G on the host (curve_utils.tcc:13-34), as the reference does for r*h
    template <typename FieldT>
    static T one_exp(const size_t scalar_size, const size_t window, const window_table<T> &t, const FieldT &pow)
    {
        if (is_lazy_table(scalar_size, window, t)) return pow * table_base(t);
        return libff_cpu_windowed_exp<T, FieldT>(scalar_size, window, t, pow);
    }
    template <typename FieldT>
    static std::vector<T> many(const size_t, const size_t, const window_table<T> &t, const FieldT *coeff,
                               const std::vector<FieldT> &v)
    {
        return b200shim::fixed_base_exp<T, FieldT>(table_base(t), v, coeff);
    }
    static void special(std::vector<T> &vec) { b200shim::to_special<T>(vec); }
};

}  // namespace b200_detail

template <typename T, typename FieldT, multi_exp_method Method>
T multi_exp(typename std::vector<T>::const_iterator vec_start, typename std::vector<T>::const_iterator vec_end,
            typename std::vector<FieldT>::const_iterator scalar_start,
            typename std::vector<FieldT>::const_iterator scalar_end, const size_t chunks)
{
    return b200_detail::msm_dispatch<T, FieldT, Method>::plain(vec_start, vec_end, scalar_start, scalar_end, chunks);
}

template <typename T, typename FieldT, multi_exp_method Method>
T multi_exp_with_mixed_addition(typename std::vector<T>::const_iterator vec_start,
                                typename std::vector<T>::const_iterator vec_end,
                                typename std::vector<FieldT>::const_iterator scalar_start,
                                typename std::vector<FieldT>::const_iterator scalar_end, const size_t chunks)
{
    return b200_detail::msm_dispatch<T, FieldT, Method>::mixed(vec_start, vec_end, scalar_start, scalar_end, chunks);
}

template <typename T>
window_table<T> get_window_table(const size_t scalar_size, const size_t window, const T &g)
{
    return b200_detail::table_dispatch<T>::make(scalar_size, window, g);
}

template <typename T, typename FieldT>
T windowed_exp(const size_t scalar_size, const size_t window, const window_table<T> &powers_of_g, const FieldT &pow)
{
    return b200_detail::table_dispatch<T>::template one_exp<FieldT>(scalar_size, window, powers_of_g, pow);
}

template <typename T, typename FieldT>
std::vector<T> batch_exp(const size_t scalar_size, const size_t window, const window_table<T> &table,
                         const std::vector<FieldT> &v)
{
    return b200_detail::table_dispatch<T>::template many<FieldT>(scalar_size, window, table, nullptr, v);
}

template <typename T, typename FieldT>
std::vector<T> batch_exp_with_coeff(const size_t scalar_size, const size_t window, const window_table<T> &table,
                                    const FieldT &coeff, const std::vector<FieldT> &v)
{
    return b200_detail::table_dispatch<T>::template many<FieldT>(scalar_size, window, table, &coeff, v);
}

template <typename T>
void batch_to_special(std::vector<T> &vec)
{
    b200_detail::table_dispatch<T>::special(vec);
}

}  // namespace libff
#endif  // B200_SHIM_MULTIEXP_HPP_
