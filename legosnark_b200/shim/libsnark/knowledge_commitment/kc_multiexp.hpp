// Shadow of libsnark's kc_multiexp.hpp (SNK/knowledge_commitment/kc_multiexp.hpp:28-63,
// kc_multiexp.tcc:21-89, 122-199; SNK = depends/libsnark/libsnark): with
// `legosnark_b200/shim` ahead of libsnark on the include path, Groth16's B-query work
//   kc_multi_exp_with_mixed_addition<G2, G1, Fr, Method>   (r1cs_gg_ppzksnark.tcc:453-463, prover)
//   kc_batch_exp<G2, G1, Fr>                               (r1cs_gg_ppzksnark.tcc:332, generator)
// runs on the B200 engine for the alt_bn128 / bn128 groups; every other (T1, T2) keeps the
// reference's template, which stays available as libsnark::libsnark_cpu_*.
//
// A knowledge commitment is a pair (g in T1, h in T2) with component-wise group law
// (knowledge_commitment.hpp:34-67), so its MSM is one G2 MSM and one G1 MSM over the same
// scalars (one upload, b200_msm_g2g1), and its fixed-base batch is one batch_exp per component.
// The reference walks a sparse_vector and calls windowed_exp per element from inside an OpenMP
// region (kc_multiexp.tcc:173-183); here the non-zero positions are gathered first and each
// component is one batched device call ("handled by batching in the shim").
#ifndef B200_SHIM_KC_MULTIEXP_HPP_
#define B200_SHIM_KC_MULTIEXP_HPP_

#define kc_multi_exp_with_mixed_addition libsnark_cpu_kc_multi_exp_with_mixed_addition
#define kc_batch_exp libsnark_cpu_kc_batch_exp
#include_next <libsnark/knowledge_commitment/kc_multiexp.hpp>
#undef kc_multi_exp_with_mixed_addition
#undef kc_batch_exp

#include <libff/algebra/scalar_multiplication/multiexp.hpp>  // the shim's (same include path)

namespace libsnark {
namespace b200_detail {

template <typename T1, typename T2>
struct kc_on_gpu {
    static const bool g2g1 = b200shim::group_traits<T1>::supported && b200shim::group_traits<T2>::supported;
    template <typename A, typename B, bool Ok>
    struct order {
        static const bool value = false;
    };
    template <typename A, typename B>
    struct order<A, B, true> {
        static const bool value = b200shim::group_traits<A>::group == 1 && b200shim::group_traits<B>::group == 0;
    };
    static const bool value = order<T1, T2, g2g1>::value;
};

template <typename T1, typename T2, typename FieldT, libff::multi_exp_method Method, bool OnGpu = kc_on_gpu<T1, T2>::value>
struct kc_dispatch {
    typedef typename std::vector<FieldT>::const_iterator SI;
    static knowledge_commitment<T1, T2> msm(const knowledge_commitment_vector<T1, T2> &vec, const size_t min_idx, const size_t max_idx,
                                            SI ss, SI se, const size_t chunks)
    {
        return libsnark_cpu_kc_multi_exp_with_mixed_addition<T1, T2, FieldT, Method>(vec, min_idx, max_idx, ss, se, chunks);
    }
};

template <typename T1, typename T2, typename FieldT, libff::multi_exp_method Method>
struct kc_dispatch<T1, T2, FieldT, Method, true> {
    typedef typename std::vector<FieldT>::const_iterator SI;
    static knowledge_commitment<T1, T2> msm(const knowledge_commitment_vector<T1, T2> &vec, const size_t min_idx, const size_t max_idx,
                                            SI ss, SI se, const size_t)
    {
        // the index walk of kc_multiexp.tcc:29-83; zero scalars are dropped here (they would be
        // skipped digit by digit on the device), ones and everything else go to the engine
        auto index_it = std::lower_bound(vec.indices.begin(), vec.indices.end(), min_idx);
        auto value_it = vec.values.begin() + (index_it - vec.indices.begin());
        const size_t scalar_length = (size_t)std::distance(ss, se);
        const size_t upper = (size_t)(vec.indices.end() - index_it);
        std::vector<T1> g;
        std::vector<T2> h;
        std::vector<FieldT> p;
        g.reserve(upper);
        h.reserve(upper);
        p.reserve(upper);
        for (; index_it != vec.indices.end() && *index_it < max_idx; ++index_it, ++value_it) {
            const size_t pos = *index_it - min_idx;
            if (pos >= scalar_length) throw std::runtime_error("kc_multi_exp: index outside the scalar range");
            const FieldT &s = *(ss + pos);
            if (s.is_zero()) continue;
            p.push_back(s);
            g.push_back(value_it->g);
            h.push_back(value_it->h);
        }
        knowledge_commitment<T1, T2> res;
        b200shim::msm_pair<T1, T2, FieldT>(g.data(), h.data(), p.data(), p.size(), res.g, res.h);
        return res;
    }
};

template <typename T1, typename T2, typename FieldT, bool OnGpu = kc_on_gpu<T1, T2>::value>
struct kc_table_dispatch {
    static knowledge_commitment_vector<T1, T2> many(const size_t scalar_size, const size_t w1, const size_t w2,
                                                    const libff::window_table<T1> &t1, const libff::window_table<T2> &t2,
                                                    const FieldT &c1, const FieldT &c2, const std::vector<FieldT> &v, const size_t chunks)
    {
        return libsnark_cpu_kc_batch_exp<T1, T2, FieldT>(scalar_size, w1, w2, t1, t2, c1, c2, v, chunks);
    }
};

template <typename T1, typename T2, typename FieldT>
struct kc_table_dispatch<T1, T2, FieldT, true> {
    static knowledge_commitment_vector<T1, T2> many(const size_t, const size_t, const size_t, const libff::window_table<T1> &t1,
                                                    const libff::window_table<T2> &t2, const FieldT &c1, const FieldT &c2,
                                                    const std::vector<FieldT> &v, const size_t)
    {
        knowledge_commitment_vector<T1, T2> res;
        res.domain_size_ = v.size();
        std::vector<FieldT> nz;
        for (size_t i = 0; i < v.size(); i++)
            if (!v[i].is_zero()) {  // kc_multiexp.tcc:106-111: zero coordinates are not stored
                nz.push_back(v[i]);
                res.indices.push_back(i);
            }
        // row 0 of any window table is {0, g, 2g, ...} (multiexp.tcc:563-578)
        const std::vector<T1> g = b200shim::fixed_base_exp<T1, FieldT>(t1[0][1], nz, &c1);
        const std::vector<T2> h = b200shim::fixed_base_exp<T2, FieldT>(t2[0][1], nz, &c2);
        res.values.reserve(nz.size());
        for (size_t i = 0; i < nz.size(); i++) res.values.emplace_back(knowledge_commitment<T1, T2>(g[i], h[i]));
        return res;
    }
};

}  // namespace b200_detail

template <typename T1, typename T2, typename FieldT, libff::multi_exp_method Method>
knowledge_commitment<T1, T2> kc_multi_exp_with_mixed_addition(const knowledge_commitment_vector<T1, T2> &vec, const size_t min_idx,
                                                              const size_t max_idx,
                                                              typename std::vector<FieldT>::const_iterator scalar_start,
                                                              typename std::vector<FieldT>::const_iterator scalar_end,
                                                              const size_t chunks)
{
    return b200_detail::kc_dispatch<T1, T2, FieldT, Method>::msm(vec, min_idx, max_idx, scalar_start, scalar_end, chunks);
}

template <typename T1, typename T2, typename FieldT>
knowledge_commitment_vector<T1, T2> kc_batch_exp(const size_t scalar_size, const size_t T1_window, const size_t T2_window,
                                                 const libff::window_table<T1> &T1_table, const libff::window_table<T2> &T2_table,
                                                 const FieldT &T1_coeff, const FieldT &T2_coeff, const std::vector<FieldT> &v,
                                                 const size_t suggested_num_chunks)
{
    return b200_detail::kc_table_dispatch<T1, T2, FieldT>::many(scalar_size, T1_window, T2_window, T1_table, T2_table, T1_coeff,
                                                                T2_coeff, v, suggested_num_chunks);
}

}  // namespace libsnark
#endif  // B200_SHIM_KC_MULTIEXP_HPP_
