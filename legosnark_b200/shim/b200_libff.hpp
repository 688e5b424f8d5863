// b200_libff.hpp — C++ host side of the drop-in: libff group / field objects <-> the
// C-ABI of include/b200_msm.h.
//
// The reference has no FFI; its boundary is the set of function templates in
//   LFF/algebra/scalar_multiplication/multiexp.hpp:56-127
// (LFF = depends/libsnark/depends/libff/libff) instantiated in the caller's TU.  The
// shadow header shim/libff/algebra/scalar_multiplication/multiexp.hpp keeps those
// templates' names and signatures and forwards the four concrete groups
//   alt_bn128_G1 / alt_bn128_G2  (LFF/algebra/curves/alt_bn128/alt_bn128_g1.hpp:35, _g2.hpp:36)
//   bn128_G1     / bn128_G2      (LFF/algebra/curves/bn128/bn128_g1.hpp:36, bn128_g2.hpp:37)
// with FieldT == T::scalar_field to the functions below.  Every other T (e.g. libsnark's
// knowledge_commitment<T1,T2>) keeps the reference's template.
//
// The in-memory images are passed as they are: a point is 3 (G1) or 6 (G2) Montgomery
// field elements of 4 x u64 limbs, R = 2^256, Jacobian X|Y|Z, for both curve flavours
// (identical limbs).  Failure of the engine (no CUDA device, CUDA
// error) throws std::runtime_error: there is no CPU fallback behind this header.
#ifndef B200_LIBFF_HPP_
#define B200_LIBFF_HPP_

#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <istream>
#include <ostream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "b200_msm.h"

namespace libff {
class alt_bn128_G1;
class alt_bn128_G2;
class bn128_G1;
class bn128_G2;
}  // namespace libff

namespace b200shim {

template <typename T>
struct group_traits {
    static const bool supported = false;
};
template <>
struct group_traits<libff::alt_bn128_G1> {
    static const bool supported = true;
    static const int group = 0;
    static const size_t limbs = 12;
};
template <>
struct group_traits<libff::bn128_G1> {
    static const bool supported = true;
    static const int group = 0;
    static const size_t limbs = 12;
};
template <>
struct group_traits<libff::alt_bn128_G2> {
    static const bool supported = true;
    static const int group = 1;
    static const size_t limbs = 24;
};
template <>
struct group_traits<libff::bn128_G2> {
    static const bool supported = true;
    static const int group = 1;
    static const size_t limbs = 24;
};

inline void check(int rc, const char *what)
{
    if (rc != B200_OK) throw std::runtime_error(std::string(what) + " failed: " + b200_last_error());
}

// One engine per process, created on first use.  B200_GPUS=<n> limits the devices an MSM is
// sharded over (default: all visible), mirroring OMP_NUM_THREADS for the reference's chunks.
inline void ensure_init()
{
    static const bool once = [] {
        const char *e = std::getenv("B200_GPUS");
        check(b200_init(e ? std::atoi(e) : 0), "b200_init");
        return true;
    }();
    (void)once;
}

// Optional: call first thing in main().  The engine comes up on a background thread while the program builds its circuit /
// reads its inputs; the first shim call then finds it ready (or waits for the rest).  Same B200_GPUS rule as ensure_init().
inline void start_engine_early()
{
    const char *e = std::getenv("B200_GPUS");
    check(b200_init_async(e ? std::atoi(e) : 0), "b200_init_async");
}

template <typename T>
inline const uint64_t *limbs_of(const T *p)
{
    static_assert(sizeof(T) == group_traits<T>::limbs * 8, "unexpected point layout");
    return reinterpret_cast<const uint64_t *>(p);
}

// engine output (x, y, 1) / (0, 1, 0)  ->  T; the zero is written as the curve's own
// T::zero() (alt_bn128: (0,1,0), alt_bn128_init.cpp:145-147; bn128: (1,1,0), bn128_init.cpp:101-103)
template <typename T>
inline T point_from_limbs(const uint64_t *l)
{
    const size_t L = group_traits<T>::limbs, zoff = 2 * L / 3;
    uint64_t z = 0;
    for (size_t i = zoff; i < L; i++) z |= l[i];
    if (z == 0) return T::zero();
    T r;
    std::memcpy(reinterpret_cast<void *>(&r), l, L * 8);
    return r;
}

template <typename T, typename FieldT>
inline T msm(const T *bases, const FieldT *scalars, size_t n)
{
    static_assert(std::is_same<FieldT, typename T::scalar_field>::value, "scalars must be the group's Fr");
    static_assert(sizeof(FieldT) == 32, "unexpected scalar layout");
    ensure_init();
    uint64_t out[24];
    const uint64_t *b = n ? limbs_of(bases) : nullptr;
    const uint64_t *s = n ? reinterpret_cast<const uint64_t *>(scalars) : nullptr;
    if (group_traits<T>::group == 0) check(b200_msm_g1(b, s, n, out), "b200_msm_g1");
    else check(b200_msm_g2(b, s, n, out), "b200_msm_g2");
    return point_from_limbs<T>(out);
}

// out[j] = sum over [offsets[j], offsets[j+1]) of scalars[i] * bases[i]: many small MSMs submitted together
// (b200_msm_batch_*).  The caller of LegoSNARK's mtxmultiexp is bound to it by shim/sparsemexp.h.
template <typename T, typename FieldT>
inline std::vector<T> msm_batch(const std::vector<T> &bases, const std::vector<FieldT> &scalars, const std::vector<uint64_t> &offsets)
{
    static_assert(std::is_same<FieldT, typename T::scalar_field>::value && sizeof(FieldT) == 32, "scalars must be the group's Fr");
    if (offsets.empty() || offsets.back() != bases.size() || bases.size() != scalars.size())
        throw std::runtime_error("msm_batch: bases, scalars and offsets disagree");
    ensure_init();
    const size_t L = group_traits<T>::limbs, count = offsets.size() - 1;
    std::vector<uint64_t> out(count * L);
    const uint64_t *b = bases.empty() ? nullptr : limbs_of(bases.data());
    const uint64_t *s = scalars.empty() ? nullptr : reinterpret_cast<const uint64_t *>(scalars.data());
    check(group_traits<T>::group == 0 ? b200_msm_batch_g1(b, s, offsets.data(), count, out.data())
                                      : b200_msm_batch_g2(b, s, offsets.data(), count, out.data()), "b200_msm_batch");
    std::vector<T> res(count, T::zero());
    for (size_t j = 0; j < count; j++) res[j] = point_from_limbs<T>(out.data() + j * L);
    return res;
}

// (sum s_i g_i, sum s_i h_i) for g_i in a G2 group and h_i in a G1 group over one scalar vector:
// the MSM behind libsnark's knowledge_commitment<T1,T2> (kc_multiexp.tcc:21-89).
template <typename T1, typename T2, typename FieldT>
inline void msm_pair(const T1 *g, const T2 *h, const FieldT *scalars, size_t n, T1 &out_g, T2 &out_h)
{
    static_assert(group_traits<T1>::group == 1 && group_traits<T2>::group == 0, "pair must be (G2, G1)");
    static_assert(std::is_same<FieldT, typename T1::scalar_field>::value && sizeof(FieldT) == 32, "scalars must be the groups' Fr");
    ensure_init();
    uint64_t o2[24], o1[12];
    check(b200_msm_g2g1(n ? limbs_of(g) : nullptr, n ? limbs_of(h) : nullptr,
                        n ? reinterpret_cast<const uint64_t *>(scalars) : nullptr, n, o2, o1), "b200_msm_g2g1");
    out_g = point_from_limbs<T1>(o2);
    out_h = point_from_limbs<T2>(o1);
}

// The device table of the last base used per group is kept and reused: libsnark's Groth16
// generator runs five batch_exps (and kc_batch_exp) over one G1 and one G2 table
// (r1cs_gg_ppzksnark.tcc:296-360).  A table built for far fewer scalars than a later call brings
// (its window grows with the batch size) is rebuilt.
struct table_cache {
    uint64_t handle = 0;
    size_t built_for = 0;
    uint64_t base[24] = {0};  // no destructor: the engine frees its tables at shutdown / process exit
};

template <typename T, typename FieldT>
inline std::vector<T> fixed_base_exp(const T &base, const std::vector<FieldT> &v, const FieldT *coeff)
{
    static_assert(std::is_same<FieldT, typename T::scalar_field>::value, "scalars must be the group's Fr");
    ensure_init();
    const size_t L = group_traits<T>::limbs, n = v.size();
    const bool g1 = group_traits<T>::group == 0;
    std::vector<T> res(n, T::zero());
    if (n == 0) return res;
    static table_cache cache[2];
    table_cache &tc = cache[group_traits<T>::group];
    if (!tc.handle || std::memcmp(tc.base, limbs_of(&base), L * 8) != 0 || n > 4 * tc.built_for) {
        if (tc.handle) b200_window_table_destroy(tc.handle);
        tc.handle = 0;
        check(g1 ? b200_window_table_create_g1(limbs_of(&base), n, &tc.handle)
                 : b200_window_table_create_g2(limbs_of(&base), n, &tc.handle), "b200_window_table_create");
        std::memcpy(tc.base, limbs_of(&base), L * 8);
        tc.built_for = n;
    }
    std::vector<uint64_t> out(n * L);
    const uint64_t *s = reinterpret_cast<const uint64_t *>(v.data());
    const uint64_t *c = coeff ? reinterpret_cast<const uint64_t *>(coeff) : nullptr;
    check(g1 ? b200_batch_exp_table_g1(tc.handle, s, n, c, out.data()) : b200_batch_exp_table_g2(tc.handle, s, n, c, out.data()),
          "b200_batch_exp_table");
    for (size_t i = 0; i < n; i++) res[i] = point_from_limbs<T>(out.data() + i * L);
    return res;
}

template <typename T>
inline void to_special(std::vector<T> &vec)
{
    ensure_init();
    const size_t L = group_traits<T>::limbs, n = vec.size();
    if (n == 0) return;
    std::vector<uint64_t> buf(n * L);
    std::memcpy(buf.data(), limbs_of(vec.data()), n * L * 8);
    if (group_traits<T>::group == 0) check(b200_batch_to_affine_g1(buf.data(), n), "b200_batch_to_affine_g1");
    else check(b200_batch_to_affine_g2(buf.data(), n), "b200_batch_to_affine_g2");
    for (size_t i = 0; i < n; i++) vec[i] = point_from_limbs<T>(buf.data() + i * L);
}

// ---- Fr-side routines next to the MSMs: LegoSNARK callers -------------
// MultiVPolyT::evalMLE (LS/prototools/polytools.h:207-234) on the device
template <typename FieldT>
inline FieldT eval_mle(const std::vector<FieldT> &v, const std::vector<FieldT> &r)
{
    static_assert(sizeof(FieldT) == 32, "unexpected scalar layout");
    if (v.size() != ((size_t)1 << r.size())) throw std::runtime_error("evalMLE: expected N == 1 << d");  // polytools.h:211
    ensure_init();
    FieldT out;
    check(b200_fr_eval_mle(reinterpret_cast<const uint64_t *>(v.data()), reinterpret_cast<const uint64_t *>(r.data()), r.size(),
                           reinterpret_cast<uint64_t *>(&out)), "b200_fr_eval_mle");
    return out;
}

// A commitment key kept on the device (CommScheme's g1s / g2s, LS/prototools/commit.h:129-147): uploaded and
// normalised once; msm() is then CommScheme::commit's multiExpMA(g1s, v) / multiExpMA(g2s, v) (commit.h:149-158)
// moving only the scalars.  precompute() adds the window multiples (b200_key_precompute_*).
template <typename T>
struct resident_key {
    uint64_t handle = 0;
    size_t n = 0;
    explicit resident_key(const std::vector<T> &bases) : n(bases.size())
    {
        ensure_init();
        check(group_traits<T>::group == 0 ? b200_pin_bases_g1(n ? limbs_of(bases.data()) : nullptr, n, &handle)
                                          : b200_pin_bases_g2(n ? limbs_of(bases.data()) : nullptr, n, &handle),
              "b200_pin_bases");
    }
    ~resident_key()
    {
        if (handle) b200_unpin_bases(handle);
    }
    resident_key(const resident_key &) = delete;
    resident_key &operator=(const resident_key &) = delete;

    void precompute(uint32_t window_bits = 0)
    {
        check(group_traits<T>::group == 0 ? b200_key_precompute_g1(handle, window_bits) : b200_key_precompute_g2(handle, window_bits),
              "b200_key_precompute");
    }
    // sum_i v[i] * bases[i] over the first min(n, v.size()) bases (multiExpMA's n = min(sizes), LS/utils/globl.h:63-78)
    template <typename FieldT>
    T msm(const std::vector<FieldT> &v) const
    {
        static_assert(std::is_same<FieldT, typename T::scalar_field>::value && sizeof(FieldT) == 32, "scalars must be the group's Fr");
        const size_t m = v.size() < n ? v.size() : n;
        uint64_t out[24];
        const uint64_t *s = m ? reinterpret_cast<const uint64_t *>(v.data()) : nullptr;
        check(group_traits<T>::group == 0 ? b200_msm_pinned_g1(handle, 0, s, m, out) : b200_msm_pinned_g2(handle, 0, s, m, out),
              "b200_msm_pinned");
        return point_from_limbs<T>(out);
    }
};

// CPPoly::prove (LS/gadgets/poly.h:45-91): witness[i] for i < d; witnessa[i] (i >= 1) equals witness[i]
// in the reference too (same MSM over the same bases, poly.h:84-86).
template <typename T, typename FieldT>
inline std::vector<T> cppoly_prove(const resident_key<T> &key, const std::vector<FieldT> &v, const std::vector<FieldT> &r)
{
    static_assert(std::is_same<FieldT, typename T::scalar_field>::value && sizeof(FieldT) == 32, "scalars must be the group's Fr");
    static_assert(group_traits<T>::group == 0, "CPPoly::prove commits in G1");
    if (b200_device_count() != 1) throw std::runtime_error("cppoly_prove: b200_cppoly_prove_g1 needs a single-device engine (set B200_GPUS=1)");
    const size_t d = r.size();
    if (v.size() != ((size_t)1 << d)) throw std::runtime_error("CPPoly::prove: expected v.size() == 1 << d");
    std::vector<uint64_t> out(12 * d);
    check(b200_cppoly_prove_g1(key.handle, reinterpret_cast<const uint64_t *>(v.data()), reinterpret_cast<const uint64_t *>(r.data()), d,
                               out.data(), nullptr), "b200_cppoly_prove_g1");
    std::vector<T> w(d, T::zero());
    for (size_t i = 0; i < d; i++) w[i] = point_from_limbs<T>(out.data() + 12 * i);
    return w;
}

// ---- wire format: std::vector<G> <-> the reference's stream image ---------------
// Byte-identical to `out << vec` / `in >> vec` of the reference (alt_bn128_g1.cpp:461-497, bn128_g1.cpp:465-492 and
// the G2 twins) for builds with -DBINARY_OUTPUT and point compression (LegoSNARK's configuration): the size line,
// then per point  '0'|'1' (is_zero), X raw bytes, '0'|'1' (Y bit).  The normalisation, Montgomery conversions and
// the square roots run on the device (b200_compress_* / b200_decompress_*); this is only the framing.
#if defined(BINARY_OUTPUT) && !defined(NO_PT_COMPRESSION)
template <typename T>
struct wire_traits {
    // bn128 writes the raw Montgomery image and takes the Y bit from it; alt_bn128 goes through as_bigint unless
    // -DMONTGOMERY_OUTPUT (fp.tcc operator<<)
    static int flavour()
    {
        if (std::is_same<T, libff::bn128_G1>::value || std::is_same<T, libff::bn128_G2>::value) return 2;
#ifdef MONTGOMERY_OUTPUT
        return 1;
#else
        return 0;
#endif
    }
};

template <typename T>
inline void write_points(std::ostream &out, const std::vector<T> &v)
{
    ensure_init();
    const size_t L = group_traits<T>::limbs, xb = L / 3 * 8, n = v.size();
    out << n << "\n";
    if (n == 0) return;
    std::vector<uint64_t> x(n * L / 3);
    std::vector<uint8_t> flags(n);
    check(group_traits<T>::group == 0 ? b200_compress_g1(limbs_of(v.data()), n, wire_traits<T>::flavour(), x.data(), flags.data())
                                      : b200_compress_g2(limbs_of(v.data()), n, wire_traits<T>::flavour(), x.data(), flags.data()),
          "b200_compress");
    std::string buf(n * (xb + 2), '0');
    for (size_t i = 0; i < n; i++) {
        char *p = &buf[i * (xb + 2)];
        p[0] = (flags[i] & 2) ? '1' : '0';
        std::memcpy(p + 1, reinterpret_cast<const char *>(x.data()) + i * xb, xb);
        p[xb + 1] = (flags[i] & 1) ? '1' : '0';
    }
    out.write(buf.data(), (std::streamsize)buf.size());
}

template <typename T>
inline void read_points(std::istream &in, std::vector<T> &v)
{
    ensure_init();
    const size_t L = group_traits<T>::limbs, xb = L / 3 * 8;
    size_t n = 0;
    in >> n;
    in.get();  // the newline after the size (consume_newline)
    v.assign(n, T::zero());
    if (n == 0) return;
    std::string buf(n * (xb + 2), '\0');
    in.read(&buf[0], (std::streamsize)buf.size());
    if ((size_t)in.gcount() != buf.size()) throw std::runtime_error("read_points: stream ended early");
    std::vector<uint64_t> x(n * L / 3), pts(n * L);
    std::vector<uint8_t> flags(n);
    for (size_t i = 0; i < n; i++) {
        const char *p = &buf[i * (xb + 2)];
        flags[i] = (uint8_t)((p[0] == '1' ? 2 : 0) | (p[xb + 1] == '1' ? 1 : 0));
        std::memcpy(reinterpret_cast<char *>(x.data()) + i * xb, p + 1, xb);
    }
    check(group_traits<T>::group == 0 ? b200_decompress_g1(x.data(), flags.data(), n, wire_traits<T>::flavour(), pts.data(), nullptr)
                                      : b200_decompress_g2(x.data(), flags.data(), n, wire_traits<T>::flavour(), pts.data(), nullptr),
          "b200_decompress");
    for (size_t i = 0; i < n; i++) v[i] = point_from_limbs<T>(pts.data() + i * L);
}
#endif  // BINARY_OUTPUT && point compression

}  // namespace b200shim
#endif  // B200_LIBFF_HPP_
