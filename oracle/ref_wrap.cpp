// oracle/_ref C-ABI wrapper around the UNMODIFIED reference (libff) sources.
//
// TEST INFRASTRUCTURE ONLY.  This translation unit contains no reference code:
// it includes the reference headers where they lie under /root/reference and
// calls their public API (libff::multi_exp, multi_exp_with_mixed_addition,
// batch_exp, batch_to_special, group/field operators).  It is compiled by
// oracle/Makefile into oracle/_ref/libffref.so, which is used to
//   * pin the plain-C restatement in oracle/bn254_oracle.c,
//   * generate tests/golden/ fixtures (tools/make_golden.py),
//   * serve as the `--impl reference` / cpu_baseline arm of bench.py.
// It is never loaded by the product library.
//
// All buffers are little-endian u64 limbs in Montgomery form (R = 2^256),
// exactly the in-memory representation of libff's Fp_model::mont_repr
// (LFF/algebra/fields/fp.hpp:42) and of bn::Fp (ATE/include/zm2.h:266).
// G1 = X|Y|Z (12 limbs), G2 = X.c0|X.c1|Y.c0|Y.c1|Z.c0|Z.c1 (24 limbs).
#include <cstdint>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include <libff/algebra/curves/alt_bn128/alt_bn128_pp.hpp>
#include <libff/algebra/scalar_multiplication/multiexp.hpp>
#include <libff/common/profiling.hpp>
#include <libff/common/rng.hpp>
#ifdef REF_WITH_BN128
#include <libff/algebra/curves/bn128/bn128_pp.hpp>
#endif

using namespace libff;

namespace {

bool g_init = false;

// ---- raw limb <-> libff object marshalling -------------------------------
inline void load(alt_bn128_Fq &f, const uint64_t *p) { memcpy(f.mont_repr.data, p, 32); }
inline void store(uint64_t *p, const alt_bn128_Fq &f) { memcpy(p, f.mont_repr.data, 32); }
inline void load(alt_bn128_Fr &f, const uint64_t *p) { memcpy(f.mont_repr.data, p, 32); }
inline void store(uint64_t *p, const alt_bn128_Fr &f) { memcpy(p, f.mont_repr.data, 32); }
inline void load(alt_bn128_Fq2 &f, const uint64_t *p) { load(f.c0, p); load(f.c1, p + 4); }
inline void store(uint64_t *p, const alt_bn128_Fq2 &f) { store(p, f.c0); store(p + 4, f.c1); }

inline void load(alt_bn128_G1 &g, const uint64_t *p) { load(g.X, p); load(g.Y, p + 4); load(g.Z, p + 8); }
inline void store(uint64_t *p, const alt_bn128_G1 &g) { store(p, g.X); store(p + 4, g.Y); store(p + 8, g.Z); }
inline void load(alt_bn128_G2 &g, const uint64_t *p) { load(g.X, p); load(g.Y, p + 8); load(g.Z, p + 16); }
inline void store(uint64_t *p, const alt_bn128_G2 &g) { store(p, g.X); store(p + 8, g.Y); store(p + 16, g.Z); }

#ifdef REF_WITH_BN128
inline void load(bn128_Fr &f, const uint64_t *p) { memcpy(f.mont_repr.data, p, 32); }
inline void load(bn128_G1 &g, const uint64_t *p) { memcpy((void *)&g.coord[0], p, 96); }
inline void store(uint64_t *p, const bn128_G1 &g) { memcpy(p, (const void *)&g.coord[0], 96); }
inline void load(bn128_G2 &g, const uint64_t *p) { memcpy((void *)&g.coord[0], p, 192); }
inline void store(uint64_t *p, const bn128_G2 &g) { memcpy(p, (const void *)&g.coord[0], 192); }
#endif

template <typename G> constexpr size_t limbs_of();
template <> constexpr size_t limbs_of<alt_bn128_G1>() { return 12; }
template <> constexpr size_t limbs_of<alt_bn128_G2>() { return 24; }
#ifdef REF_WITH_BN128
template <> constexpr size_t limbs_of<bn128_G1>() { return 12; }
template <> constexpr size_t limbs_of<bn128_G2>() { return 24; }
#endif

template <typename G, typename Fr>
int msm_impl(const uint64_t *bases, const uint64_t *scalars, size_t n, size_t chunks, int variant,
             int normalise, uint64_t *out)
{
    const size_t L = limbs_of<G>();
    std::vector<G> g(n);
    std::vector<Fr> s(n);
    for (size_t i = 0; i < n; i++) {
        load(g[i], bases + i * L);
        load(s[i], scalars + i * 4);
    }
    G r;
    switch (variant) {
    case 0: r = multi_exp<G, Fr, multi_exp_method_BDLO12>(g.begin(), g.end(), s.begin(), s.end(), chunks); break;
    case 1: r = multi_exp_with_mixed_addition<G, Fr, multi_exp_method_BDLO12>(g.begin(), g.end(), s.begin(), s.end(), chunks); break;
    case 2: r = multi_exp<G, Fr, multi_exp_method_bos_coster>(g.begin(), g.end(), s.begin(), s.end(), chunks); break;
    case 3: r = multi_exp<G, Fr, multi_exp_method_naive>(g.begin(), g.end(), s.begin(), s.end(), chunks); break;
    case 4: r = multi_exp<G, Fr, multi_exp_method_naive_plain>(g.begin(), g.end(), s.begin(), s.end(), chunks); break;
    default: return 1;
    }
    if (normalise) r.to_affine_coordinates();
    store(out, r);
    return 0;
}

template <typename G, typename Fr>
int batch_exp_impl(const uint64_t *base, const uint64_t *scalars, size_t n, const uint64_t *coeff,
                   size_t window, int normalise, uint64_t *out)
{
    const size_t L = limbs_of<G>();
    G g;
    load(g, base);
    std::vector<Fr> s(n);
    for (size_t i = 0; i < n; i++) load(s[i], scalars + i * 4);
    const size_t scalar_bits = Fr::size_in_bits();
    if (window == 0) window = get_exp_window_size<G>(n);
    window_table<G> table = get_window_table(scalar_bits, window, g);
    std::vector<G> res;
    if (coeff) {
        Fr c;
        load(c, coeff);
        res = batch_exp_with_coeff(scalar_bits, window, table, c, s);
    } else {
        res = batch_exp(scalar_bits, window, table, s);
    }
    if (normalise) batch_to_special(res);
    for (size_t i = 0; i < n; i++) store(out + i * L, res[i]);
    return 0;
}

template <typename G>
int to_special_impl(uint64_t *pts, size_t n)
{
    const size_t L = limbs_of<G>();
    std::vector<G> v(n);
    for (size_t i = 0; i < n; i++) load(v[i], pts + i * L);
    batch_to_special(v);
    for (size_t i = 0; i < n; i++) store(pts + i * L, v[i]);
    return 0;
}

// op: 0 add(operator+), 1 mixed_add, 2 dbl, 3 to_affine, 4 neg, 5 add()
template <typename G>
int group_op_impl(int op, const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out)
{
    const size_t L = limbs_of<G>();
    for (size_t i = 0; i < n; i++) {
        G p, q, r;
        load(p, a + i * L);
        if (b) load(q, b + i * L);
        switch (op) {
        case 0: r = p + q; break;
        case 1: r = p.mixed_add(q); break;
        case 2: r = p.dbl(); break;
        case 3: r = p; r.to_affine_coordinates(); break;
        case 4: r = -p; break;
        case 5: r = p.add(q); break;
        default: return 1;
        }
        store(out + i * L, r);
    }
    return 0;
}

template <typename G, typename Fr>
int scalar_mul_impl(const uint64_t *base, const uint64_t *scalars, size_t n, int stride_base,
                    int normalise, uint64_t *out)
{
    const size_t L = limbs_of<G>();
#ifdef _OPENMP
#pragma omp parallel for
#endif
    for (size_t i = 0; i < n; i++) {
        G g, r;
        Fr s;
        load(g, base + (stride_base ? i * L : 0));
        load(s, scalars + i * 4);
        r = s * g;
        if (normalise) r.to_affine_coordinates();
        store(out + i * L, r);
    }
    return 0;
}

// op: 0 mul, 1 squared, 2 add, 3 sub, 4 inverse, 5 neg, 6 from-bigint (Fp(b)), 7 as_bigint
template <typename F>
int field_op_impl(int op, const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out, size_t L)
{
    for (size_t i = 0; i < n; i++) {
        F x, y, r;
        load(x, a + i * L);
        if (b) load(y, b + i * L);
        switch (op) {
        case 0: r = x * y; break;
        case 1: r = x.squared(); break;
        case 2: r = x + y; break;
        case 3: r = x - y; break;
        case 4: r = x.inverse(); break;
        case 5: r = -x; break;
        default: return 1;
        }
        store(out + i * L, r);
    }
    return 0;
}

} // namespace

extern "C" {

int ref_init(void)
{
    if (g_init) return 0;
    inhibit_profiling_info = true;
    inhibit_profiling_counters = true;
    alt_bn128_pp::init_public_params();
#ifdef REF_WITH_BN128
    bn128_pp::init_public_params();
#endif
    g_init = true;
    return 0;
}

int ref_has_bn128(void)
{
#ifdef REF_WITH_BN128
    return 1;
#else
    return 0;
#endif
}

int ref_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// curve: 0 alt_bn128, 1 bn128.  variant: 0 multi_exp<BDLO12>, 1 multi_exp_with_mixed_addition<BDLO12>,
// 2 bos_coster, 3 naive, 4 naive_plain.
int ref_msm_g1(int curve, const uint64_t *bases, const uint64_t *scalars, size_t n, size_t chunks,
               int variant, int normalise, uint64_t *out)
{
    ref_init();
    if (curve == 0) return msm_impl<alt_bn128_G1, alt_bn128_Fr>(bases, scalars, n, chunks, variant, normalise, out);
#ifdef REF_WITH_BN128
    if (curve == 1) return msm_impl<bn128_G1, bn128_Fr>(bases, scalars, n, chunks, variant, normalise, out);
#endif
    return 2;
}

int ref_msm_g2(int curve, const uint64_t *bases, const uint64_t *scalars, size_t n, size_t chunks,
               int variant, int normalise, uint64_t *out)
{
    ref_init();
    if (curve == 0) return msm_impl<alt_bn128_G2, alt_bn128_Fr>(bases, scalars, n, chunks, variant, normalise, out);
#ifdef REF_WITH_BN128
    if (curve == 1) return msm_impl<bn128_G2, bn128_Fr>(bases, scalars, n, chunks, variant, normalise, out);
#endif
    return 2;
}

int ref_batch_exp_g1(int curve, const uint64_t *base, const uint64_t *scalars, size_t n,
                     const uint64_t *coeff, size_t window, int normalise, uint64_t *out)
{
    ref_init();
    if (curve == 0) return batch_exp_impl<alt_bn128_G1, alt_bn128_Fr>(base, scalars, n, coeff, window, normalise, out);
#ifdef REF_WITH_BN128
    if (curve == 1) return batch_exp_impl<bn128_G1, bn128_Fr>(base, scalars, n, coeff, window, normalise, out);
#endif
    return 2;
}

int ref_batch_exp_g2(int curve, const uint64_t *base, const uint64_t *scalars, size_t n,
                     const uint64_t *coeff, size_t window, int normalise, uint64_t *out)
{
    ref_init();
    if (curve == 0) return batch_exp_impl<alt_bn128_G2, alt_bn128_Fr>(base, scalars, n, coeff, window, normalise, out);
#ifdef REF_WITH_BN128
    if (curve == 1) return batch_exp_impl<bn128_G2, bn128_Fr>(base, scalars, n, coeff, window, normalise, out);
#endif
    return 2;
}

size_t ref_exp_window_size_g1(size_t n) { ref_init(); return get_exp_window_size<alt_bn128_G1>(n); }
size_t ref_exp_window_size_g2(size_t n) { ref_init(); return get_exp_window_size<alt_bn128_G2>(n); }

int ref_batch_to_special_g1(int curve, uint64_t *pts, size_t n)
{
    ref_init();
    if (curve == 0) return to_special_impl<alt_bn128_G1>(pts, n);
#ifdef REF_WITH_BN128
    if (curve == 1) return to_special_impl<bn128_G1>(pts, n);
#endif
    return 2;
}

int ref_batch_to_special_g2(int curve, uint64_t *pts, size_t n)
{
    ref_init();
    if (curve == 0) return to_special_impl<alt_bn128_G2>(pts, n);
#ifdef REF_WITH_BN128
    if (curve == 1) return to_special_impl<bn128_G2>(pts, n);
#endif
    return 2;
}

int ref_g1_op(int curve, int op, const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out)
{
    ref_init();
    if (curve == 0) return group_op_impl<alt_bn128_G1>(op, a, b, n, out);
#ifdef REF_WITH_BN128
    if (curve == 1) return group_op_impl<bn128_G1>(op, a, b, n, out);
#endif
    return 2;
}

int ref_g2_op(int curve, int op, const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out)
{
    ref_init();
    if (curve == 0) return group_op_impl<alt_bn128_G2>(op, a, b, n, out);
#ifdef REF_WITH_BN128
    if (curve == 1) return group_op_impl<bn128_G2>(op, a, b, n, out);
#endif
    return 2;
}

// out[i] = scalars[i] * base (stride_base=0) or scalars[i] * base[i] (stride_base=1)
int ref_scalar_mul_g1(const uint64_t *base, const uint64_t *scalars, size_t n, int stride_base,
                      int normalise, uint64_t *out)
{
    ref_init();
    return scalar_mul_impl<alt_bn128_G1, alt_bn128_Fr>(base, scalars, n, stride_base, normalise, out);
}

int ref_scalar_mul_g2(const uint64_t *base, const uint64_t *scalars, size_t n, int stride_base,
                      int normalise, uint64_t *out)
{
    ref_init();
    return scalar_mul_impl<alt_bn128_G2, alt_bn128_Fr>(base, scalars, n, stride_base, normalise, out);
}

int ref_fq_op(int op, const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out)
{
    ref_init();
    return field_op_impl<alt_bn128_Fq>(op, a, b, n, out, 4);
}
int ref_fr_op(int op, const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out)
{
    ref_init();
    return field_op_impl<alt_bn128_Fr>(op, a, b, n, out, 4);
}
int ref_fq2_op(int op, const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out)
{
    ref_init();
    return field_op_impl<alt_bn128_Fq2>(op, a, b, n, out, 8);
}

// Montgomery <-> standard representation (Fp_model(bigint) ctor / as_bigint()).
int ref_fr_from_bigint(const uint64_t *a, size_t n, uint64_t *out)
{
    ref_init();
    for (size_t i = 0; i < n; i++) {
        bigint<4> b;
        memcpy(b.data, a + 4 * i, 32);
        alt_bn128_Fr f(b);
        store(out + 4 * i, f);
    }
    return 0;
}
int ref_fr_as_bigint(const uint64_t *a, size_t n, uint64_t *out)
{
    ref_init();
    for (size_t i = 0; i < n; i++) {
        alt_bn128_Fr f;
        load(f, a + 4 * i);
        bigint<4> b = f.as_bigint();
        memcpy(out + 4 * i, b.data, 32);
    }
    return 0;
}
int ref_fq_from_bigint(const uint64_t *a, size_t n, uint64_t *out)
{
    ref_init();
    for (size_t i = 0; i < n; i++) {
        bigint<4> b;
        memcpy(b.data, a + 4 * i, 32);
        alt_bn128_Fq f(b);
        store(out + 4 * i, f);
    }
    return 0;
}

// scalars[i] = SHA512_rng<Fr>(idx0 + i), Montgomery form (LFF/common/rng.tcc:26-72)
int ref_sha512_rng_fr(uint64_t idx0, size_t n, uint64_t *out)
{
    ref_init();
    for (size_t i = 0; i < n; i++) {
        alt_bn128_Fr f = SHA512_rng<alt_bn128_Fr>(idx0 + i);
        store(out + 4 * i, f);
    }
    return 0;
}

int ref_g1_one(uint64_t *out) { ref_init(); store(out, alt_bn128_G1::one()); return 0; }
int ref_g2_one(uint64_t *out) { ref_init(); store(out, alt_bn128_G2::one()); return 0; }
int ref_g1_zero(int curve, uint64_t *out)
{
    ref_init();
    if (curve == 0) { store(out, alt_bn128_G1::zero()); return 0; }
#ifdef REF_WITH_BN128
    if (curve == 1) { store(out, bn128_G1::zero()); return 0; }
#endif
    return 2;
}
int ref_g2_zero(int curve, uint64_t *out)
{
    ref_init();
    if (curve == 0) { store(out, alt_bn128_G2::zero()); return 0; }
#ifdef REF_WITH_BN128
    if (curve == 1) { store(out, bn128_G2::zero()); return 0; }
#endif
    return 2;
}

} // extern "C"

// ---- wire format: the reference's own operator<< / operator>> with point compression,
// through a string stream.  With -DBINARY_OUTPUT (how this library is built) a point is
//   '0'|'1' (is_zero)  |  X raw bytes (32 for G1, 64 for G2)  |  '0'|'1' (Y bit)
// alt_bn128: X = as_bigint (no -DMONTGOMERY_OUTPUT here) -> flavour 0 of include/b200_msm.h; bn128: flavour 2.
template <typename G>
int compress_impl(const uint64_t *pts, size_t n, uint64_t *x_out, uint8_t *flags)
{
    const size_t L = limbs_of<G>(), xb = L / 3 * 8;
    for (size_t i = 0; i < n; i++) {
        G g;
        load(g, pts + i * L);
        std::ostringstream os;
        os << g;
        const std::string s = os.str();
        if (s.size() != xb + 2) return 3;
        memcpy((char *)x_out + i * xb, s.data() + 1, xb);
        flags[i] = (uint8_t)((s[0] == '1' ? 2 : 0) | (s[xb + 1] == '1' ? 1 : 0));
    }
    return 0;
}
template <typename G>
int decompress_impl(const uint64_t *x, const uint8_t *flags, size_t n, uint64_t *pts_out)
{
    const size_t L = limbs_of<G>(), xb = L / 3 * 8;
    for (size_t i = 0; i < n; i++) {
        std::string s(xb + 2, '0');
        s[0] = (flags[i] & 2) ? '1' : '0';
        memcpy(&s[1], (const char *)x + i * xb, xb);
        s[xb + 1] = (flags[i] & 1) ? '1' : '0';
        std::istringstream is(s);
        G g;
        is >> g;
        store(pts_out + i * L, g);
    }
    return 0;
}

extern "C" {
int ref_compress_g1(int curve, const uint64_t *pts, size_t n, uint64_t *x_out, uint8_t *flags)
{
    ref_init();
    if (curve == 0) return compress_impl<alt_bn128_G1>(pts, n, x_out, flags);
#ifdef REF_WITH_BN128
    if (curve == 1) return compress_impl<bn128_G1>(pts, n, x_out, flags);
#endif
    return 2;
}
int ref_compress_g2(int curve, const uint64_t *pts, size_t n, uint64_t *x_out, uint8_t *flags)
{
    ref_init();
    if (curve == 0) return compress_impl<alt_bn128_G2>(pts, n, x_out, flags);
#ifdef REF_WITH_BN128
    if (curve == 1) return compress_impl<bn128_G2>(pts, n, x_out, flags);
#endif
    return 2;
}
int ref_decompress_g1(int curve, const uint64_t *x, const uint8_t *flags, size_t n, uint64_t *pts_out)
{
    ref_init();
    if (curve == 0) return decompress_impl<alt_bn128_G1>(x, flags, n, pts_out);
#ifdef REF_WITH_BN128
    if (curve == 1) return decompress_impl<bn128_G1>(x, flags, n, pts_out);
#endif
    return 2;
}
int ref_decompress_g2(int curve, const uint64_t *x, const uint8_t *flags, size_t n, uint64_t *pts_out)
{
    ref_init();
    if (curve == 0) return decompress_impl<alt_bn128_G2>(x, flags, n, pts_out);
#ifdef REF_WITH_BN128
    if (curve == 1) return decompress_impl<bn128_G2>(x, flags, n, pts_out);
#endif
    return 2;
}
}  // extern "C"
