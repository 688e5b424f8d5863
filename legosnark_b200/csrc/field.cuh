// field.cuh — BN254 prime fields on 8 x 32-bit register-resident limbs.
//
// Replaces, on the device, libff's Fp_model<4, p> over GMP mpn limbs
// (LFF/algebra/fields/fp.tcc: mul_reduce :161-186, operator+= :309-420,
// operator-= :422-510, squared :593-639, as_bigint :227-238) and ate-pairing's
// JIT'd mie::Fp (ATE/src/zm2.cpp:1282-, R = 2^256 :3627).  Same Montgomery
// representation (R = 2^256, little-endian), so the 32-byte host images are
// used as-is; every result is fully reduced to [0, p) like the reference, which
// makes outputs bit-comparable.
//
// Multiplication: operand-scanning Montgomery product with the running sum
// split into two 8-limb accumulators, one for products that start on even
// columns and one for those that start on odd columns, so that every
// 32x32->64 product is added with one IMAD.WIDE.U32 (mad.lo.cc + madc.hi.cc
// pair) on an uninterrupted carry chain.  After each reduction step the value
// is divided by 2^32, which swaps the roles of the two accumulators.
// Cost: 8 x (8 + 8) wide multiply-adds + 8 IMAD for the quotient digits = 136
// multiply-add instructions.
#pragma once
#include "ptx_ops.cuh"

namespace b200 {

// ---- field constants (alt_bn128_init.cpp:40-53, 66-79; bn128_init.cpp:38-76) ----
struct FqParams {
    // q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
    static constexpr uint32_t INV = 0xe4866389u;  // -q^{-1} mod 2^32
    B200_HD static constexpr uint32_t mod(int i)
    {
        constexpr uint32_t m[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u,
                                   0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return m[i];
    }
    B200_HD static constexpr uint32_t one(int i)  // R mod q
    {
        constexpr uint32_t m[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u,
                                   0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return m[i];
    }
    B200_HD static constexpr uint32_t r2(int i)  // R^2 mod q
    {
        constexpr uint32_t m[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u,
                                   0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
        return m[i];
    }
};

struct FrParams {
    // r = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    static constexpr uint32_t INV = 0xefffffffu;
    B200_HD static constexpr uint32_t mod(int i)
    {
        constexpr uint32_t m[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                                   0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return m[i];
    }
    B200_HD static constexpr uint32_t one(int i)
    {
        constexpr uint32_t m[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u,
                                   0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return m[i];
    }
    B200_HD static constexpr uint32_t r2(int i)
    {
        constexpr uint32_t m[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u,
                                   0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
        return m[i];
    }
};

namespace detail {

// acc[0..7] (pairs starting at limb 0,2,4,6) += x[xoff], x[xoff+2], ... times m; one carry chain.
// Returns nothing: the caller picks up the carry-out with addc().
template <class GetX>
B200_HD void wide_mad_row(uint32_t acc[8], GetX x, uint32_t m)
{
    wmad_cc(acc[0], acc[1], x(0), m, acc[0], acc[1]);
#pragma unroll
    for (int j = 2; j < 8; j += 2) wmadc_cc(acc[j], acc[j + 1], x(j), m, acc[j], acc[j + 1]);
}

// One operand-scanning step: T += a * bi; T += mi * p; T /= 2^32 (implicit: the
// caller swaps `even` and `odd` for the next step).
//   T = sum even[k] 2^(32k) + sum odd[k] 2^(32(k+1))
template <class P>
B200_HD void mont_step(uint32_t even[8], uint32_t odd[8], const uint32_t a[8], uint32_t bi, bool first)
{
    if (first) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            even[j] = a[j] * bi;
            even[j + 1] = mul_hi(a[j], bi);
            odd[j] = a[j + 1] * bi;
            odd[j + 1] = mul_hi(a[j + 1], bi);
        }
    } else {
        // `odd` is last step's even accumulator: its limb 0 is zero, limb 1 sits on
        // the new column 0, limbs 2..7 on the new columns 1..6.
        even[0] = add_cc(even[0], odd[1]);
#pragma unroll
        for (int j = 0; j < 6; j += 2) wmadc_cc(odd[j], odd[j + 1], a[j + 1], bi, odd[j + 2], odd[j + 3]);
        wmadc(odd[6], odd[7], a[7], bi, 0u, 0u);
        wide_mad_row(even, [&](int j) { return a[j]; }, bi);
        odd[7] = addc(odd[7], 0u);
    }
    const uint32_t mi = even[0] * P::INV;
    wide_mad_row(odd, [&](int j) { return P::mod(j + 1); }, mi);  // no carry-out: T < 2^287
    wide_mad_row(even, [&](int j) { return P::mod(j); }, mi);
    odd[7] = addc(odd[7], 0u);
    // now even[0] == 0
}

// r = (r >= p) ? r - p : r
template <class P>
B200_HD void final_sub(uint32_t r[8])
{
    uint32_t t[8];
    t[0] = sub_cc(r[0], P::mod(0));
#pragma unroll
    for (int i = 1; i < 8; i++) t[i] = subc_cc(r[i], P::mod(i));
    const uint32_t borrow = subc(0u, 0u);  // 0xffffffff if r < p
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = borrow ? r[i] : t[i];
}

}  // namespace detail

template <class P>
struct alignas(16) Fp {
    uint32_t l[8];

    B200_HD static Fp zero()
    {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = 0;
        return r;
    }
    B200_HD static Fp one()
    {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = P::one(i);
        return r;
    }
    B200_HD static Fp r2()
    {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = P::r2(i);
        return r;
    }
    B200_HD bool is_zero() const
    {
        uint32_t o = l[0];
#pragma unroll
        for (int i = 1; i < 8; i++) o |= l[i];
        return o == 0;
    }
    B200_HD bool operator==(const Fp &b) const
    {
        uint32_t o = l[0] ^ b.l[0];
#pragma unroll
        for (int i = 1; i < 8; i++) o |= l[i] ^ b.l[i];
        return o == 0;
    }
    B200_HD bool operator!=(const Fp &b) const { return !(*this == b); }

    // Montgomery product a*b/R mod p, fully reduced.
    B200_HD static Fp mul(const Fp &a, const Fp &b)
    {
        uint32_t even[8], odd[8];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            detail::mont_step<P>(even, odd, a.l, b.l[i], i == 0);
            detail::mont_step<P>(odd, even, a.l, b.l[i + 1], false);
        }
        // T/2^32: result[k] = even[k] + odd[k+1]
        Fp r;
        r.l[0] = add_cc(even[0], odd[1]);
#pragma unroll
        for (int k = 1; k < 7; k++) r.l[k] = addc_cc(even[k], odd[k + 1]);
        r.l[7] = addc(even[7], 0u);
        detail::final_sub<P>(r.l);
        return r;
    }
    B200_HD static Fp sqr(const Fp &a) { return mul(a, a); }

    // a/R mod p: Fp_model::as_bigint() (fp.tcc:227-238) — Montgomery product with the integer 1.
    B200_HD static Fp from_mont(const Fp &a)
    {
        Fp o = zero();
        o.l[0] = 1;
        return mul(a, o);
    }
    // a*R mod p: Fp_model(bigint) (fp.tcc:189-194)
    B200_HD static Fp to_mont(const Fp &a) { return mul(a, r2()); }

    B200_HD static Fp add(const Fp &a, const Fp &b)
    {
        Fp r;
        r.l[0] = add_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < 7; i++) r.l[i] = addc_cc(a.l[i], b.l[i]);
        r.l[7] = addc(a.l[7], b.l[7]);  // 2p < 2^255: no carry out
        detail::final_sub<P>(r.l);
        return r;
    }
    B200_HD static Fp dbl(const Fp &a) { return add(a, a); }

    B200_HD static Fp sub(const Fp &a, const Fp &b)
    {
        Fp r;
        r.l[0] = sub_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < 8; i++) r.l[i] = subc_cc(a.l[i], b.l[i]);
        const uint32_t borrow = subc(0u, 0u);  // all ones if a < b
        r.l[0] = add_cc(r.l[0], borrow & P::mod(0));
#pragma unroll
        for (int i = 1; i < 7; i++) r.l[i] = addc_cc(r.l[i], borrow & P::mod(i));
        r.l[7] = addc(r.l[7], borrow & P::mod(7));
        return r;
    }
    B200_HD static Fp neg(const Fp &a)
    {
        // zero stays zero (fp.tcc:551-567)
        const uint32_t nz = a.is_zero() ? 0u : 0xffffffffu;
        Fp r;
        r.l[0] = sub_cc(nz & P::mod(0), a.l[0]);
#pragma unroll
        for (int i = 1; i < 7; i++) r.l[i] = subc_cc(nz & P::mod(i), a.l[i]);
        r.l[7] = subc(nz & P::mod(7), a.l[7]);
        return r;
    }
    B200_HD static Fp cneg(const Fp &a, bool flag) { return flag ? neg(a) : a; }

    // a^(p-2): Fermat inversion (the reference runs mpn_gcdext, fp.tcc:641-685; the
    // inverse is unique so the limbs agree).  4-bit fixed window: 252 sqr + ~64 mul + 14.
    B200_HD static Fp inv(const Fp &a)
    {
        Fp tab[16];
        tab[0] = one();
        tab[1] = a;
#pragma unroll 1
        for (int i = 2; i < 16; i++) tab[i] = mul(tab[i - 1], a);
        // exponent e = p - 2, scanned 4 bits at a time from the top
        uint32_t e[8];
        e[0] = sub_cc(P::mod(0), 2u);
#pragma unroll
        for (int i = 1; i < 7; i++) e[i] = subc_cc(P::mod(i), 0u);
        e[7] = subc(P::mod(7), 0u);
        Fp r = one();
#pragma unroll 1
        for (int w = 63; w >= 0; w--) {
            r = sqr(r);
            r = sqr(r);
            r = sqr(r);
            r = sqr(r);
            const uint32_t d = (e[w >> 3] >> ((w & 7) * 4)) & 15u;
            if (d) r = mul(r, tab[d]);
        }
        return r;
    }
};

typedef Fp<FqParams> Fq;
typedef Fp<FrParams> Fr;

// ---- Fq2 = Fq[u]/(u^2 + 1): Fp2_model (LFF/algebra/fields/fp2.tcc), non_residue = -1
// (alt_bn128_init.cpp:95), layout c0,c1 (fp2.hpp:49) ----
struct Fq2 {
    Fq c0, c1;

    B200_HD static Fq2 zero() { return Fq2{Fq::zero(), Fq::zero()}; }
    B200_HD static Fq2 one() { return Fq2{Fq::one(), Fq::zero()}; }
    B200_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    B200_HD bool operator==(const Fq2 &b) const { return c0 == b.c0 && c1 == b.c1; }
    B200_HD bool operator!=(const Fq2 &b) const { return !(*this == b); }
    // Karatsuba, fp2.tcc:72-84
    B200_HD static Fq2 mul(const Fq2 &x, const Fq2 &y)
    {
        const Fq aA = Fq::mul(x.c0, y.c0);
        const Fq bB = Fq::mul(x.c1, y.c1);
        const Fq s = Fq::mul(Fq::add(x.c0, x.c1), Fq::add(y.c0, y.c1));
        return Fq2{Fq::sub(aA, bB), Fq::sub(Fq::sub(s, aA), bB)};
    }
    // complex squaring, fp2.tcc:111-120
    B200_HD static Fq2 sqr(const Fq2 &x)
    {
        const Fq ab = Fq::mul(x.c0, x.c1);
        const Fq c0 = Fq::mul(Fq::add(x.c0, x.c1), Fq::sub(x.c0, x.c1));
        return Fq2{c0, Fq::dbl(ab)};
    }
    B200_HD static Fq2 add(const Fq2 &a, const Fq2 &b) { return Fq2{Fq::add(a.c0, b.c0), Fq::add(a.c1, b.c1)}; }
    B200_HD static Fq2 dbl(const Fq2 &a) { return Fq2{Fq::dbl(a.c0), Fq::dbl(a.c1)}; }
    B200_HD static Fq2 sub(const Fq2 &a, const Fq2 &b) { return Fq2{Fq::sub(a.c0, b.c0), Fq::sub(a.c1, b.c1)}; }
    B200_HD static Fq2 neg(const Fq2 &a) { return Fq2{Fq::neg(a.c0), Fq::neg(a.c1)}; }
    B200_HD static Fq2 cneg(const Fq2 &a, bool flag) { return flag ? neg(a) : a; }
    // fp2.tcc:122-136
    B200_HD static Fq2 inv(const Fq2 &x)
    {
        const Fq t = Fq::inv(Fq::add(Fq::sqr(x.c0), Fq::sqr(x.c1)));
        return Fq2{Fq::mul(x.c0, t), Fq::neg(Fq::mul(x.c1, t))};
    }
};

}  // namespace b200
