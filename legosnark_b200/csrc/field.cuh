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


// Operand-scanning step for a SUM OF TWO PRODUCTS: T += a * bi + c * di; T += mi * p; T /= 2^32.
// One reduction row serves both products (25 multiply-adds per step instead of 34), which is
// what the point formulas' "R (Q - X3) - Y1 PPP" shape wants.  Bound: a, c < p and
// bi, di, mi < 2^32 give T < 3p 2^32 + 3p < 2^288 inside the step, so the 9-column window
// (even: columns 0..7, odd: columns 1..8) never overflows.
template <class P>
B200_HD void mont_step2(uint32_t even[8], uint32_t odd[8], const uint32_t a[8], uint32_t bi, const uint32_t c[8], uint32_t di,
                        bool first)
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
        even[0] = add_cc(even[0], odd[1]);
#pragma unroll
        for (int j = 0; j < 6; j += 2) wmadc_cc(odd[j], odd[j + 1], a[j + 1], bi, odd[j + 2], odd[j + 3]);
        wmadc(odd[6], odd[7], a[7], bi, 0u, 0u);
        wide_mad_row(even, [&](int j) { return a[j]; }, bi);
        odd[7] = addc(odd[7], 0u);
    }
    wide_mad_row(odd, [&](int j) { return c[j + 1]; }, di);  // odd <= T / 2^32 < 2^256: no carry-out
    wide_mad_row(even, [&](int j) { return c[j]; }, di);
    odd[7] = addc(odd[7], 0u);
    const uint32_t mi = even[0] * P::INV;
    wide_mad_row(odd, [&](int j) { return P::mod(j + 1); }, mi);
    wide_mad_row(even, [&](int j) { return P::mod(j); }, mi);
    odd[7] = addc(odd[7], 0u);
}

// t[0..15] = a^2 as a plain 512-bit integer: the 28 products a_i a_j (i < j) once, doubled by
// a one-bit shift, plus the 8 squares a_i^2 = 36 wide multiply-adds instead of 64
// (fp.tcc:593-639 `squared` has the same structure on 64-bit limbs).  Products that start on
// even columns go to E, those on odd columns to O (limb k of O sits on column k + 1), so every
// row is an uninterrupted IMAD.WIDE carry chain like in mont_step.  Each chain ends on the top
// limb pair written so far (or opens a fresh pair), so its carry-out lands in a limb that
// holds nothing else yet.
B200_HD void sqr_wide(uint32_t t[16], const uint32_t a[8])
{
    uint32_t E[16], O[16];
#pragma unroll
    for (int k = 0; k < 16; k++) E[k] = O[k] = 0u;
#pragma unroll
    for (int i = 0; i < 7; i++) {
        // odd columns i + j: j = i + 1, i + 3, ...
        {
            int last = -1;
#pragma unroll
            for (int j = i + 1; j < 8; j += 2) {
                const int k = i + j - 1;
                if (last < 0) wmad_cc(O[k], O[k + 1], a[i], a[j], O[k], O[k + 1]);
                else wmadc_cc(O[k], O[k + 1], a[i], a[j], O[k], O[k + 1]);
                last = k;
            }
            if (last >= 0 && last + 2 < 14) O[last + 2] = addc(O[last + 2], 0u);
        }
        // even columns: j = i + 2, i + 4, ...
        {
            int last = -1;
#pragma unroll
            for (int j = i + 2; j < 8; j += 2) {
                const int k = i + j;
                if (last < 0) wmad_cc(E[k], E[k + 1], a[i], a[j], E[k], E[k + 1]);
                else wmadc_cc(E[k], E[k + 1], a[i], a[j], E[k], E[k + 1]);
                last = k;
            }
            if (last >= 0 && last + 2 < 14) E[last + 2] = addc(E[last + 2], 0u);
        }
    }
    // D = E + O * 2^32 (off-diagonal sum, < 2^479); E uses limbs 2..13, O limbs 0..13
    uint32_t D[16];
    D[0] = 0u;
    D[1] = O[0];
    D[2] = add_cc(E[2], O[1]);
#pragma unroll
    for (int k = 3; k < 14; k++) D[k] = addc_cc(E[k], O[k - 1]);
    D[14] = addc(O[13], 0u);
    // t = 2 D + sum a_i^2 2^(64 i)
#pragma unroll
    for (int k = 15; k > 0; k--) t[k] = (k == 15 ? 0u : (D[k] << 1)) | (D[k - 1] >> 31);
    t[0] = 0u;
    wmad_cc(t[0], t[1], a[0], a[0], t[0], t[1]);
#pragma unroll
    for (int i = 1; i < 7; i++) wmadc_cc(t[2 * i], t[2 * i + 1], a[i], a[i], t[2 * i], t[2 * i + 1]);
    wmadc(t[14], t[15], a[7], a[7], t[14], t[15]);
}

// One row of the stand-alone Montgomery reduction: the 9-column window slides down one column
// (`odd` is last row's even accumulator, as in mont_step) and mi * p is added.  The shift of the
// old even limbs rides on the addends of the p_odd chain.
template <class P>
B200_HD void redc_row(uint32_t even[8], uint32_t odd[8])
{
    const uint32_t mi = (even[0] + odd[1]) * P::INV;
    even[0] = add_cc(even[0], odd[1]);
#pragma unroll
    for (int j = 0; j < 6; j += 2) wmadc_cc(odd[j], odd[j + 1], P::mod(j + 1), mi, odd[j + 2], odd[j + 3]);
    wmadc(odd[6], odd[7], P::mod(7), mi, 0u, 0u);
    wide_mad_row(even, [&](int j) { return P::mod(j); }, mi);
    odd[7] = addc(odd[7], 0u);
}

// r = t / 2^256 mod p for t < p 2^256 (16 limbs), fully reduced: (t_lo + M p) / 2^256 <= p by the
// eight rows above (8 x (8 wide + 1) = 72 multiply-adds), plus t_hi.
template <class P>
B200_HD void redc_wide(uint32_t r[8], const uint32_t t[16])
{
    uint32_t even[8], odd[8];
#pragma unroll
    for (int k = 0; k < 8; k++) even[k] = t[k];
    {
        const uint32_t mi = even[0] * P::INV;
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            odd[j] = P::mod(j + 1) * mi;
            odd[j + 1] = mul_hi(P::mod(j + 1), mi);
        }
        wide_mad_row(even, [&](int j) { return P::mod(j); }, mi);
        odd[7] = addc(odd[7], 0u);
    }
#pragma unroll
    for (int i = 1; i < 8; i += 2) {
        redc_row<P>(odd, even);
        if (i + 1 < 8) redc_row<P>(even, odd);
    }
    // the last row ran with the roles swapped, like the last mont_step of mul(): T / 2^32 = even[k] + odd[k + 1]
    r[0] = add_cc(even[0], odd[1]);
#pragma unroll
    for (int k = 1; k < 7; k++) r[k] = addc_cc(even[k], odd[k + 1]);
    r[7] = addc(even[7], 0u);
    r[0] = add_cc(r[0], t[8]);
#pragma unroll
    for (int k = 1; k < 7; k++) r[k] = addc_cc(r[k], t[8 + k]);
    r[7] = addc(r[7], t[15]);
    final_sub<P>(r);
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
    // a^2 / R mod p with the dedicated 36-product square and a stand-alone reduction:
    // 108 multiply-adds instead of 136.
    B200_HD static Fp sqr(const Fp &a)
    {
        uint32_t t[16];
        detail::sqr_wide(t, a.l);
        Fp r;
        detail::redc_wide<P>(r.l, t);
        return r;
    }
    // (a b + c d) / R mod p with ONE interleaved reduction: 200 multiply-adds instead of 272.
    // The result is below p (2p / R + 1) < 1.4 p before the final subtraction.
    B200_HD static Fp mul_add(const Fp &a, const Fp &b, const Fp &c, const Fp &d)
    {
        uint32_t even[8], odd[8];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            detail::mont_step2<P>(even, odd, a.l, b.l[i], c.l, d.l[i], i == 0);
            detail::mont_step2<P>(odd, even, a.l, b.l[i + 1], c.l, d.l[i + 1], false);
        }
        Fp r;
        r.l[0] = add_cc(even[0], odd[1]);
#pragma unroll
        for (int k = 1; k < 7; k++) r.l[k] = addc_cc(even[k], odd[k + 1]);
        r.l[7] = addc(even[7], 0u);
        detail::final_sub<P>(r.l);
        return r;
    }
    // a b - c d
    B200_HD static Fp mul_sub(const Fp &a, const Fp &b, const Fp &c, const Fp &d) { return mul_add(a, b, neg(c), d); }

    // Two INDEPENDENT products r0 = a0 b0, r1 = a1 b1 with their operand-scanning rows alternating in the instruction stream.
    // A product is a chain of dependent carry rows (row i + 1 needs row i's even[0] for the quotient digit); a warp that runs one
    // product at a time stalls on that chain whenever its scheduler has no other warp to issue (k_accumulate<Fq2>: two warps per
    // scheduler, `wait` is the top stall).  With the rows of two products adjacent ptxas overlaps the tail of one with the head of
    // the other.  Same instructions, same results.
    B200_HD static void mul2(Fp &r0, Fp &r1, const Fp &a0, const Fp &b0, const Fp &a1, const Fp &b1)
    {
        uint32_t e0[8], o0[8], e1[8], o1[8];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            detail::mont_step<P>(e0, o0, a0.l, b0.l[i], i == 0);
            detail::mont_step<P>(e1, o1, a1.l, b1.l[i], i == 0);
            detail::mont_step<P>(o0, e0, a0.l, b0.l[i + 1], false);
            detail::mont_step<P>(o1, e1, a1.l, b1.l[i + 1], false);
        }
        r0.l[0] = add_cc(e0[0], o0[1]);
#pragma unroll
        for (int k = 1; k < 7; k++) r0.l[k] = addc_cc(e0[k], o0[k + 1]);
        r0.l[7] = addc(e0[7], 0u);
        r1.l[0] = add_cc(e1[0], o1[1]);
#pragma unroll
        for (int k = 1; k < 7; k++) r1.l[k] = addc_cc(e1[k], o1[k + 1]);
        r1.l[7] = addc(e1[7], 0u);
        detail::final_sub<P>(r0.l);
        detail::final_sub<P>(r1.l);
    }
    // r0 = a0 b0 + c0 d0, r1 = a1 b1 + c1 d1: two fused two-product passes (mul_add) with alternating rows
    B200_HD static void mul_add2(Fp &r0, Fp &r1, const Fp &a0, const Fp &b0, const Fp &c0, const Fp &d0, const Fp &a1, const Fp &b1,
                                 const Fp &c1, const Fp &d1)
    {
        uint32_t e0[8], o0[8], e1[8], o1[8];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            detail::mont_step2<P>(e0, o0, a0.l, b0.l[i], c0.l, d0.l[i], i == 0);
            detail::mont_step2<P>(e1, o1, a1.l, b1.l[i], c1.l, d1.l[i], i == 0);
            detail::mont_step2<P>(o0, e0, a0.l, b0.l[i + 1], c0.l, d0.l[i + 1], false);
            detail::mont_step2<P>(o1, e1, a1.l, b1.l[i + 1], c1.l, d1.l[i + 1], false);
        }
        r0.l[0] = add_cc(e0[0], o0[1]);
#pragma unroll
        for (int k = 1; k < 7; k++) r0.l[k] = addc_cc(e0[k], o0[k + 1]);
        r0.l[7] = addc(e0[7], 0u);
        r1.l[0] = add_cc(e1[0], o1[1]);
#pragma unroll
        for (int k = 1; k < 7; k++) r1.l[k] = addc_cc(e1[k], o1[k + 1]);
        r1.l[7] = addc(e1[7], 0u);
        detail::final_sub<P>(r0.l);
        detail::final_sub<P>(r1.l);
    }

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
    // fp2.tcc:72-84
    B200_HD static Fq2 mul(const Fq2 &x, const Fq2 &y)
    {
        // same value as the three-product Karatsuba form; each coordinate is one fused
        // two-product Montgomery pass (2 x 200 multiply-adds instead of 3 x 136)
#if defined(B200_FQ2_INTERLEAVED_PRODUCTS)
        // the two coordinates are independent: their rows alternate (Fq::mul_add2).  Measured on k_accumulate<Fq2> at 2^20
        // (profiles/r4g_stage.jsonl): 6.45 ms against 6.37 ms for the two passes one after the other -> off.
        Fq2 r;
        Fq::mul_add2(r.c0, r.c1, x.c0, y.c0, Fq::neg(x.c1), y.c1, x.c0, y.c1, x.c1, y.c0);
        return r;
#else
        return Fq2{Fq::mul_sub(x.c0, y.c0, x.c1, y.c1), Fq::mul_add(x.c0, y.c1, x.c1, y.c0)};
#endif
    }
    // complex squaring, fp2.tcc:111-120
    B200_HD static Fq2 sqr(const Fq2 &x)
    {
#if defined(B200_FQ2_INTERLEAVED_PRODUCTS)
        Fq ab, c0;
        Fq::mul2(ab, c0, x.c0, x.c1, Fq::add(x.c0, x.c1), Fq::sub(x.c0, x.c1));
        return Fq2{c0, Fq::dbl(ab)};
#else
        const Fq ab = Fq::mul(x.c0, x.c1);
        const Fq c0 = Fq::mul(Fq::add(x.c0, x.c1), Fq::sub(x.c0, x.c1));
        return Fq2{c0, Fq::dbl(ab)};
#endif
    }
    B200_HD static Fq2 mul_sub(const Fq2 &a, const Fq2 &b, const Fq2 &c, const Fq2 &d) { return sub(mul(a, b), mul(c, d)); }
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
