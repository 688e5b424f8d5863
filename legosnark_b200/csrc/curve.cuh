// curve.cuh — BN254 G1 / G2 point arithmetic for the MSM engine, templated on the
// coordinate field F (Fq for G1, Fq2 for G2).  Curve y^2 = x^3 + b with a = 0
// (alt_bn128_init.cpp:134-136), so no formula below needs b.
//
// The reference works in Jacobian coordinates (alt_bn128_g1.cpp: operator+
// :139-195 add-2007-bl, mixed_add :256-323 madd-2007-bl, dbl :325-358
// dbl-2009-l; ate-pairing's ECAdd/ECDouble, ATE/include/bn.h:2496-2577).  Group
// elements are only ever compared after affine normalisation, so
// the engine is free to accumulate in extended Jacobian "XYZZ" coordinates
// (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2), whose mixed addition needs no Z^3:
//   affine + XYZZ : 8M + 2S = 10 modmul  (EFD madd-2008-s)   vs 11 for madd-2007-bl
//   XYZZ  + XYZZ  : 12M + 2S = 14 modmul (EFD add-2008-s)    vs 16 for add-2007-bl
//   2 * XYZZ      : 6M + 4S (a = 0: 6M + 3S = 9)  (EFD dbl-2008-s-1)
// Every exceptional case the reference handles is handled here as well: infinity
// on either side, P == Q (falls through to doubling) and P == -Q (-> infinity).
#pragma once
#include "field.cuh"

namespace b200 {

// Affine point; (0,0) encodes infinity (never on the curve since b != 0).
template <class F>
struct Affine {
    F x, y;
    B200_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
    B200_HD static Affine inf() { return Affine{F::zero(), F::zero()}; }
};

// Jacobian point in the reference's memory layout X|Y|Z; Z == 0 <=> infinity.
template <class F>
struct Jacobian {
    F x, y, z;
    B200_HD bool is_inf() const { return z.is_zero(); }
    // alt_bn128 G1_zero/G2_zero = (0,1,0) (alt_bn128_init.cpp:145-147, 205-207)
    B200_HD static Jacobian inf() { return Jacobian{F::zero(), F::one(), F::zero()}; }
};

template <class F>
struct XYZZ {
    F x, y, zz, zzz;

    B200_HD bool is_inf() const { return zz.is_zero(); }
    B200_HD static XYZZ inf() { return XYZZ{F::zero(), F::zero(), F::zero(), F::zero()}; }
    B200_HD static XYZZ from_affine(const Affine<F> &p)
    {
        if (p.is_inf()) return inf();
        return XYZZ{p.x, p.y, F::one(), F::one()};
    }
    // Jacobian (X*ZZ, Y*ZZZ, ZZ): Z = ZZ gives Z^2 = ZZ^2, Z^3 = ZZ^3 = ZZZ^2.
    B200_HD Jacobian<F> to_jacobian() const
    {
        if (is_inf()) return Jacobian<F>::inf();
        return Jacobian<F>{F::mul(x, zz), F::mul(y, zzz), zz};
    }
    B200_HD static XYZZ from_jacobian(const Jacobian<F> &p)
    {
        if (p.is_inf()) return inf();
        const F zz = F::sqr(p.z);
        return XYZZ{p.x, p.y, zz, F::mul(zz, p.z)};
    }
};

// 2 * (x, y) for an affine point: dbl-2008-s-1 with ZZ = ZZZ = 1 (a = 0).
template <class F>
B200_COLD XYZZ<F> xyzz_dbl_affine(const F &x, const F &y)
{
    const F U = F::dbl(y);
    const F V = F::sqr(U);
    const F W = F::mul(U, V);
    const F S = F::mul(x, V);
    const F xx = F::sqr(x);
    const F M = F::add(F::dbl(xx), xx);
    XYZZ<F> r;
    r.x = F::sub(F::sqr(M), F::dbl(S));
    r.y = F::mul_sub(M, F::sub(S, r.x), W, y);
    r.zz = V;
    r.zzz = W;
    return r;  // y == 0 would give zz == 0 == infinity (no such point in the prime-order groups)
}

// 2 * P, dbl-2008-s-1 (a = 0).
template <class F>
B200_HD XYZZ<F> xyzz_dbl(const XYZZ<F> &p)
{
    if (p.is_inf()) return p;
    const F U = F::dbl(p.y);
    const F V = F::sqr(U);
    const F W = F::mul(U, V);
    const F S = F::mul(p.x, V);
    const F xx = F::sqr(p.x);
    const F M = F::add(F::dbl(xx), xx);
    XYZZ<F> r;
    r.x = F::sub(F::sqr(M), F::dbl(S));
    r.y = F::mul_sub(M, F::sub(S, r.x), W, p.y);
    r.zz = F::mul(V, p.zz);
    r.zzz = F::mul(W, p.zzz);
    return r;
}

// acc += (x2, +-y2): mixed addition madd-2008-s; (x2, y2) must not be infinity.
template <class F>
B200_HD void xyzz_madd(XYZZ<F> &acc, const F &x2, const F &y2_in, bool negate)
{
    const F y2 = F::cneg(y2_in, negate);
    if (acc.is_inf()) {
        acc.x = x2;
        acc.y = y2;
        acc.zz = F::one();
        acc.zzz = F::one();
        return;
    }
    const F U2 = F::mul(x2, acc.zz);
    const F S2 = F::mul(y2, acc.zzz);
    const F P = F::sub(U2, acc.x);
    const F R = F::sub(S2, acc.y);
    if (P.is_zero()) {
        // same x: either the same point (double it) or its negative (-> infinity)
        if (R.is_zero()) acc = xyzz_dbl_affine(x2, y2);
        else acc = XYZZ<F>::inf();
        return;
    }
    const F PP = F::sqr(P);
    const F PPP = F::mul(P, PP);
    const F Q = F::mul(acc.x, PP);
    const F X3 = F::sub(F::sub(F::sqr(R), PPP), F::dbl(Q));
    acc.y = F::mul_sub(R, F::sub(Q, X3), acc.y, PPP);
    acc.x = X3;
    acc.zz = F::mul(acc.zz, PP);
    acc.zzz = F::mul(acc.zzz, PPP);
}

// The same mixed addition with its independent products issued in pairs (Fp::mul2: alternating rows of two Montgomery products),
// for base fields that have mul2 (Fq).  Same products, same result; option `g1_paired` of k_accumulate.
template <class F>
B200_HD void xyzz_madd_paired(XYZZ<F> &acc, const F &x2, const F &y2_in, bool negate)
{
    const F y2 = F::cneg(y2_in, negate);
    if (acc.is_inf()) {
        acc.x = x2;
        acc.y = y2;
        acc.zz = F::one();
        acc.zzz = F::one();
        return;
    }
    F U2, S2;
    F::mul2(U2, S2, x2, acc.zz, y2, acc.zzz);
    const F P = F::sub(U2, acc.x);
    const F R = F::sub(S2, acc.y);
    if (P.is_zero()) {
        if (R.is_zero()) acc = xyzz_dbl_affine(x2, y2);
        else acc = XYZZ<F>::inf();
        return;
    }
    const F PP = F::sqr(P);
    F PPP, Q;
    F::mul2(PPP, Q, P, PP, acc.x, PP);
    const F X3 = F::sub(F::sub(F::sqr(R), PPP), F::dbl(Q));
    acc.y = F::mul_sub(R, F::sub(Q, X3), acc.y, PPP);
    acc.x = X3;
    F::mul2(acc.zz, acc.zzz, acc.zz, PP, acc.zzz, PPP);
}

template <class F>
B200_COLD void xyzz_dbl_cold(XYZZ<F> *p);

// acc += q, add-2008-s.
template <class F>
B200_HD void xyzz_add(XYZZ<F> &acc, const XYZZ<F> &q)
{
    if (q.is_inf()) return;
    if (acc.is_inf()) {
        acc = q;
        return;
    }
    const F U1 = F::mul(acc.x, q.zz);
    const F U2 = F::mul(q.x, acc.zz);
    const F S1 = F::mul(acc.y, q.zzz);
    const F S2 = F::mul(q.y, acc.zzz);
    const F P = F::sub(U2, U1);
    const F R = F::sub(S2, S1);
    if (P.is_zero()) {
        if (R.is_zero()) xyzz_dbl_cold(&acc);
        else acc = XYZZ<F>::inf();
        return;
    }
    const F PP = F::sqr(P);
    const F PPP = F::mul(P, PP);
    const F Q = F::mul(U1, PP);
    const F X3 = F::sub(F::sub(F::sqr(R), PPP), F::dbl(Q));
    acc.y = F::mul_sub(R, F::sub(Q, X3), S1, PPP);
    acc.x = X3;
    acc.zz = F::mul(F::mul(acc.zz, q.zz), PP);
    acc.zzz = F::mul(F::mul(acc.zzz, q.zzz), PPP);
}

// ---- out-of-line variants for the cold kernels (window reduction, bucket combine,
// table build, parity hooks): one copy of each adder per kernel instead of one per
// call site keeps code size and compile time down; the hot accumulation loop uses
// the inlined xyzz_madd above. ----
template <class F>
B200_COLD void xyzz_add_cold(XYZZ<F> *acc, const XYZZ<F> *q)
{
    xyzz_add(*acc, *q);
}
template <class F>
B200_COLD void xyzz_dbl_cold(XYZZ<F> *p)
{
    *p = xyzz_dbl(*p);
}
template <class F>
B200_COLD void xyzz_madd_cold(XYZZ<F> *acc, const Affine<F> *p, bool negate)
{
    xyzz_madd(*acc, p->x, p->y, negate);
}

// k * p for a small unsigned k (MSB-first double-and-add), used to weight
// bucket-segment sums in the window reduction.
template <class F>
B200_HD XYZZ<F> xyzz_mul_small(const XYZZ<F> &p, uint32_t k)
{
    XYZZ<F> r = XYZZ<F>::inf();
    if (k == 0 || p.is_inf()) return r;
    int top = 31;
    while (!((k >> top) & 1u)) top--;
    r = p;
#pragma unroll 1
    for (int i = top - 1; i >= 0; i--) {
        xyzz_dbl_cold(&r);
        if ((k >> i) & 1u) xyzz_add_cold(&r, &p);
    }
    return r;
}

}  // namespace b200
