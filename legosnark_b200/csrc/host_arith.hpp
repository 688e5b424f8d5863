// host_arith.hpp — host-side BN254 arithmetic for the O(254) tail of an MSM:
// Horner combination of the per-window sums the GPU returns
// (result = sum_k 2^(c k) W_k: 254 doublings + W additions), the sum of the
// per-GPU partials (north star: "the partials are summed on the host"), the
// final affine normalisation and the 2^(o w) g row bases of a fixed-base table.
// 4 x u64 Montgomery limbs (R = 2^256) — the reference's own memory image, so
// results are written straight into the caller's buffers.
//
// This is product code, independent of oracle/.  It never touches the n-sized
// inputs; the hot path (bucket accumulation and reduction) runs on the GPU only.
#pragma once
#include <cstdint>
#include <cstring>

namespace b200 {
namespace host {

typedef unsigned __int128 u128;

struct HFq {
    uint64_t l[4];

    static constexpr uint64_t MOD[4] = {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL,
                                        0x30644e72e131a029ULL};
    static constexpr uint64_t INV = 0x87d20782e4866389ULL;
    static constexpr uint64_t ONE[4] = {0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL,
                                        0x0e0a77c19a07df2fULL};

    static HFq zero() { return HFq{{0, 0, 0, 0}}; }
    static HFq one() { return HFq{{ONE[0], ONE[1], ONE[2], ONE[3]}}; }
    bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
    bool operator==(const HFq &o) const { return memcmp(l, o.l, 32) == 0; }

    static bool geq_mod(const uint64_t a[4])
    {
        for (int i = 3; i >= 0; --i) {
            if (a[i] > MOD[i]) return true;
            if (a[i] < MOD[i]) return false;
        }
        return true;
    }
    static void sub_mod(uint64_t a[4])
    {
        uint64_t borrow = 0;
        for (int i = 0; i < 4; i++) {
            u128 d = (u128)a[i] - MOD[i] - borrow;
            a[i] = (uint64_t)d;
            borrow = (uint64_t)(d >> 64) & 1;
        }
    }
    static HFq mul(const HFq &a, const HFq &b)
    {
        uint64_t t[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 4; i++) {
            u128 carry = 0;
            for (int j = 0; j < 4; j++) {
                u128 s = (u128)a.l[j] * b.l[i] + t[j] + carry;
                t[j] = (uint64_t)s;
                carry = s >> 64;
            }
            u128 s = (u128)t[4] + carry;
            t[4] = (uint64_t)s;
            t[5] = (uint64_t)(s >> 64);
            const uint64_t m = t[0] * INV;
            carry = ((u128)m * MOD[0] + t[0]) >> 64;
            for (int j = 1; j < 4; j++) {
                s = (u128)m * MOD[j] + t[j] + carry;
                t[j - 1] = (uint64_t)s;
                carry = s >> 64;
            }
            s = (u128)t[4] + carry;
            t[3] = (uint64_t)s;
            t[4] = t[5] + (uint64_t)(s >> 64);
        }
        if (t[4] || geq_mod(t)) sub_mod(t);
        return HFq{{t[0], t[1], t[2], t[3]}};
    }
    static HFq sqr(const HFq &a) { return mul(a, a); }
    static HFq add(const HFq &a, const HFq &b)
    {
        uint64_t t[4];
        uint64_t carry = 0;
        for (int i = 0; i < 4; i++) {
            u128 s = (u128)a.l[i] + b.l[i] + carry;
            t[i] = (uint64_t)s;
            carry = (uint64_t)(s >> 64);
        }
        if (carry || geq_mod(t)) sub_mod(t);
        return HFq{{t[0], t[1], t[2], t[3]}};
    }
    static HFq dbl(const HFq &a) { return add(a, a); }
    static HFq sub(const HFq &a, const HFq &b)
    {
        uint64_t t[4];
        uint64_t borrow = 0;
        for (int i = 0; i < 4; i++) {
            u128 d = (u128)a.l[i] - b.l[i] - borrow;
            t[i] = (uint64_t)d;
            borrow = (uint64_t)(d >> 64) & 1;
        }
        if (borrow) {
            uint64_t carry = 0;
            for (int i = 0; i < 4; i++) {
                u128 s = (u128)t[i] + MOD[i] + carry;
                t[i] = (uint64_t)s;
                carry = (uint64_t)(s >> 64);
            }
        }
        return HFq{{t[0], t[1], t[2], t[3]}};
    }
    static HFq neg(const HFq &a) { return a.is_zero() ? a : sub(zero(), a); }
    static HFq inv(const HFq &a)
    {
        // a^(q-2)
        uint64_t e[4] = {MOD[0] - 2, MOD[1], MOD[2], MOD[3]};
        HFq r = one(), base = a;
        for (int i = 0; i < 254; i++) {
            if ((e[i >> 6] >> (i & 63)) & 1) r = mul(r, base);
            base = sqr(base);
        }
        return r;
    }
};

struct HFq2 {
    HFq c0, c1;
    static HFq2 zero() { return HFq2{HFq::zero(), HFq::zero()}; }
    static HFq2 one() { return HFq2{HFq::one(), HFq::zero()}; }
    bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    bool operator==(const HFq2 &o) const { return c0 == o.c0 && c1 == o.c1; }
    static HFq2 mul(const HFq2 &x, const HFq2 &y)
    {
        const HFq aA = HFq::mul(x.c0, y.c0), bB = HFq::mul(x.c1, y.c1);
        const HFq s = HFq::mul(HFq::add(x.c0, x.c1), HFq::add(y.c0, y.c1));
        return HFq2{HFq::sub(aA, bB), HFq::sub(HFq::sub(s, aA), bB)};
    }
    static HFq2 sqr(const HFq2 &x)
    {
        const HFq ab = HFq::mul(x.c0, x.c1);
        return HFq2{HFq::mul(HFq::add(x.c0, x.c1), HFq::sub(x.c0, x.c1)), HFq::dbl(ab)};
    }
    static HFq2 add(const HFq2 &a, const HFq2 &b) { return HFq2{HFq::add(a.c0, b.c0), HFq::add(a.c1, b.c1)}; }
    static HFq2 dbl(const HFq2 &a) { return add(a, a); }
    static HFq2 sub(const HFq2 &a, const HFq2 &b) { return HFq2{HFq::sub(a.c0, b.c0), HFq::sub(a.c1, b.c1)}; }
    static HFq2 neg(const HFq2 &a) { return HFq2{HFq::neg(a.c0), HFq::neg(a.c1)}; }
    static HFq2 inv(const HFq2 &x)
    {
        const HFq t = HFq::inv(HFq::add(HFq::sqr(x.c0), HFq::sqr(x.c1)));
        return HFq2{HFq::mul(x.c0, t), HFq::neg(HFq::mul(x.c1, t))};
    }
};

// Jacobian point in the reference layout X|Y|Z (Z == 0 <=> infinity).
template <class F>
struct HJac {
    F x, y, z;
    bool is_inf() const { return z.is_zero(); }
    static HJac inf() { return HJac{F::zero(), F::one(), F::zero()}; }  // alt_bn128 zero (0,1,0)
};

// dbl-2009-l (a = 0)
template <class F>
HJac<F> jac_dbl(const HJac<F> &p)
{
    if (p.is_inf()) return p;
    const F A = F::sqr(p.x), B = F::sqr(p.y), C = F::sqr(B);
    F D = F::sub(F::sub(F::sqr(F::add(p.x, B)), A), C);
    D = F::dbl(D);
    const F E = F::add(F::dbl(A), A), Fv = F::sqr(E);
    HJac<F> r;
    r.x = F::sub(Fv, F::dbl(D));
    const F C8 = F::dbl(F::dbl(F::dbl(C)));
    r.y = F::sub(F::mul(E, F::sub(D, r.x)), C8);
    r.z = F::dbl(F::mul(p.y, p.z));
    return r;
}

// add-2007-bl with the doubling / inverse cases
template <class F>
HJac<F> jac_add(const HJac<F> &p, const HJac<F> &q)
{
    if (p.is_inf()) return q;
    if (q.is_inf()) return p;
    const F Z1Z1 = F::sqr(p.z), Z2Z2 = F::sqr(q.z);
    const F U1 = F::mul(p.x, Z2Z2), U2 = F::mul(q.x, Z1Z1);
    const F S1 = F::mul(p.y, F::mul(q.z, Z2Z2)), S2 = F::mul(q.y, F::mul(p.z, Z1Z1));
    if (U1 == U2) {
        if (S1 == S2) return jac_dbl(p);
        return HJac<F>::inf();
    }
    const F H = F::sub(U2, U1);
    const F I = F::sqr(F::dbl(H));
    const F J = F::mul(H, I);
    const F r = F::dbl(F::sub(S2, S1));
    const F V = F::mul(U1, I);
    HJac<F> o;
    o.x = F::sub(F::sub(F::sqr(r), J), F::dbl(V));
    o.y = F::sub(F::mul(r, F::sub(V, o.x)), F::dbl(F::mul(S1, J)));
    o.z = F::mul(F::sub(F::sub(F::sqr(F::add(p.z, q.z)), Z1Z1), Z2Z2), H);
    return o;
}

// to_affine_coordinates(): (X/Z^2, Y/Z^3, 1), or the curve's zero
template <class F>
HJac<F> jac_normalise(const HJac<F> &p)
{
    if (p.is_inf()) return HJac<F>::inf();
    const F zi = F::inv(p.z);
    const F z2 = F::sqr(zi);
    return HJac<F>{F::mul(p.x, z2), F::mul(p.y, F::mul(z2, zi)), F::one()};
}

// XYZZ (x = X/ZZ, y = Y/ZZZ) -> Jacobian (X ZZ, Y ZZZ, ZZ)
template <class F>
HJac<F> jac_from_xyzz(const F &X, const F &Y, const F &ZZ, const F &ZZZ)
{
    if (ZZ.is_zero()) return HJac<F>::inf();
    return HJac<F>{F::mul(X, ZZ), F::mul(Y, ZZZ), ZZ};
}

}  // namespace host
}  // namespace b200
