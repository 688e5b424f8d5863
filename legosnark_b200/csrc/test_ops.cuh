// test_ops.cuh — element-wise dispatch used by the parity hooks
// (b200_test_field_op / b200_test_group_op in the C-ABI, and tests/emu on the
// host).  One definition shared by the CUDA kernels and the host emulation so
// both exercise the same arithmetic sources.
#pragma once
#include "curve.cuh"

namespace b200 {

// op: 0 mul, 1 sqr, 2 add, 3 sub, 4 inverse, 5 neg, 6 from_mont (as_bigint), 7 to_mont,
//     8 a b - b (a + b)  (the fused two-product form mul_sub)
template <class F>
B200_HD F field_test_op(int op, const F &a, const F &b)
{
    switch (op) {
    case 0: return F::mul(a, b);
    case 1: return F::sqr(a);
    case 2: return F::add(a, b);
    case 3: return F::sub(a, b);
    case 4: return F::inv(a);
    case 5: return F::neg(a);
    case 8: return F::mul_sub(a, b, b, F::add(a, b));
    default: return a;
    }
}

template <class P>
B200_HD Fp<P> prime_field_test_op(int op, const Fp<P> &a, const Fp<P> &b)
{
    if (op == 6) return Fp<P>::from_mont(a);
    if (op == 7) return Fp<P>::to_mont(a);
    return field_test_op<Fp<P>>(op, a, b);
}

// Inputs/outputs are Jacobian in the reference layout.
// op: 0 a + b (XYZZ add), 1 a + affine(b) (mixed; b must have Z == 1 or be zero),
//     2 2a, 6 a - affine(b), 7 k*a with k = b.x limb 0 (32 bits), 8 round trip
template <class F>
B200_HD Jacobian<F> group_test_op(int op, const Jacobian<F> &a, const Jacobian<F> &b, uint32_t k)
{
    XYZZ<F> acc = XYZZ<F>::from_jacobian(a);
    switch (op) {
    case 0: {
        const XYZZ<F> q = XYZZ<F>::from_jacobian(b);
        xyzz_add_cold(&acc, &q);
        break;
    }
    case 1:
    case 6:
        if (!b.is_inf()) {
            const Affine<F> q{b.x, b.y};
            xyzz_madd_cold(&acc, &q, op == 6);
        }
        break;
    case 2: xyzz_dbl_cold(&acc); break;
    case 7: acc = xyzz_mul_small(acc, k); break;
    default: break;
    }
    return acc.to_jacobian();
}

}  // namespace b200
