// Host emulation of the device arithmetic headers (legosnark_b200/csrc/*.cuh).
// TEST SCAFFOLDING for `pytest -m "not gpu"`: compiled with g++, the PTX
// primitives of ptx_ops.cuh are emulated with an explicit carry flag, so the
// limb/column logic of field.cuh and the point formulas of curve.cuh are checked
// bit-for-bit against the oracle without a GPU.  Not part of the product.
#include <cstdint>
#include <cstring>

#include "test_ops.cuh"

using namespace b200;

template <class T>
static T ld(const uint64_t *p)
{
    T t;
    memcpy(&t, p, sizeof(T));
    return t;
}
template <class T>
static void st(uint64_t *p, const T &t)
{
    memcpy(p, &t, sizeof(T));
}

extern "C" {

// field: 0 Fq, 1 Fr, 2 Fq2
int emu_field_op(int field, int op, const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out)
{
    for (size_t i = 0; i < n; i++) {
        if (field == 0) {
            Fq y = b ? ld<Fq>(b + 4 * i) : Fq::zero();
            st(out + 4 * i, prime_field_test_op(op, ld<Fq>(a + 4 * i), y));
        } else if (field == 1) {
            Fr y = b ? ld<Fr>(b + 4 * i) : Fr::zero();
            st(out + 4 * i, prime_field_test_op(op, ld<Fr>(a + 4 * i), y));
        } else {
            Fq2 y = b ? ld<Fq2>(b + 8 * i) : Fq2::zero();
            st(out + 8 * i, field_test_op(op, ld<Fq2>(a + 8 * i), y));
        }
    }
    return 0;
}

// The paired forms (field.cuh: Fp::mul2, Fp::mul_add2; curve.cuh: xyzz_madd_paired), measured options of the engine:
// which 0: (a b, c d) by mul2; 1: (a b + c d, a d + c b) by mul_add2.  field: 0 Fq, 1 Fr.
int emu_pair_field_op(int field, int which, const uint64_t *a, const uint64_t *b, const uint64_t *c, const uint64_t *d, size_t n,
                      uint64_t *out0, uint64_t *out1)
{
    for (size_t i = 0; i < n; i++) {
        if (field == 0) {
            const Fq A = ld<Fq>(a + 4 * i), B = ld<Fq>(b + 4 * i), C = ld<Fq>(c + 4 * i), D = ld<Fq>(d + 4 * i);
            Fq r0, r1;
            if (which == 0) Fq::mul2(r0, r1, A, B, C, D);
            else Fq::mul_add2(r0, r1, A, B, C, D, A, D, C, B);
            st(out0 + 4 * i, r0);
            st(out1 + 4 * i, r1);
        } else {
            const Fr A = ld<Fr>(a + 4 * i), B = ld<Fr>(b + 4 * i), C = ld<Fr>(c + 4 * i), D = ld<Fr>(d + 4 * i);
            Fr r0, r1;
            if (which == 0) Fr::mul2(r0, r1, A, B, C, D);
            else Fr::mul_add2(r0, r1, A, B, C, D, A, D, C, B);
            st(out0 + 4 * i, r0);
            st(out1 + 4 * i, r1);
        }
    }
    return 0;
}

// a (Jacobian) + affine(b) (negated when neg != 0) through xyzz_madd_paired; b must have Z == 1 or be zero (G1 only)
int emu_madd_paired_g1(const uint64_t *a, const uint64_t *b, size_t n, int neg, uint64_t *out)
{
    typedef Jacobian<Fq> J;
    for (size_t i = 0; i < n; i++) {
        XYZZ<Fq> acc = XYZZ<Fq>::from_jacobian(ld<J>(a + 12 * i));
        const J q = ld<J>(b + 12 * i);
        if (!q.is_inf()) xyzz_madd_paired(acc, q.x, q.y, neg != 0);
        st(out + 12 * i, acc.to_jacobian());
    }
    return 0;
}

// group: 0 G1, 1 G2
int emu_group_op(int group, int op, const uint64_t *a, const uint64_t *b, size_t n, uint32_t k, uint64_t *out)
{
    for (size_t i = 0; i < n; i++) {
        if (group == 0) {
            typedef Jacobian<Fq> J;
            J y = b ? ld<J>(b + 12 * i) : J::inf();
            st(out + 12 * i, group_test_op(op, ld<J>(a + 12 * i), y, k));
        } else {
            typedef Jacobian<Fq2> J;
            J y = b ? ld<J>(b + 24 * i) : J::inf();
            st(out + 24 * i, group_test_op(op, ld<J>(a + 24 * i), y, k));
        }
    }
    return 0;
}
}
