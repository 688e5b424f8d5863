// wire_kernels.cuh — point (de)compression for the reference's stream formats:
// persisting / loading commitment keys made on the device.
//
// Reference (LFF = depends/libsnark/depends/libff/libff): operator<< / operator>> of
//   alt_bn128_G1  LFF/algebra/curves/alt_bn128/alt_bn128_g1.cpp:404-459   (G2: alt_bn128_g2.cpp:414-475)
//   bn128_G1      LFF/algebra/curves/bn128/bn128_g1.cpp:344-463           (G2: bn128_g2.cpp:374-470)
// with point compression on (the default): a point goes out as  is_zero | X | lsb(Y)  after
// to_affine_coordinates(); coming back, Y = +-sqrt(X^3 + b) with the sign chosen by the stored bit.
// The arithmetic is here; the byte framing ('0'/'1' characters, separators) is host work in the shim.
//   flavour 0  alt_bn128:                    X as_bigint (standard form, fp.tcc operator<<), bit = lsb of Y.as_bigint()
//   flavour 1  alt_bn128 -DMONTGOMERY_OUTPUT: X Montgomery image,                           bit = lsb of Y.as_bigint()
//   flavour 2  bn128 -DBINARY_OUTPUT:         X Montgomery image (raw bn::Fp bytes),        bit = lsb of Y's Montgomery image
// G2: X = (c0, c1) / (a_, b_), bit taken from Y.c0 / Y.a_.
#pragma once
#include <cuda_runtime.h>

#include "curve.cuh"

namespace b200 {

template <class F> struct WireField;
template <> struct WireField<Fq> {
    __device__ static Fq to_wire(const Fq &x, int fl) { return fl == 0 ? Fq::from_mont(x) : x; }
    __device__ static Fq from_wire(const Fq &x, int fl) { return fl == 0 ? Fq::to_mont(x) : x; }
    __device__ static uint32_t parity(const Fq &y, int fl) { return (fl == 2 ? y.l[0] : Fq::from_mont(y).l[0]) & 1u; }
    // b = 3 (alt_bn128_init.cpp:134), Montgomery form
    __device__ static Fq coeff_b()
    {
        Fq b;
        const uint32_t v[8] = {0x50ad28d7u, 0x7a17caa9u, 0xe15521b9u, 0x1f6ac17au, 0x696bd284u, 0x334bea4eu, 0xce179d8eu, 0x2a1f6744u};
#pragma unroll
        for (int i = 0; i < 8; i++) b.l[i] = v[i];
        return b;
    }
    // Fp_model::sqrt (fp.tcc, Tonelli-Shanks) with s = 1 (alt_bn128_init.cpp:82): b = a^t = 1 for every square, the
    // loop never runs and the result is a * a^((t-1)/2) = a^((q+1)/4)
    __device__ static Fq sqrt(const Fq &a)
    {
        const uint32_t e[8] = {0xb61f3f52u, 0x4f082305u, 0x5a1c72a3u, 0x65e05aa4u, 0xa0605617u, 0x6e14116du, 0xb84c680au, 0x0c19139cu};
        Fq r = Fq::one();
#pragma unroll 1
        for (int i = 251; i >= 0; i--) {  // (q+1)/4 has 252 bits
            r = Fq::sqr(r);
            if ((e[i >> 5] >> (i & 31)) & 1u) r = Fq::mul(r, a);
        }
        return r;
    }
};
template <> struct WireField<Fq2> {
    __device__ static Fq2 to_wire(const Fq2 &x, int fl) { return Fq2{WireField<Fq>::to_wire(x.c0, fl), WireField<Fq>::to_wire(x.c1, fl)}; }
    __device__ static Fq2 from_wire(const Fq2 &x, int fl) { return Fq2{WireField<Fq>::from_wire(x.c0, fl), WireField<Fq>::from_wire(x.c1, fl)}; }
    __device__ static uint32_t parity(const Fq2 &y, int fl) { return WireField<Fq>::parity(y.c0, fl); }
    // twist b' = 3 / (9 + u) (alt_bn128_init.cpp:135-136)
    __device__ static Fq2 coeff_b()
    {
        Fq2 b;
        const uint32_t c0[8] = {0x77b802a8u, 0x3bf938e3u, 0x3633535du, 0x020b1b27u, 0x49755260u, 0x26b7edf0u, 0x4384a86du, 0x2514c632u};
        const uint32_t c1[8] = {0xd1dcff67u, 0x38e7ecccu, 0x93ce0d3eu, 0x65f0b37du, 0x22ac00aau, 0xd749d0ddu, 0x4a688d4du, 0x0141b9ceu};
#pragma unroll
        for (int i = 0; i < 8; i++) {
            b.c0.l[i] = c0[i];
            b.c1.l[i] = c1[i];
        }
        return b;
    }
    // Fp2_model::sqrt (fp2.tcc:146-200): Tonelli-Shanks with s = 4, t_minus_1_over_2 and nqr_to_t of alt_bn128_init.cpp:92-98
    __device__ static Fq2 sqrt(const Fq2 &a)
    {
        const uint32_t e[16] = {0x113aeb4du, 0x09daa2c5u, 0x684f5608u, 0xe5301039u, 0xe36cb656u, 0x425280c4u, 0xabd09216u, 0x682344f4u,
                                0xe1a6359cu, 0x31376fd2u, 0x88b1bab0u, 0xe5805c2au, 0xe01a4690u, 0xe2ccd37bu, 0xc3b1e5fcu, 0x00492e25u};
        const uint32_t z0[8] = {0x87961532u, 0x801dd976u, 0x3e84d778u, 0xb2fe144bu, 0x98f81824u, 0x936464b8u, 0xad99ce67u, 0x2581f70bu};
        const uint32_t z1[8] = {0x07394ed9u, 0x60b5b575u, 0x808492c9u, 0xf3a19577u, 0xeb1419ecu, 0xd0048196u, 0x9ba98a59u, 0x0e752acfu};
        const Fq2 one = Fq2::one();
        Fq2 z;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            z.c0.l[i] = z0[i];
            z.c1.l[i] = z1[i];
        }
        Fq2 w = one;
#pragma unroll 1
        for (int i = 502; i >= 0; i--) {  // w = a^((t-1)/2), a 503-bit exponent
            w = Fq2::sqr(w);
            if ((e[i >> 5] >> (i & 31)) & 1u) w = Fq2::mul(w, a);
        }
        Fq2 x = Fq2::mul(a, w);
        Fq2 b = Fq2::mul(x, w);  // a^t
        uint32_t v = 4;
        // at most s - 1 rounds for a square; bounded so that a non-square cannot spin (the caller checks x^2 == a)
#pragma unroll 1
        for (int round = 0; round < 4 && b != one; round++) {
            uint32_t m = 0;
            Fq2 b2m = b;
            while (b2m != one && m < v) {
                b2m = Fq2::sqr(b2m);
                m++;
            }
            if (m >= v) break;  // not a square
            w = z;
            for (int j = (int)v - (int)m - 1; j > 0; j--) w = Fq2::sqr(w);
            z = Fq2::sqr(w);
            b = Fq2::mul(b, z);
            x = Fq2::mul(x, w);
            v = m;
        }
        return x;
    }
};

// affine points (from k_ingest) -> X in wire form + flags (bit 0: parity of Y, bit 1: zero).  A zero goes out the way
// the reference writes it: to_affine_coordinates() makes it (0, 1, 0) on both curves (alt_bn128_g1.cpp:60-67,
// bn128_g1.cpp:106-112), so X = 0 and the bit is the parity of one (of its Montgomery image for bn128).
template <class F>
__global__ void __launch_bounds__(128) k_compress(const Affine<F> *__restrict__ pts, const uint8_t *__restrict__ zero, size_t n, int fl,
                                                   F *__restrict__ x_out, uint8_t *__restrict__ flags)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (zero[i]) {
        x_out[i] = F::zero();
        flags[i] = (uint8_t)(2u | WireField<F>::parity(F::one(), fl));
        return;
    }
    const Affine<F> p = pts[i];
    x_out[i] = WireField<F>::to_wire(p.x, fl);
    flags[i] = (uint8_t)WireField<F>::parity(p.y, fl);
}

// X + flags -> (X, Y, 1) Montgomery Jacobian, or the zero (0, 1, 0).  bad[i] = 1 when X^3 + b is not a square.
template <class F>
__global__ void __launch_bounds__(128) k_decompress(const F *__restrict__ x_in, const uint8_t *__restrict__ flags, size_t n, int fl,
                                                     Jacobian<F> *__restrict__ out, uint8_t *__restrict__ bad)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t f = flags[i];
    if (f & 2u) {
        out[i] = Jacobian<F>::inf();
        if (bad) bad[i] = 0;
        return;
    }
    const F x = WireField<F>::from_wire(x_in[i], fl);
    const F y2 = F::add(F::mul(F::sqr(x), x), WireField<F>::coeff_b());
    F y = WireField<F>::sqrt(y2);
    const bool ok = F::sqr(y) == y2;
    if (WireField<F>::parity(y, fl) != (uint32_t)(f & 1u)) y = F::neg(y);
    out[i] = Jacobian<F>{x, y, F::one()};
    if (bad) bad[i] = ok ? 0 : 1;
}

}  // namespace b200
