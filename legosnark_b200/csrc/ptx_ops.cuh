// ptx_ops.cuh — the 32-bit carry-chain instruction vocabulary the field
// arithmetic is written in.
//
// On the device every primitive is ONE PTX instruction issued through
// `asm volatile` (volatile keeps NVVM from reordering the chain; ptxas then sees
// a plain PTX carry chain and fuses each mad.lo.cc/madc.hi.cc pair into a single
// IMAD.WIDE.U32[.X] and each add.cc/addc into IADD3[.X] with a predicate carry).
//
// When the headers are compiled by a host compiler (tests/emu: the CPU unit
// tests of the limb arithmetic) the same primitives are emulated with an
// explicit carry flag, so the column/alignment logic above them is checked
// against the oracle without a GPU.  The host emulation is test scaffolding;
// the product library is built by nvcc and runs the PTX path only.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define B200_HD __host__ __device__ __forceinline__
#define B200_D __device__ __forceinline__
#define B200_COLD __host__ __device__ __noinline__
#else
#define B200_HD inline
#define B200_D inline
#define B200_COLD inline
#endif

namespace b200 {

#if defined(__CUDA_ARCH__)

B200_D uint32_t add_cc(uint32_t a, uint32_t b)
{
    uint32_t r;
    asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
B200_D uint32_t addc_cc(uint32_t a, uint32_t b)
{
    uint32_t r;
    asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
B200_D uint32_t addc(uint32_t a, uint32_t b)
{
    uint32_t r;
    asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
B200_D uint32_t sub_cc(uint32_t a, uint32_t b)
{
    uint32_t r;
    asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
B200_D uint32_t subc_cc(uint32_t a, uint32_t b)
{
    uint32_t r;
    asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
B200_D uint32_t subc(uint32_t a, uint32_t b)
{
    uint32_t r;
    asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
B200_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
B200_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
B200_D uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
B200_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
B200_D uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
B200_D uint32_t mul_hi(uint32_t a, uint32_t b) { return __umulhi(a, b); }

// ---- fused pairs: one asm statement per 32x32->64 multiply-add so that the
// lo/hi halves stay adjacent in the PTX stream (ptxas only forms
// IMAD.WIDE.U32[.X] from adjacent mad.lo.cc / madc.hi.cc pairs). ----
// (lo,hi) = a*b + (clo,chi); carry-out to CC
B200_D void wmad_cc(uint32_t &lo, uint32_t &hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi)
{
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %4;\n\tmadc.hi.cc.u32 %1, %2, %3, %5;"
                 : "=r"(lo), "=r"(hi) : "r"(a), "r"(b), "r"(clo), "r"(chi));
}
// (lo,hi) = a*b + (clo,chi) + CC; carry-out to CC
B200_D void wmadc_cc(uint32_t &lo, uint32_t &hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi)
{
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %4;\n\tmadc.hi.cc.u32 %1, %2, %3, %5;"
                 : "=r"(lo), "=r"(hi) : "r"(a), "r"(b), "r"(clo), "r"(chi));
}
// (lo,hi) = a*b + (clo,chi) + CC; no carry-out
B200_D void wmadc(uint32_t &lo, uint32_t &hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi)
{
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %4;\n\tmadc.hi.u32 %1, %2, %3, %5;"
                 : "=r"(lo), "=r"(hi) : "r"(a), "r"(b), "r"(clo), "r"(chi));
}

#else  // ---- host emulation (tests/emu only) --------------------------------

namespace emu {
inline uint32_t &cf()
{
    static thread_local uint32_t flag = 0;
    return flag;
}
inline uint32_t add3(uint32_t a, uint32_t b, uint32_t cin, bool set)
{
    uint64_t s = (uint64_t)a + b + cin;
    if (set) cf() = (uint32_t)(s >> 32);
    return (uint32_t)s;
}
inline uint32_t sub3(uint32_t a, uint32_t b, uint32_t bin, bool set)
{
    uint64_t d = (uint64_t)a - b - bin;
    if (set) cf() = (uint32_t)((d >> 32) & 1);  // PTX: CC.CF holds the borrow for sub.cc/subc
    return (uint32_t)d;
}
inline uint32_t lo(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b); }
inline uint32_t hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
}  // namespace emu

inline uint32_t add_cc(uint32_t a, uint32_t b) { return emu::add3(a, b, 0, true); }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { return emu::add3(a, b, emu::cf(), true); }
inline uint32_t addc(uint32_t a, uint32_t b) { return emu::add3(a, b, emu::cf(), false); }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { return emu::sub3(a, b, 0, true); }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { return emu::sub3(a, b, emu::cf(), true); }
inline uint32_t subc(uint32_t a, uint32_t b) { return emu::sub3(a, b, emu::cf(), false); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return emu::add3(emu::lo(a, b), c, 0, true); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return emu::add3(emu::lo(a, b), c, emu::cf(), true); }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return emu::add3(emu::hi(a, b), c, 0, true); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return emu::add3(emu::hi(a, b), c, emu::cf(), true); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return emu::add3(emu::hi(a, b), c, emu::cf(), false); }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return emu::hi(a, b); }

inline void wmad_cc(uint32_t &lo, uint32_t &hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi)
{
    const uint32_t l = mad_lo_cc(a, b, clo);
    hi = madc_hi_cc(a, b, chi);
    lo = l;
}
inline void wmadc_cc(uint32_t &lo, uint32_t &hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi)
{
    const uint32_t l = madc_lo_cc(a, b, clo);
    hi = madc_hi_cc(a, b, chi);
    lo = l;
}
inline void wmadc(uint32_t &lo, uint32_t &hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi)
{
    const uint32_t l = madc_lo_cc(a, b, clo);
    hi = madc_hi(a, b, chi);
    lo = l;
}

#endif

}  // namespace b200
