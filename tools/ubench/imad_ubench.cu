// imad_ubench.cu — integer-pipe microbenchmarks that decide how the Montgomery
// multiplication is written (DESIGN.md "IMAD roofline").  Standalone: built by
// tools/ubench/Makefile, run on the GPU box, prints one JSON line per variant.
//
// Each variant runs `iters` rounds of a fixed instruction pattern in every thread,
// 8 CTAs x 256 threads per SM (full occupancy), and reports warp-instructions per
// cycle per SM sub-partition (SMSP) derived from the CUDA-event time and the SM
// clock read from %clock64 deltas inside the kernel.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../legosnark_b200/csrc/field.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <int V>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t iters, uint32_t seed, long long *clk)
{
    uint32_t a0 = seed + threadIdx.x, a1 = seed * 7 + blockIdx.x, a2 = a0 ^ 0x9e3779b9u, a3 = a1 * 13 + 1;
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 16; i++) r[i] = i * 0x01000193u + threadIdx.x;
    const long long t0 = clock64();
#pragma unroll 1
    for (uint32_t it = 0; it < iters; it++) {
        if (V == 0) {  // 32 IMAD.WIDE.U32, no carries; both multiplicands loop-carried (non-linear: ptxas cannot restructure)
            uint64_t *q = reinterpret_cast<uint64_t *>(r);
#pragma unroll
            for (int rep = 0; rep < 4; rep++) {
#pragma unroll
                for (int j = 0; j < 8; j++)
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(q[j]) : "r"((uint32_t)q[j]), "r"((uint32_t)q[(j + 1) & 7]));
            }
        } else if (V == 1) {  // 2 real Montgomery steps of Fq::mul (field.cuh mont_step): 32 IMAD.WIDE.U32[.X] + 2 IMAD
            b200::detail::mont_step<b200::FqParams>(r, r + 8, r + 8, r[11] ^ a0, false);
            b200::detail::mont_step<b200::FqParams>(r + 8, r, r, r[3] ^ a1, false);
        } else if (V == 2) {  // same 8 products, carry-OUT only on each (no carry-in): isolates the cost of .X
            asm volatile(
                "mad.lo.cc.u32 %0, %16, %17, %0;\n\tmadc.hi.u32 %1, %16, %17, %1;\n\t"
                "mad.lo.cc.u32 %2, %18, %17, %2;\n\tmadc.hi.u32 %3, %18, %17, %3;\n\t"
                "mad.lo.cc.u32 %4, %19, %17, %4;\n\tmadc.hi.u32 %5, %19, %17, %5;\n\t"
                "mad.lo.cc.u32 %6, %16, %17, %6;\n\tmadc.hi.u32 %7, %16, %17, %7;\n\t"
                "mad.lo.cc.u32 %8, %16, %18, %8;\n\tmadc.hi.u32 %9, %16, %18, %9;\n\t"
                "mad.lo.cc.u32 %10, %18, %18, %10;\n\tmadc.hi.u32 %11, %18, %18, %11;\n\t"
                "mad.lo.cc.u32 %12, %19, %18, %12;\n\tmadc.hi.u32 %13, %19, %18, %13;\n\t"
                "mad.lo.cc.u32 %14, %16, %18, %14;\n\tmadc.hi.u32 %15, %16, %18, %15;\n\t"
                : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                  "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                : "r"(a0), "r"(a1), "r"(a2), "r"(a3));
        } else if (V == 3) {  // 16 x 32-bit IMAD (lo) independent
#pragma unroll
            for (int j = 0; j < 16; j++) asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(r[j]) : "r"(r[(j + 3) & 15]));
        } else if (V == 4) {  // 16 x IMAD.HI independent
#pragma unroll
            for (int j = 0; j < 16; j++) asm volatile("mad.hi.u32 %0, %0, %1, %0;" : "+r"(r[j]) : "r"(r[(j + 3) & 15]));
        } else if (V == 5) {  // two chains of 8 IADD3.X (add.cc / addc.cc)
            asm volatile(
                "add.cc.u32 %0, %0, %16;\n\taddc.cc.u32 %1, %1, %17;\n\taddc.cc.u32 %2, %2, %18;\n\taddc.cc.u32 %3, %3, %19;\n\t"
                "addc.cc.u32 %4, %4, %16;\n\taddc.cc.u32 %5, %5, %17;\n\taddc.cc.u32 %6, %6, %18;\n\taddc.u32 %7, %7, %19;\n\t"
                "add.cc.u32 %8, %8, %16;\n\taddc.cc.u32 %9, %9, %17;\n\taddc.cc.u32 %10, %10, %18;\n\taddc.cc.u32 %11, %11, %19;\n\t"
                "addc.cc.u32 %12, %12, %16;\n\taddc.cc.u32 %13, %13, %17;\n\taddc.cc.u32 %14, %14, %18;\n\taddc.u32 %15, %15, %19;\n\t"
                : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                  "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                : "r"(a0), "r"(a1), "r"(a2), "r"(a3));
        } else if (V == 6) {  // 16 independent IADD3 (no carry)
#pragma unroll
            for (int j = 0; j < 16; j++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[j]) : "r"(r[(j + 3) & 15]), "r"(a0));
        } else if (V == 7) {  // 8 IMAD.WIDE (no carry) interleaved with 8 IADD3: do fma and alu pipes dual-issue?
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(*(uint64_t *)&r[j]) : "r"(r[(j + 2) & 7]), "r"(a1));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(r[8 + j]) : "r"(r[9 + j]));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(*(uint64_t *)&r[j]) : "r"(r[(j + 4) & 7]), "r"(a0));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(r[9 + j]) : "r"(r[8 + j]));
            }
        } else if (V == 8) {  // 64-bit column accumulation: 8 x IMAD.WIDE into the SAME pair chain of 2 (dependent depth 4)
#pragma unroll
            for (int j = 0; j < 4; j += 2) {
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(*(uint64_t *)&r[j]) : "r"(r[j + 4]), "r"(a1));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(*(uint64_t *)&r[j]) : "r"(r[j + 5]), "r"(a2));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(*(uint64_t *)&r[j]) : "r"(r[j + 8]), "r"(a3));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(*(uint64_t *)&r[j]) : "r"(r[j + 9]), "r"(a0));
            }
        } else if (V == 9) {  // IMAD.WIDE.U32.X only carry-IN (predicate set once outside): madc.lo + madc.hi (no cc out)
            asm volatile(
                "add.cc.u32 %0, %0, %16;\n\t"
                "madc.lo.cc.u32 %0, %16, %17, %0;\n\tmadc.hi.u32 %1, %16, %17, %1;\n\t"
                "add.cc.u32 %2, %2, %16;\n\t"
                "madc.lo.cc.u32 %2, %18, %17, %2;\n\tmadc.hi.u32 %3, %18, %17, %3;\n\t"
                "add.cc.u32 %4, %4, %16;\n\t"
                "madc.lo.cc.u32 %4, %19, %17, %4;\n\tmadc.hi.u32 %5, %19, %17, %5;\n\t"
                "add.cc.u32 %6, %6, %16;\n\t"
                "madc.lo.cc.u32 %6, %16, %17, %6;\n\tmadc.hi.u32 %7, %16, %17, %7;\n\t"
                : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                  "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                : "r"(a0), "r"(a1), "r"(a2), "r"(a3));
        }
        else if (V == 10) {  // 8 independent DFMA
            double *d = reinterpret_cast<double *>(r);
#pragma unroll
            for (int j = 0; j < 8; j++) asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(d[j]) : "d"(d[(j + 3) & 7]), "d"(1.0000001));
        } else if (V == 11) {  // 4 DFMA interleaved with 4 IMAD.WIDE.U32: do the fp64 and fmaheavy pipes overlap?
            double *d = reinterpret_cast<double *>(r);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(d[j]) : "d"(d[(j + 1) & 3]), "d"(1.0000001));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(*(uint64_t *)&r[8 + 2 * j]) : "r"(r[8 + ((2 * j + 2) & 7)]), "r"(a1));
            }
        } else if (V == 12) {  // 8 IMAD.WIDE.U32 + 8 IMAD (lo): shared pipe?
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(*(uint64_t *)&r[j]) : "r"(r[(j + 2) & 7]), "r"(a1));
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r[8 + j]) : "r"(r[9 + j]), "r"(a0));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(*(uint64_t *)&r[j]) : "r"(r[(j + 4) & 7]), "r"(a0));
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r[9 + j]) : "r"(r[8 + j]), "r"(a1));
            }
        } else if (V == 13) {  // 16 independent FFMA (reference: fmaheavy + fmalite)
            float *f = reinterpret_cast<float *>(r);
#pragma unroll
            for (int j = 0; j < 16; j++) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(f[j]) : "f"(f[(j + 3) & 15]), "f"(1.0001f));
        }
    }
    const long long t1 = clock64();
    uint32_t s = a2;
#pragma unroll
    for (int i = 0; i < 16; i++) s ^= r[i];
    if (s == 0x12345678u) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

struct Variant { const char *name; int ops; const char *what; };

template <int V>
void run(const Variant &v, int sms, uint32_t iters, uint32_t *d_out, long long *d_clk)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int blocks = sms * 8;
    float ms = 0;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(e0));
        k<V><<<blocks, 256>>>(d_out, iters, 12345u, d_clk);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    long long clk = 0;
    CK(cudaMemcpy(&clk, d_clk, 8, cudaMemcpyDeviceToHost));
    const double warp_instr = (double)v.ops * iters * blocks * 8.0;  // 8 warps per CTA
    const double lane_ops_per_s = warp_instr * 32.0 / (ms * 1e-3);
    // cycles: a block's resident time ~ kernel time (one wave); SMSP-cycles = 4 * sms * clk
    const double per_smsp_per_clk = warp_instr / (4.0 * sms * (double)clk);
    printf("{\"variant\": %d, \"name\": \"%s\", \"what\": \"%s\", \"ms\": %.4f, \"clk\": %lld, \"sm_mhz_est\": %.0f, "
           "\"lane_ops_per_s\": %.4e, \"warp_instr_per_clk_per_smsp\": %.4f}\n",
           V, v.name, v.what, ms, clk, clk / (ms * 1e3), lane_ops_per_s, per_smsp_per_clk);
    fflush(stdout);
}

int main(int argc, char **argv)
{
    const uint32_t iters = argc > 1 ? (uint32_t)atoi(argv[1]) : 1u << 14;
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    uint32_t *d_out;
    long long *d_clk;
    CK(cudaMalloc(&d_out, 64));
    CK(cudaMalloc(&d_clk, 64));
    const int sms = p.multiProcessorCount;
    const Variant vs[] = {
        {"imad_wide", 32, "32 IMAD.WIDE.U32 (8 chains x 4), no carries"},
        {"mont_step_x2", 34, "2 Montgomery steps as in Fq::mul: 32 IMAD.WIDE.U32[.X] + 2 IMAD (+ carries)"},
        {"imad_wide_ccout", 8, "8 IMAD.WIDE.U32 with carry-out only"},
        {"imad_lo", 16, "16 independent IMAD (lo)"},
        {"imad_hi", 16, "16 independent IMAD.HI.U32"},
        {"iadd3_x_chain", 16, "2 chains of 8 IADD3.X"},
        {"lop3", 16, "16 independent LOP3"},
        {"imad_wide+lop", 16, "8 IMAD.WIDE.U32 interleaved with 8 LOP3 (xor)"},
        {"imad_wide_dep4", 8, "2 x 4 dependent IMAD.WIDE.U32 on one accumulator"},
        {"imad_wide_x_in", 12, "4 x [IADD3 cc-out + IMAD.WIDE.U32.X carry-in only] (8 fma + 4 alu)"},
        {"dfma", 8, "8 independent DFMA"},
        {"dfma+imad_wide", 8, "4 DFMA interleaved with 4 IMAD.WIDE.U32"},
        {"imad_wide+imad", 16, "8 IMAD.WIDE.U32 interleaved with 8 IMAD"},
        {"ffma", 16, "16 independent FFMA"},
    };
    run<0>(vs[0], sms, iters, d_out, d_clk);
    run<1>(vs[1], sms, iters, d_out, d_clk);
    run<2>(vs[2], sms, iters, d_out, d_clk);
    run<3>(vs[3], sms, iters, d_out, d_clk);
    run<4>(vs[4], sms, iters, d_out, d_clk);
    run<5>(vs[5], sms, iters, d_out, d_clk);
    run<6>(vs[6], sms, iters, d_out, d_clk);
    run<7>(vs[7], sms, iters, d_out, d_clk);
    run<8>(vs[8], sms, iters, d_out, d_clk);
    run<9>(vs[9], sms, iters, d_out, d_clk);
    run<10>(vs[10], sms, iters, d_out, d_clk);
    run<11>(vs[11], sms, iters, d_out, d_clk);
    run<12>(vs[12], sms, iters, d_out, d_clk);
    run<13>(vs[13], sms, iters, d_out, d_clk);
    return 0;
}
