// add_latency.cu — latency of ONE dependent chain of XYZZ point additions per warp, the regime of the window reduction and of
// every tree sum (DESIGN.md §4 "reduction").  Variants: the out-of-line serial adder (xyzz_add_cold), the inlined one
// (xyzz_add), the quad-cooperative one (xyzz_add_quad); warps per scheduler 1, 2, 4.  Standalone:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o add_latency add_latency.cu && ./add_latency
// prints one JSON line per (field, variant, warps per scheduler): microseconds per addition along the chain.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../legosnark_b200/csrc/msm_kernels.cuh"

using namespace b200;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <class F, int V>
__global__ void __launch_bounds__(128) k_chain(const XYZZ<F> *__restrict__ in, uint32_t steps, XYZZ<F> *__restrict__ out)
{
    // every lane (V < 2) or every quad (V == 2) runs its own chain acc += q, q fixed; lanes of a quad hold the same values
    const uint32_t lane = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t id = V == 2 ? lane >> 2 : lane;
    XYZZ<F> acc = in[2 * (id & 63)], q = in[2 * (id & 63) + 1];
#pragma unroll 1
    for (uint32_t s = 0; s < steps; s++) {
        if (V == 0) xyzz_add_cold(&acc, &q);
        else if (V == 1) xyzz_add(acc, q);
        else xyzz_add_quad(&acc, &q);
    }
    if (acc.is_inf()) out[lane & 127] = acc;  // keeps the chain alive; never true for these inputs
    if (lane == 0) out[128] = acc;
}

template <class F>
__global__ void k_make(XYZZ<F> *pts, const Affine<F> g)
{
    // 128 distinct multiples of g in XYZZ with non-trivial ZZ: (i + 2) g, built by repeated addition
    XYZZ<F> acc = XYZZ<F>::from_affine(g);
    acc = xyzz_dbl(acc);
    for (int i = 0; i < 128; i++) {
        pts[i] = acc;
        xyzz_madd(acc, g.x, g.y, false);
    }
}

template <class F, int V>
void run(const char *field, const char *name, const XYZZ<F> *d_in, XYZZ<F> *d_out, int sms)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int wps : {1, 2, 4}) {
        const uint32_t steps = 64;
        const int blocks = sms * wps;  // 128 threads = 4 warps = one per scheduler
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            CK(cudaEventRecord(e0));
            k_chain<F, V><<<blocks, 128>>>(d_in, steps, d_out);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep && ms < best) best = ms;
        }
        printf("{\"field\": \"%s\", \"adder\": \"%s\", \"warps_per_scheduler\": %d, \"us_per_addition\": %.3f}\n", field, name, wps,
               best * 1e3 / steps);
        fflush(stdout);
    }
}

int main()
{
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    {
        XYZZ<Fq> *d_in, *d_out;
        CK(cudaMalloc(&d_in, 128 * sizeof(XYZZ<Fq>)));
        CK(cudaMalloc(&d_out, 129 * sizeof(XYZZ<Fq>)));
        Affine<Fq> g{Fq::one(), Fq::add(Fq::one(), Fq::one())};  // (1, 2) on y^2 = x^3 + 3
        k_make<Fq><<<1, 1>>>(d_in, g);
        CK(cudaDeviceSynchronize());
        run<Fq, 0>("Fq", "xyzz_add_cold (serial, out of line)", d_in, d_out, p.multiProcessorCount);
        run<Fq, 1>("Fq", "xyzz_add (serial, inlined)", d_in, d_out, p.multiProcessorCount);
        run<Fq, 2>("Fq", "xyzz_add_quad (4 lanes)", d_in, d_out, p.multiProcessorCount);
        // the three adders agree on the chain's end point
        XYZZ<Fq> h[3];
        for (int v = 0; v < 3; v++) {
            if (v == 0) k_chain<Fq, 0><<<1, 32>>>(d_in, 64, d_out);
            if (v == 1) k_chain<Fq, 1><<<1, 32>>>(d_in, 64, d_out);
            if (v == 2) k_chain<Fq, 2><<<1, 32>>>(d_in, 64, d_out);
            CK(cudaMemcpy(&h[v], d_out + 128, sizeof(XYZZ<Fq>), cudaMemcpyDeviceToHost));
        }
        printf("{\"field\": \"Fq\", \"same_end_point\": %s}\n",
               memcmp(&h[0], &h[1], sizeof h[0]) == 0 && memcmp(&h[0], &h[2], sizeof h[0]) == 0 ? "true" : "false");
    }
    {
        XYZZ<Fq2> *d_in, *d_out;
        CK(cudaMalloc(&d_in, 128 * sizeof(XYZZ<Fq2>)));
        CK(cudaMalloc(&d_out, 129 * sizeof(XYZZ<Fq2>)));
        // any pair (x, y) works for the timing: the formulas do not use the curve constant (a = 0) and the chain never closes
        Affine<Fq2> g{Fq2{Fq::one(), Fq::add(Fq::one(), Fq::one())}, Fq2{Fq::add(Fq::one(), Fq::one()), Fq::one()}};
        k_make<Fq2><<<1, 1>>>(d_in, g);
        CK(cudaDeviceSynchronize());
        run<Fq2, 0>("Fq2", "xyzz_add_cold (serial, out of line)", d_in, d_out, p.multiProcessorCount);
        run<Fq2, 2>("Fq2", "xyzz_add_quad (4 lanes)", d_in, d_out, p.multiProcessorCount);
    }
    return 0;
}
