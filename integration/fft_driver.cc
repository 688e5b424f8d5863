// fft_driver.cc — libfqfft's evaluation domains over Fr as libsnark's r1cs_to_qap_witness_map and
// LegoSNARK's Interpolator (LS/prototools/interp.h:61-65) / lipmaa.cc:94-185 use them:
// get_evaluation_domain(m) then iFFT, cosetFFT, divide_by_Z_on_coset, icosetFFT (the QAP witness map's
// sequence, SNK/reductions/r1cs_to_qap/r1cs_to_qap.tcc) and a plain FFT.  Sizes: a power of two
// (basic_radix2_domain) and, when m2 is given, a size that get_evaluation_domain maps to the extended /
// step radix-2 domains, which reach _basic_radix2_FFT with sub-domain roots.  Built three times from this
// one file by integration/Makefile (reference headers single-thread / OpenMP, shim headers).
//
//   fft_{cpu,cpuomp,b200} [log2m = 16] [m2 = 0]
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace std;

#include <libff/common/default_types/ec_pp.hpp>
#include <libff/common/profiling.hpp>
#include <libfqfft/evaluation_domain/get_evaluation_domain.hpp>
#include "harness.h"
using harness::now_ms;
typedef libff::Fr<libff::default_ec_pp> FrT;

static void fp_vec(harness::Fingerprint &fp, const vector<FrT> &v)
{
    for (const FrT &x : v) {
        const auto b = x.as_bigint();
        fp.bytes(b.data, sizeof b.data);
    }
}

static void run(size_t m, const char *impl)
{
    const auto dom = libfqfft::get_evaluation_domain<FrT>(m);
    vector<FrT> a = harness::scalars<FrT>(dom->m, 3);
    const vector<FrT> a0 = a;
    harness::Fingerprint fp;
    double t0 = now_ms();
    dom->FFT(a);
    const double fft_ms = now_ms() - t0;
    fp_vec(fp, a);
    t0 = now_ms();
    dom->iFFT(a);
    const double ifft_ms = now_ms() - t0;
    const bool round_trip = a == a0;
    // the witness map's coset sequence
    const FrT g = FrT::multiplicative_generator;
    t0 = now_ms();
    dom->cosetFFT(a, g);
    dom->divide_by_Z_on_coset(a);
    dom->icosetFFT(a, g);
    const double coset_ms = now_ms() - t0;
    fp_vec(fp, a);
    // warm second FFT (twiddle tables cached on the device in the shim build)
    a = a0;
    t0 = now_ms();
    dom->FFT(a);
    const double fft2_ms = now_ms() - t0;
    // the generators' side of the domain (r1cs_to_qap_instance_map_with_evaluation, r1cs_to_qap.tcc:127-190): all Lagrange
    // coefficients at a random point, and at a point of the domain itself (the reference's unit-vector branch)
    harness::Fingerprint fl;
    const FrT tpt = harness::scalars<FrT>(1, 11)[0];
    t0 = now_ms();
    const vector<FrT> lag = dom->evaluate_all_lagrange_polynomials(tpt);
    const double lagrange_ms = now_ms() - t0;
    fp_vec(fl, lag);
    fp_vec(fl, dom->evaluate_all_lagrange_polynomials(dom->get_domain_element(dom->m / 3)));
    printf("{\"example\": \"fft\", \"impl\": \"%s\", \"m\": %zu, \"domain_m\": %zu, \"fft_ms_first\": %.3f, \"fft_ms\": %.3f, "
           "\"ifft_ms\": %.3f, \"coset_fft_divide_icoset_ms\": %.3f, \"round_trip\": %s, \"fingerprint\": \"%s\", "
           "\"lagrange_ms\": %.3f, \"lagrange_fingerprint\": \"%s\"}\n",
           impl, m, (size_t)dom->m, fft_ms, fft2_ms, ifft_ms, coset_ms, round_trip ? "true" : "false", fp.hex().c_str(), lagrange_ms,
           fl.hex().c_str());
}

int main(int argc, char **argv)
{
    const int l = argc > 1 ? atoi(argv[1]) : 16;
    const size_t m2 = argc > 2 ? (size_t)atoll(argv[2]) : 0;
    libff::inhibit_profiling_info = true;
    libff::inhibit_profiling_counters = true;
    libff::default_ec_pp::init_public_params();
#ifdef B200_SHIM_BASIC_RADIX2_DOMAIN_AUX_HPP_
    const char *impl = "b200";
#elif defined(MULTICORE)
    const char *impl = "libfqfft-cpu-omp";
#else
    const char *impl = "libfqfft-cpu";
#endif
    run((size_t)1 << l, impl);
    if (m2) run(m2, impl);
    return 0;
}
