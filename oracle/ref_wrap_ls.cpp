// ref_wrap_ls.cpp — C-ABI over the UNMODIFIED reference's Fr-side routines next to the MSMs
// (SURVEY.md §8(f) rows 2 and 3), compiled where the sources lie by oracle/Makefile into
// oracle/_ref/liblsref.so.  TEST INFRASTRUCTURE ONLY.  Every function calls the reference's own
// class / template; nothing is restated here:
//   ref_fr_eval_mle       MultiVPolyT::evalMLE          (src/prototools/polytools.h:207-234)
//   ref_fr_mle_bind       DPMle::pushRandomness         (src/prototools/mle.h:199-210)
//   ref_cppoly_prove_g1   CPPoly::prove                 (src/gadgets/poly.h:45-91) over an installed key
//   ref_fr_fft            libfqfft basic_radix2_domain  (libfqfft/evaluation_domain/domains/basic_radix2_domain.tcc)
// LegoSNARK compiles with CURVE=BN128 only: LFr = bn128 Fr, same Montgomery limbs as
// alt_bn128's.
#include <cstdint>
#include <cstring>
#include <vector>
using namespace std;

#include "poly.h"
#include "mle.h"
#include <libfqfft/evaluation_domain/domains/basic_radix2_domain.hpp>

namespace {
void init_once()
{
    static bool done = false;
    if (done) return;
    libff::inhibit_profiling_info = true;
    libff::inhibit_profiling_counters = true;
    default_ec_pp::init_public_params();
    done = true;
}
vector<LFr> load_fr(const uint64_t *p, size_t n)
{
    vector<LFr> v(n);
    for (size_t i = 0; i < n; i++) memcpy(v[i].mont_repr.data, p + 4 * i, 32);
    return v;
}
void store_fr(uint64_t *p, const LFr &x) { memcpy(p, x.mont_repr.data, 32); }
}  // namespace

extern "C" {

int ref_ls_init(void)
{
    init_once();
    return 0;
}

int ref_fr_eval_mle(const uint64_t *v, const uint64_t *r, size_t d, uint64_t *out)
{
    init_once();
    const vector<LFr> vv = load_fr(v, (size_t)1 << d), rr = load_fr(r, d);
    store_fr(out, MultiVPolyT::evalMLE(vv, rr));
    return 0;
}

// one pushRandomness(r, 0) on a table of 2 * half = 2^d values
int ref_fr_mle_bind(const uint64_t *table, size_t half, const uint64_t *r, uint64_t *out)
{
    init_once();
    size_t d = 0;
    while (((size_t)1 << d) < 2 * half) d++;
    if (((size_t)1 << d) != 2 * half) return 1;
    DPMle mle(d, 2 * half, load_fr(table, 2 * half));
    mle.pushRandomness(load_fr(r, 1)[0], 0);
    for (size_t p = 0; p < half; p++) store_fr(out + 4 * p, mle.getVTable(0, p));
    return 0;
}

// CPPoly::prove with g1s = the given bases (Jacobian X|Y|Z in bn128 layout); witness: d points, affine-normalised
int ref_cppoly_prove_g1(const uint64_t *bases, size_t nbases, const uint64_t *v, const uint64_t *r, size_t d, uint64_t *witness,
                        uint64_t *witnessa)
{
    init_once();
    struct Key : public CommScheme {
        void install(const vector<LG1> &a)
        {
            n = (long)a.size();
            g1s = a;
        }
    } key;
    vector<LG1> g(nbases);
    for (size_t i = 0; i < nbases; i++) memcpy((void *)&g[i].coord[0], bases + 12 * i, 96);
    key.install(g);
    CPPoly cp(&key);
    const vector<LFr> vv = load_fr(v, (size_t)1 << d), rr = load_fr(r, d);
    PolyPf pf;
    CommOut dummy;
    cp.prove(vv, dummy, rr, pf);
    for (size_t i = 0; i < d; i++) {
        LG1 w = pf.witness[i];
        w.to_affine_coordinates();
        memcpy(witness + 12 * i, (const void *)&w.coord[0], 96);
        if (witnessa) {
            LG1 wa = i ? pf.witnessa[i] : pf.witness[i];
            wa.to_affine_coordinates();
            memcpy(witnessa + 12 * i, (const void *)&wa.coord[0], 96);
        }
    }
    return 0;
}

// mode 0 FFT, 1 iFFT, 2 cosetFFT(g), 3 icosetFFT(g), 4 _basic_radix2_FFT(a, omega^-1)
int ref_fr_fft(uint64_t *a, size_t log_n, int mode, const uint64_t *g)
{
    init_once();
    const size_t n = (size_t)1 << log_n;
    vector<LFr> v = load_fr(a, n);
    try {
        libfqfft::basic_radix2_domain<LFr> dom(n);
        const LFr gg = g ? load_fr(g, 1)[0] : LFr::one();
        if (mode == 0) dom.FFT(v);
        else if (mode == 1) dom.iFFT(v);
        else if (mode == 2) dom.cosetFFT(v, gg);
        else if (mode == 3) dom.icosetFFT(v, gg);
        else if (mode == 4) libfqfft::_basic_radix2_FFT(v, libff::get_root_of_unity<LFr>(n).inverse());
        else return 1;
    } catch (...) {
        return 2;
    }
    for (size_t i = 0; i < n; i++) store_fr(a + 4 * i, v[i]);
    return 0;
}

}  // extern "C"
