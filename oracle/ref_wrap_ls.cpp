// ref_wrap_ls.cpp — C-ABI over the UNMODIFIED reference's Fr-side routines next to the MSMs
// (SURVEY.md §8(f) rows 2 and 3), compiled where the sources lie by oracle/Makefile into
// oracle/_ref/liblsref.so.  TEST INFRASTRUCTURE ONLY.  Every function calls the reference's own
// class / template; nothing is restated here:
//   ref_fr_eval_mle       MultiVPolyT::evalMLE          (src/prototools/polytools.h:207-234)
//   ref_fr_mle_bind       DPMle::pushRandomness         (src/prototools/mle.h:199-210)
//   ref_cppoly_prove_g1   CPPoly::prove                 (src/gadgets/poly.h:45-91) over an installed key
//   ref_fr_eq_table       DPBeta::compute_eq_tbl        (src/prototools/mle.h:93-105)
//   ref_fr_matrix_mle     DPMatrixMle::DPMatrixMle      (src/prototools/mle.h:241-259)
//   ref_sumcheck_h_polys  the round loop of CPSumcheck::prove (src/gadgets/sumcheck.cc:56-70): make_new_h_poly
//                         (src/gadgets/sumcheck.h:85-106) + pushRandomness, with DPBeta or DPBetaDummy
//   ref_fr_beta_suffix    DPBeta's suffix table after precomputeAll (src/prototools/mle.h:130-150)
//   ref_fr_step_fft       libfqfft step_radix2_domain   (libfqfft/evaluation_domain/domains/step_radix2_domain.tcc:38-152)
//   ref_fr_step_divide_z  step_radix2_domain::divide_by_Z_on_coset (:213-241)
//   ref_fr_lagrange       basic_radix2_domain / step_radix2_domain ::evaluate_all_lagrange_polynomials (basic_radix2_domain_aux.tcc:183-236, step_radix2_domain.tcc:161-186)
//   ref_fr_fft            libfqfft basic_radix2_domain  (libfqfft/evaluation_domain/domains/basic_radix2_domain.tcc)
// LegoSNARK compiles with CURVE=BN128 only: LFr = bn128 Fr, same Montgomery limbs as
// alt_bn128's.
#include <cstdint>
#include <cstring>
#include <vector>
using namespace std;

#include "poly.h"
#include "mle.h"
#include "sumcheck.h"
#include <libfqfft/evaluation_domain/domains/basic_radix2_domain.hpp>
#include <libfqfft/evaluation_domain/domains/step_radix2_domain.hpp>

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

int ref_fr_step_fft(uint64_t *a, size_t log_big, size_t log_small, int mode, const uint64_t *g)
{
    init_once();
    const size_t m = ((size_t)1 << log_big) + ((size_t)1 << log_small);
    libfqfft::step_radix2_domain<LFr> dom(m);
    vector<LFr> v = load_fr(a, m);
    const LFr gg = g ? load_fr(g, 1)[0] : LFr::one();
    if (mode == 0) dom.FFT(v);
    else if (mode == 1) dom.iFFT(v);
    else if (mode == 2) dom.cosetFFT(v, gg);
    else if (mode == 3) dom.icosetFFT(v, gg);
    else return 1;
    for (size_t i = 0; i < m; i++) store_fr(a + 4 * i, v[i]);
    return 0;
}

int ref_fr_step_divide_z(uint64_t *a, size_t log_big, size_t log_small)
{
    init_once();
    const size_t m = ((size_t)1 << log_big) + ((size_t)1 << log_small);
    libfqfft::step_radix2_domain<LFr> dom(m);
    vector<LFr> v = load_fr(a, m);
    dom.divide_by_Z_on_coset(v);
    for (size_t i = 0; i < m; i++) store_fr(a + 4 * i, v[i]);
    return 0;
}

// evaluate_all_lagrange_polynomials(t): basic_radix2_domain (log_small == (size_t)-1) or step_radix2_domain
int ref_fr_lagrange(uint64_t *out, size_t log_big, size_t log_small, const uint64_t *t)
{
    init_once();
    const LFr tt = load_fr(t, 1)[0];
    vector<LFr> u;
    if (log_small == (size_t)-1) {
        libfqfft::basic_radix2_domain<LFr> dom((size_t)1 << log_big);
        u = dom.evaluate_all_lagrange_polynomials(tt);
    } else {
        libfqfft::step_radix2_domain<LFr> dom(((size_t)1 << log_big) + ((size_t)1 << log_small));
        u = dom.evaluate_all_lagrange_polynomials(tt);
    }
    for (size_t i = 0; i < u.size(); i++) store_fr(out + 4 * i, u[i]);
    return 0;
}

int ref_fr_eq_table(const uint64_t *r, size_t d, uint64_t *out)
{
    init_once();
    const size_t N = (size_t)1 << d;
    Ins dst(N), tmp(N);
    DPBeta::compute_eq_tbl(d, dst, tmp, load_fr(r, d));
    for (size_t p = 0; p < N; p++) store_fr(out + 4 * p, dst[p]);
    return 0;
}

int ref_fr_matrix_mle(const uint64_t *A, const uint64_t *rho, size_t d, uint64_t *v)
{
    init_once();
    const size_t n = (size_t)1 << d;
    DPMatrixMle m(d, n, load_fr(A, n * n), load_fr(rho, d));
    const Ins vv = m.getV();
    for (size_t p = 0; p < n; p++) store_fr(v + 4 * p, vv[p]);
    return 0;
}

int ref_fr_beta_suffix(const uint64_t *rho, size_t d, uint64_t *out)
{
    init_once();
    DPBeta beta(d, load_fr(rho, d));
    for (size_t p = 0; p < ((size_t)1 << (d - 1)); p++) store_fr(out + 4 * p, beta.beta_suff_rho_cur[p]);
    return 0;
}

// h[i] for every round i < d (coefficients, low degree first, `stride` slots of 4 limbs per round; *ncoef = coefficients per
// polynomial: 3 without beta, 4 with); rho == nullptr selects DPBetaDummy (CPSumcheckMatrix::init_beta, sumcheck.h:118-121)
int ref_sumcheck_h_polys(const uint64_t *a, const uint64_t *b, const uint64_t *rho, const uint64_t *r, size_t d, size_t stride,
                         uint64_t *out, size_t *ncoef)
{
    init_once();
    const size_t N = (size_t)1 << d;
    CPSumcheck sc(nullptr, nullptr);
    shared_ptr<DPBeta> beta = rho ? make_shared<DPBeta>(d, load_fr(rho, d)) : shared_ptr<DPBeta>(make_shared<DPBetaDummy>());
    vector<shared_ptr<DPMle>> mles;
    mles.push_back(make_shared<DPMle>(d, N, load_fr(a, N)));
    mles.push_back(make_shared<DPMle>(d, N, load_fr(b, N)));
    const vector<LFr> rr = load_fr(r, d);
    for (size_t i = 0; i < d; i++) {
        const PolyT h = sc.make_new_h_poly(d, i, beta, mles);
        if (h.vRepr.size() > stride) return 1;
        *ncoef = h.vRepr.size();
        for (size_t k = 0; k < h.vRepr.size(); k++) store_fr(out + 4 * (i * stride + k), h.vRepr[k]);
        if (i + 1 < d) {
            beta->pushRandomness(rr[i], i);
            for (auto &m : mles) m->pushRandomness(rr[i], i);
        }
    }
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
