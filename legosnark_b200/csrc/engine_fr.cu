// engine_fr.cu — host side of the Fr vector kernels (fr_kernels.cuh): multilinear folding
// (CPPoly::prove / evalMLE / DPMle) and libfqfft's basic radix-2 domain.
#include "engine_common.hpp"
#include "host_copy.hpp"
#include "fr_kernels.cuh"

namespace b200 {
namespace eng {

// ------------------------------------------------------------------------------
// folding
// ------------------------------------------------------------------------------
const void *fr_fold_device(Device &D, const uint64_t *v, const uint64_t *r, size_t d, bool want_w)
{
    const size_t N = (size_t)1 << d;
    cudaStream_t st = D.stream;
    D.fr_a.ensure(N * sizeof(Fr));
    D.fr_b.ensure(std::max<size_t>(N / 2, 1) * sizeof(Fr));
    D.fr_r.ensure(std::max<size_t>(d, 1) * sizeof(Fr));
    if (want_w) {
        D.fr_w.ensure(N * sizeof(Fr));
        // w_coeffs is a zero-initialised vector of 2^d entries of which 2^d - 1 are written (poly.h:52)
        CK(cudaMemsetAsync((char *)D.fr_w.p + (N - 1) * sizeof(Fr), 0, sizeof(Fr), st));
    }
    h2d(D, D.fr_a.p, v, N * sizeof(Fr), st);
    if (d) CK(cudaMemcpyAsync(D.fr_r.p, r, d * sizeof(Fr), cudaMemcpyHostToDevice, st));
    Fr *cur = D.fr_a.as<Fr>(), *nxt = D.fr_b.as<Fr>();
    for (uint32_t i0 = 0; i0 < d;) {
        const uint32_t nl = std::min<uint32_t>(FOLD_LEVELS, (uint32_t)d - i0);
        const size_t pairs = (size_t)1 << (d - i0 - 1);
        LAUNCH(D, k_fr_fold, cdiv(pairs, FOLD_THREADS), FOLD_THREADS, 0, st, (const Fr *)cur, D.fr_r.as<Fr>(), (uint32_t)d, i0, nl, nxt,
               want_w ? D.fr_w.as<Fr>() : (Fr *)nullptr);
        std::swap(cur, nxt);
        i0 += nl;
    }
    return cur;
}

int fr_fold_witness(const uint64_t *v, const uint64_t *r, size_t d, uint64_t *w_coeffs, uint64_t *eval)
{
    if (!g_init) return fail(B200_ERR_NOT_INIT, "b200_init has not been called (no CUDA device => no result: there is no CPU fallback)");
    if (!v || (d && !r) || (!w_coeffs && !eval)) return fail(B200_ERR_ARG, "null argument");
    if (d > 30) return fail(B200_ERR_ARG, "d out of range");
    try {
        Device &D = g_devs[0];
        D.launches = 0;
        CK(cudaSetDevice(D.id));
        const size_t N = (size_t)1 << d;
        const void *fin = fr_fold_device(D, v, r, d, w_coeffs != nullptr);
        if (w_coeffs) d2h(D, w_coeffs, D.fr_w.p, N * sizeof(Fr), D.stream);
        if (eval) CK(cudaMemcpyAsync(eval, fin, sizeof(Fr), cudaMemcpyDeviceToHost, D.stream));
        CK(cudaStreamSynchronize(D.stream));
        g_stats = b200_stats_t{};
        g_stats.n = N;
        g_stats.kernel_launches = D.launches;
        g_stats.h2d_bytes = (double)(N + d) * sizeof(Fr);
        g_stats.d2h_bytes = (double)((w_coeffs ? N : 0) + (eval ? 1 : 0)) * sizeof(Fr);
        return B200_OK;
    } catch (const CudaError &e) {
        return fail(B200_ERR_CUDA, "%s", e.msg.c_str());
    } catch (const std::exception &ex) {  // nothing may unwind through the extern "C" boundary
        return fail(B200_ERR_CUDA, "host error: %s", ex.what());
    } catch (...) {
        return fail(B200_ERR_CUDA, "unknown host error");
    }
}

int fr_mle_bind(const uint64_t *table, size_t half, const uint64_t *r, uint64_t *out)
{
    if (!g_init) return fail(B200_ERR_NOT_INIT, "b200_init has not been called");
    if (half == 0) return B200_OK;
    if (!table || !r || !out) return fail(B200_ERR_ARG, "null argument");
    try {
        Device &D = g_devs[0];
        D.launches = 0;
        CK(cudaSetDevice(D.id));
        cudaStream_t st = D.stream;
        D.fr_a.ensure(2 * half * sizeof(Fr));
        D.fr_b.ensure(half * sizeof(Fr));
        D.fr_r.ensure(sizeof(Fr));
        h2d(D, D.fr_a.p, table, 2 * half * sizeof(Fr), st);
        CK(cudaMemcpyAsync(D.fr_r.p, r, sizeof(Fr), cudaMemcpyHostToDevice, st));
        LAUNCH(D, k_fr_bind_hi, cdiv(half, 256), 256, 0, st, D.fr_a.as<Fr>(), half, D.fr_r.as<Fr>(), D.fr_b.as<Fr>());
        d2h(D, out, D.fr_b.p, half * sizeof(Fr), st);
        CK(cudaStreamSynchronize(st));
        g_stats = b200_stats_t{};
        g_stats.n = half;
        g_stats.kernel_launches = D.launches;
        return B200_OK;
    } catch (const CudaError &e) {
        return fail(B200_ERR_CUDA, "%s", e.msg.c_str());
    } catch (const std::exception &ex) {  // nothing may unwind through the extern "C" boundary
        return fail(B200_ERR_CUDA, "host error: %s", ex.what());
    } catch (...) {
        return fail(B200_ERR_CUDA, "unknown host error");
    }
}

// ------------------------------------------------------------------------------
// sum-check tables (mle.h / sumcheck.h)
// ------------------------------------------------------------------------------
// eq table of d challenges already in D.fr_r -> returned device pointer (D.fr_a or D.fr_b), 2^d entries
static const Fr *eq_table_device(Device &D, cudaStream_t st, size_t d)
{
    const size_t N = (size_t)1 << d;
    D.fr_a.ensure(N * sizeof(Fr));
    D.fr_b.ensure(N * sizeof(Fr));
    Fr *cur = D.fr_a.as<Fr>(), *nxt = D.fr_b.as<Fr>();
    for (uint32_t j = 0; j < d; j++) {
        LAUNCH(D, k_fr_eq_step, cdiv((size_t)2 << j, 256), 256, 0, st, (const Fr *)cur, (const Fr *)D.fr_r.as<Fr>(), j, nxt);
        std::swap(cur, nxt);
    }
    return cur;
}

int fr_eq_table(const uint64_t *r, size_t d, uint64_t *out)
{
    if (!g_init) return fail(B200_ERR_NOT_INIT, "b200_init has not been called (no CUDA device => no result: there is no CPU fallback)");
    if (!r || !out) return fail(B200_ERR_ARG, "null argument");
    if (d < 1 || d > 28) return fail(B200_ERR_ARG, "d out of range (1..28)");
    try {
        Device &D = g_devs[0];
        D.launches = 0;
        CK(cudaSetDevice(D.id));
        cudaStream_t st = D.stream;
        D.fr_r.ensure(d * sizeof(Fr));
        CK(cudaMemcpyAsync(D.fr_r.p, r, d * sizeof(Fr), cudaMemcpyHostToDevice, st));
        const Fr *tbl = eq_table_device(D, st, d);
        d2h(D, out, tbl, ((size_t)1 << d) * sizeof(Fr), st);
        CK(cudaStreamSynchronize(st));
        g_stats = b200_stats_t{};
        g_stats.n = (size_t)1 << d;
        g_stats.kernel_launches = D.launches;
        return B200_OK;
    } catch (const CudaError &e) {
        return fail(B200_ERR_CUDA, "%s", e.msg.c_str());
    } catch (const std::exception &ex) {  // nothing may unwind through the extern "C" boundary
        return fail(B200_ERR_CUDA, "host error: %s", ex.what());
    } catch (...) {
        return fail(B200_ERR_CUDA, "unknown host error");
    }
}

int fr_matrix_mle(const uint64_t *A, const uint64_t *rho, size_t d, uint64_t *v)
{
    if (!g_init) return fail(B200_ERR_NOT_INIT, "b200_init has not been called (no CUDA device => no result: there is no CPU fallback)");
    if (!A || !rho || !v) return fail(B200_ERR_ARG, "null argument");
    if (d < 1 || d > 14) return fail(B200_ERR_ARG, "d out of range (1..14)");
    try {
        Device &D = g_devs[0];
        D.launches = 0;
        CK(cudaSetDevice(D.id));
        cudaStream_t st = D.stream;
        const size_t n = (size_t)1 << d, N = n * n;
        D.fr_r.ensure(d * sizeof(Fr));
        D.fr_w.ensure(N * sizeof(Fr));
        CK(cudaMemcpyAsync(D.fr_r.p, rho, d * sizeof(Fr), cudaMemcpyHostToDevice, st));
        h2d(D, D.fr_w.p, A, N * sizeof(Fr), st);
        const Fr *eq = eq_table_device(D, st, d);
        // rows split over gridDim.y so that ~4 blocks per SM are in flight
        const uint32_t bx = cdiv(n, 32);
        uint32_t ny = std::max<uint32_t>(1, std::min<uint32_t>((uint32_t)(n / 8), (uint32_t)D.sms * 4 / bx));
        D.scalars.ensure((size_t)ny * n * sizeof(Fr));
        D.prefix.ensure(n * sizeof(Fr));
        LAUNCH(D, k_fr_matrix_mle, dim3(bx, ny), MMLE_THREADS, 0, st, (const Fr *)D.fr_w.as<Fr>(), eq, (uint32_t)d, D.scalars.as<Fr>());
        LAUNCH(D, k_fr_vec_add_slices, cdiv(n, 256), 256, 0, st, (const Fr *)D.scalars.as<Fr>(), n, ny, D.prefix.as<Fr>());
        CK(cudaMemcpyAsync(v, D.prefix.p, n * sizeof(Fr), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        g_stats = b200_stats_t{};
        g_stats.n = N;
        g_stats.kernel_launches = D.launches;
        g_stats.h2d_bytes = (double)(N + d) * sizeof(Fr);
        g_stats.d2h_bytes = (double)n * sizeof(Fr);
        return B200_OK;
    } catch (const CudaError &e) {
        return fail(B200_ERR_CUDA, "%s", e.msg.c_str());
    } catch (const std::exception &ex) {  // nothing may unwind through the extern "C" boundary
        return fail(B200_ERR_CUDA, "host error: %s", ex.what());
    } catch (...) {
        return fail(B200_ERR_CUDA, "unknown host error");
    }
}

static void sumcheck_round_device(Device &D, cudaStream_t st, const Fr *a, const Fr *b, const Fr *w, size_t half, Fr *out3)
{
    const uint32_t blocks = std::max<uint32_t>(1, std::min<uint32_t>(cdiv(half, SC_THREADS), (uint32_t)D.sms * 4));
    D.coeff.ensure((size_t)blocks * 3 * sizeof(Fr));
    LAUNCH(D, k_fr_sumcheck_round, blocks, SC_THREADS, 0, st, a, b, w, half, D.coeff.as<Fr>());
    LAUNCH(D, k_fr_sum3, 1, 96, 0, st, (const Fr *)D.coeff.as<Fr>(), blocks, out3);
}

int fr_sumcheck_round(const uint64_t *a, const uint64_t *b, const uint64_t *w, size_t half, uint64_t *out)
{
    if (!g_init) return fail(B200_ERR_NOT_INIT, "b200_init has not been called (no CUDA device => no result: there is no CPU fallback)");
    if (!a || !b || !out || half == 0) return fail(B200_ERR_ARG, "bad argument");
    try {
        Device &D = g_devs[0];
        D.launches = 0;
        CK(cudaSetDevice(D.id));
        cudaStream_t st = D.stream;
        D.fr_a.ensure(2 * half * sizeof(Fr));
        D.fr_b.ensure(2 * half * sizeof(Fr));
        D.fr_w.ensure(half * sizeof(Fr));
        D.fr_r.ensure(3 * sizeof(Fr));
        h2d(D, D.fr_a.p, a, 2 * half * sizeof(Fr), st);
        h2d(D, D.fr_b.p, b, 2 * half * sizeof(Fr), st);
        if (w) h2d(D, D.fr_w.p, w, half * sizeof(Fr), st);
        sumcheck_round_device(D, st, D.fr_a.as<Fr>(), D.fr_b.as<Fr>(), w ? D.fr_w.as<Fr>() : (const Fr *)nullptr, half, D.fr_r.as<Fr>());
        CK(cudaMemcpyAsync(out, D.fr_r.p, 3 * sizeof(Fr), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        g_stats = b200_stats_t{};
        g_stats.n = half;
        g_stats.kernel_launches = D.launches;
        return B200_OK;
    } catch (const CudaError &e) {
        return fail(B200_ERR_CUDA, "%s", e.msg.c_str());
    } catch (const std::exception &ex) {  // nothing may unwind through the extern "C" boundary
        return fail(B200_ERR_CUDA, "host error: %s", ex.what());
    } catch (...) {
        return fail(B200_ERR_CUDA, "unknown host error");
    }
}

// all d rounds of the sum-check over two tables without a beta factor (CPSumcheckMatrix: DPBetaDummy,
// sumcheck.h:118-131): round i's three sums, then both tables are bound to r[i] on the device
// (DPMle::pushRandomness, mle.h:199-210); the tables never return to the host.
int fr_sumcheck_rounds(const uint64_t *a, const uint64_t *b, const uint64_t *r, size_t d, uint64_t *h)
{
    if (!g_init) return fail(B200_ERR_NOT_INIT, "b200_init has not been called (no CUDA device => no result: there is no CPU fallback)");
    if (!a || !b || !r || !h) return fail(B200_ERR_ARG, "null argument");
    if (d < 1 || d > 28) return fail(B200_ERR_ARG, "d out of range (1..28)");
    try {
        Device &D = g_devs[0];
        D.launches = 0;
        CK(cudaSetDevice(D.id));
        cudaStream_t st = D.stream;
        const size_t N = (size_t)1 << d;
        D.fr_a.ensure(N * sizeof(Fr));
        D.fr_b.ensure(N * sizeof(Fr));
        D.fr_w.ensure(N * sizeof(Fr));  // ping-pong partners: a in fr_w[0, N/2), b in fr_w[N/2, N)
        D.fr_r.ensure(d * sizeof(Fr));
        D.out_norm.ensure(d * 3 * sizeof(Fr));
        h2d(D, D.fr_a.p, a, N * sizeof(Fr), st);
        h2d(D, D.fr_b.p, b, N * sizeof(Fr), st);
        CK(cudaMemcpyAsync(D.fr_r.p, r, d * sizeof(Fr), cudaMemcpyHostToDevice, st));
        Fr *ca = D.fr_a.as<Fr>(), *cb = D.fr_b.as<Fr>();
        Fr *na = D.fr_w.as<Fr>(), *nb = D.fr_w.as<Fr>() + N / 2;
        for (size_t i = 0; i < d; i++) {
            const size_t half = (size_t)1 << (d - i - 1);
            sumcheck_round_device(D, st, ca, cb, nullptr, half, D.out_norm.as<Fr>() + 3 * i);
            if (i + 1 < d) {
                LAUNCH(D, k_fr_bind_hi, cdiv(half, 256), 256, 0, st, (const Fr *)ca, half, (const Fr *)(D.fr_r.as<Fr>() + i), na);
                LAUNCH(D, k_fr_bind_hi, cdiv(half, 256), 256, 0, st, (const Fr *)cb, half, (const Fr *)(D.fr_r.as<Fr>() + i), nb);
                std::swap(ca, na);
                std::swap(cb, nb);
            }
        }
        CK(cudaMemcpyAsync(h, D.out_norm.p, d * 3 * sizeof(Fr), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        g_stats = b200_stats_t{};
        g_stats.n = N;
        g_stats.kernel_launches = D.launches;
        g_stats.h2d_bytes = (double)(2 * N + d) * sizeof(Fr);
        g_stats.d2h_bytes = (double)d * 3 * sizeof(Fr);
        return B200_OK;
    } catch (const CudaError &e) {
        return fail(B200_ERR_CUDA, "%s", e.msg.c_str());
    } catch (const std::exception &ex) {  // nothing may unwind through the extern "C" boundary
        return fail(B200_ERR_CUDA, "host error: %s", ex.what());
    } catch (...) {
        return fail(B200_ERR_CUDA, "unknown host error");
    }
}

// P[i] *= (c1 ratio^i - c0)^-1 for i < n_geo, then P[n_geo + i] *= tail for i < n_tail (tail may be NULL):
// step_radix2_domain::divide_by_Z_on_coset (step_radix2_domain.tcc:213-241) with the constants formed by the caller
int fr_scale_inv_geometric(uint64_t *P, size_t n_geo, const uint64_t *c1, const uint64_t *ratio, const uint64_t *c0, size_t n_tail,
                           const uint64_t *tail)
{
    if (!g_init) return fail(B200_ERR_NOT_INIT, "b200_init has not been called (no CUDA device => no result: there is no CPU fallback)");
    if (!P || !c1 || !ratio || !c0 || (n_tail && !tail)) return fail(B200_ERR_ARG, "null argument");
    const size_t n = n_geo + n_tail;
    if (n == 0) return B200_OK;
    try {
        Device &D = g_devs[0];
        D.launches = 0;
        CK(cudaSetDevice(D.id));
        cudaStream_t st = D.stream;
        D.fr_a.ensure(n * sizeof(Fr));
        D.fr_r.ensure(4 * sizeof(Fr));
        uint64_t consts[16];
        memcpy(consts, c1, 32);
        memcpy(consts + 4, ratio, 32);
        memcpy(consts + 8, c0, 32);
        if (tail) memcpy(consts + 12, tail, 32);
        h2d(D, D.fr_a.p, P, n * sizeof(Fr), st);
        CK(cudaMemcpyAsync(D.fr_r.p, consts, sizeof consts, cudaMemcpyHostToDevice, st));
        if (n_geo)
            LAUNCH(D, k_fr_scale_inv_geometric, cdiv(cdiv(n_geo, INVG_RUN), 128), 128, 0, st, D.fr_a.as<Fr>(), n_geo, (const Fr *)D.fr_r.as<Fr>());
        if (n_tail) LAUNCH(D, k_fr_scale, cdiv(n_tail, 256), 256, 0, st, D.fr_a.as<Fr>() + n_geo, n_tail, (const Fr *)(D.fr_r.as<Fr>() + 3));
        d2h(D, P, D.fr_a.p, n * sizeof(Fr), st);
        CK(cudaStreamSynchronize(st));  // consts lives on this frame
        g_stats = b200_stats_t{};
        g_stats.n = n;
        g_stats.kernel_launches = D.launches;
        g_stats.h2d_bytes = g_stats.d2h_bytes = (double)n * sizeof(Fr);
        return B200_OK;
    } catch (const CudaError &e) {
        return fail(B200_ERR_CUDA, "%s", e.msg.c_str());
    } catch (const std::exception &ex) {  // nothing may unwind through the extern "C" boundary
        return fail(B200_ERR_CUDA, "host error: %s", ex.what());
    } catch (...) {
        return fail(B200_ERR_CUDA, "unknown host error");
    }
}

// out[i] = in[i] * a0 * a_ratio^i / prod_f (c1_f ratio_f^i - c0_f): the Lagrange-coefficient vectors of libfqfft's radix-2 and
// step domains (basic_radix2_domain_aux.tcc:183-236, step_radix2_domain.tcc:161-186) with the constants formed by the caller.
// consts = a0, a_ratio, then (c1, ratio, c0) per factor: (2 + 3 nf) x 4 limbs.  in may be NULL (ones) or equal to out.
int fr_geometric_quotients(uint64_t *out, const uint64_t *in, size_t n, const uint64_t *consts, size_t nf)
{
    if (!g_init) return fail(B200_ERR_NOT_INIT, "b200_init has not been called (no CUDA device => no result: there is no CPU fallback)");
    if (!out || !consts) return fail(B200_ERR_ARG, "null argument");
    if (nf < 1 || nf > 2) return fail(B200_ERR_ARG, "one or two denominator factors, got %zu", nf);
    if (n == 0) return B200_OK;
    try {
        Device &D = g_devs[0];
        D.launches = 0;
        CK(cudaSetDevice(D.id));
        cudaStream_t st = D.stream;
        D.fr_a.ensure(n * sizeof(Fr));
        D.fr_r.ensure(8 * sizeof(Fr));
        if (in) h2d(D, D.fr_a.p, in, n * sizeof(Fr), st);
        CK(cudaMemcpyAsync(D.fr_r.p, consts, (2 + 3 * nf) * sizeof(Fr), cudaMemcpyHostToDevice, st));
        LAUNCH(D, k_fr_geometric_quotients, cdiv(cdiv(n, INVG_RUN), 128), 128, 0, st, in ? (const Fr *)D.fr_a.as<Fr>() : (const Fr *)nullptr,
               D.fr_a.as<Fr>(), n, (const Fr *)D.fr_r.as<Fr>(), (int)nf);
        d2h(D, out, D.fr_a.p, n * sizeof(Fr), st);
        CK(cudaStreamSynchronize(st));  // consts belongs to the caller
        g_stats = b200_stats_t{};
        g_stats.n = n;
        g_stats.kernel_launches = D.launches;
        g_stats.h2d_bytes = in ? (double)n * sizeof(Fr) : 0.0;
        g_stats.d2h_bytes = (double)n * sizeof(Fr);
        return B200_OK;
    } catch (const CudaError &e) {
        return fail(B200_ERR_CUDA, "%s", e.msg.c_str());
    } catch (const std::exception &ex) {  // nothing may unwind through the extern "C" boundary
        return fail(B200_ERR_CUDA, "host error: %s", ex.what());
    } catch (...) {
        return fail(B200_ERR_CUDA, "unknown host error");
    }
}

// ------------------------------------------------------------------------------
// radix-2 domains: twiddle tables cached per (device, log n); coset tables per last shift g
// ------------------------------------------------------------------------------
struct FrDomain {
    DevBuf consts;              // omega, omega^-1, n^-1, g, g^-1
    DevBuf tw_fwd, tw_inv;      // omega^i, omega^-i, i < n/2
    DevBuf coset_pre;           // g^i, i < n
    DevBuf coset_post;          // n^-1 g^-i
    DevBuf g_dev;
    bool has_g = false;
    uint64_t g[4] = {0, 0, 0, 0};
};
static std::vector<std::map<uint32_t, std::unique_ptr<FrDomain>>> g_domains;

void fr_step_release();
void fr_release()
{
    fr_step_release();
    for (size_t di = 0; di < g_domains.size(); di++) {
        if (di < g_devs.size()) cudaSetDevice(g_devs[di].id);
        for (auto &kv : g_domains[di]) {
            FrDomain &dm = *kv.second;
            DevBuf *all[] = {&dm.consts, &dm.tw_fwd, &dm.tw_inv, &dm.coset_pre, &dm.coset_post, &dm.g_dev};
            for (DevBuf *b : all) b->release();
        }
    }
    g_domains.clear();
}

static FrDomain &domain_for(Device &D, size_t dev_index, uint32_t logn, const uint64_t *g, cudaStream_t st)
{
    if (g_domains.size() < g_devs.size()) g_domains.resize(g_devs.size());
    auto &slot = g_domains[dev_index][logn];
    const size_t n = (size_t)1 << logn;
    const bool fresh = !slot;
    if (fresh) slot = std::make_unique<FrDomain>();
    FrDomain &dm = *slot;
    const bool new_g = g && (!dm.has_g || memcmp(dm.g, g, 32) != 0);
    if (fresh || new_g) {
        dm.consts.ensure(5 * sizeof(Fr));
        dm.g_dev.ensure(sizeof(Fr));
        const Fr *gd = nullptr;
        if (g) {
            // the table kernels run on `st`; a pageable source is copied before the call returns
            CK(cudaMemcpyAsync(dm.g_dev.p, g, sizeof(Fr), cudaMemcpyHostToDevice, st));
            gd = dm.g_dev.as<Fr>();
        } else if (dm.has_g) {
            gd = dm.g_dev.as<Fr>();  // keep the cached shift when a plain transform creates nothing new
        }
        LAUNCH(D, k_fr_domain_consts, 1, 32, 0, st, logn, gd, dm.consts.as<Fr>());
    }
    const Fr *c = dm.consts.as<Fr>();
    if (fresh) {
        const size_t half = std::max<size_t>(n / 2, 1);
        dm.tw_fwd.ensure(half * sizeof(Fr));
        dm.tw_inv.ensure(half * sizeof(Fr));
        LAUNCH(D, k_fr_pow_table, cdiv(cdiv(half, 16), 128), 128, 0, st, c + 0, (const Fr *)nullptr, half, dm.tw_fwd.as<Fr>());
        LAUNCH(D, k_fr_pow_table, cdiv(cdiv(half, 16), 128), 128, 0, st, c + 1, (const Fr *)nullptr, half, dm.tw_inv.as<Fr>());
    }
    if (new_g) {
        dm.coset_pre.ensure(n * sizeof(Fr));
        dm.coset_post.ensure(n * sizeof(Fr));
        LAUNCH(D, k_fr_pow_table, cdiv(cdiv(n, 16), 128), 128, 0, st, c + 3, (const Fr *)nullptr, n, dm.coset_pre.as<Fr>());
        LAUNCH(D, k_fr_pow_table, cdiv(cdiv(n, 16), 128), 128, 0, st, c + 4, c + 2, n, dm.coset_post.as<Fr>());
        memcpy(dm.g, g, 32);
        dm.has_g = true;
    }
    return dm;
}

// mode 0 FFT, 1 iFFT, 2 cosetFFT(g), 3 icosetFFT(g)  (basic_radix2_domain.tcc); 4: the unscaled inverse
// transform _basic_radix2_FFT(a, omega^-1) that iFFT and the extended / step domains build on.
// a: host vector transformed in place, or d_a: device vector transformed in place on `stream`.
int fr_fft(uint64_t *a, void *d_a, size_t log_n, int mode, const uint64_t *g, void *stream)
{
    if (!g_init) return fail(B200_ERR_NOT_INIT, "b200_init has not been called (no CUDA device => no result: there is no CPU fallback)");
    if ((!a && !d_a) || mode < 0 || mode > 4) return fail(B200_ERR_ARG, "bad argument");
    if (log_n < 1 || log_n > FR_TWO_ADICITY) return fail(B200_ERR_ARG, "basic_radix2: expected 1 <= log2(m) <= Fr::s = 28");
    if ((mode == 2 || mode == 3) && !g) return fail(B200_ERR_ARG, "coset transforms need the shift g");
    try {
        Device &D = g_devs[0];
        D.launches = 0;
        CK(cudaSetDevice(D.id));
        cudaStream_t st = (d_a && stream) ? (cudaStream_t)stream : D.stream;
        const uint32_t logn = (uint32_t)log_n;
        const size_t n = (size_t)1 << logn;
        engine_enter(D, st);  // tables / D.fr_b may still be in use by an earlier asynchronous call on another stream
        FrDomain &dm = domain_for(D, 0, logn, (mode == 2 || mode == 3) ? g : nullptr, st);
        Fr *data;
        D.fr_b.ensure(n * sizeof(Fr));
        if (d_a) {
            data = reinterpret_cast<Fr *>(d_a);
        } else {
            D.fr_a.ensure(n * sizeof(Fr));
            h2d(D, D.fr_a.p, a, n * sizeof(Fr), st);
            data = D.fr_a.as<Fr>();
        }
        Fr *work = D.fr_b.as<Fr>();
        const bool inverse = mode == 1 || mode == 3 || mode == 4;
        const Fr *tw = inverse ? dm.tw_inv.as<Fr>() : dm.tw_fwd.as<Fr>();
        const Fr *pre = mode == 2 ? dm.coset_pre.as<Fr>() : nullptr;
        const Fr *post = mode == 3 ? dm.coset_post.as<Fr>() : nullptr;
        const Fr *post_scalar = mode == 1 ? dm.consts.as<Fr>() + 2 : nullptr;
        // pass plan: the first pass does up to 10 stages on contiguous tiles; the remaining stages are
        // split evenly over passes of at most 8 stages, each block taking 2^(10-k) low-index neighbours
        std::vector<FftPass> plan;
        const uint32_t k1 = std::min<uint32_t>(FFT_TILE_LOG, logn);
        plan.push_back(FftPass{logn, 0, k1, 0});
        uint32_t rem = logn - k1;
        if (rem) {
            const uint32_t np = (rem + 7) / 8;
            uint32_t s0 = k1;
            for (uint32_t p = 0; p < np; p++) {
                const uint32_t k = (rem + (np - p) - 1) / (np - p);
                plan.push_back(FftPass{logn, s0, k, (uint32_t)FFT_TILE_LOG - k});
                s0 += k;
                rem -= k;
            }
        }
        // data -> work (bit reversal cannot run in place), middle passes in place on work, the last
        // pass writes back to data; a one-pass transform is copied back
        for (size_t pi = 0; pi < plan.size(); pi++) {
            const FftPass &P = plan[pi];
            const bool last = pi + 1 == plan.size();
            const Fr *src = pi == 0 ? data : work;
            Fr *dst = (last && pi > 0) ? data : work;
            const uint32_t E = 1u << (P.k + P.t);
            LAUNCH(D, k_fr_fft_pass, (uint32_t)(n / E), FFT_THREADS, (size_t)8 * E * 4, st, src, dst, P, tw, pi == 0 ? pre : (const Fr *)nullptr,
                   last ? post : (const Fr *)nullptr, last ? post_scalar : (const Fr *)nullptr);
        }
        if (plan.size() == 1) CK(cudaMemcpyAsync(data, work, n * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
        if (!d_a) {
            d2h(D, a, data, n * sizeof(Fr), st);
            CK(cudaStreamSynchronize(st));
        } else {
            engine_leave(D, st);
        }
        g_stats = b200_stats_t{};
        g_stats.n = n;
        g_stats.kernel_launches = D.launches;
        g_stats.num_windows = (uint32_t)plan.size();
        g_stats.h2d_bytes = d_a ? 0.0 : (double)n * sizeof(Fr);
        g_stats.d2h_bytes = d_a ? 0.0 : (double)n * sizeof(Fr);
        return B200_OK;
    } catch (const CudaError &e) {
        return fail(B200_ERR_CUDA, "%s", e.msg.c_str());
    } catch (const std::exception &ex) {  // nothing may unwind through the extern "C" boundary
        return fail(B200_ERR_CUDA, "host error: %s", ex.what());
    } catch (...) {
        return fail(B200_ERR_CUDA, "unknown host error");
    }
}

// ------------------------------------------------------------------------------
// step_radix2_domain transforms (2^log_big + 2^log_small points): mode 0 FFT, 1 iFFT, 2 cosetFFT(g), 3 icosetFFT(g)
// (step_radix2_domain.tcc:38-152).  The two radix-2 transforms inside are fr_fft on device buffers; the loops around
// them are the k_step_* kernels.  One upload, one download.
// ------------------------------------------------------------------------------
static DevBuf g_st_a, g_st_c, g_st_d, g_st_e, g_st_part, g_st_gp, g_st_const, g_st_half;
void fr_qap_release();
void fr_step_release()
{
    DevBuf *all[] = {&g_st_a, &g_st_c, &g_st_d, &g_st_e, &g_st_part, &g_st_gp, &g_st_const, &g_st_half};
    for (DevBuf *b : all) b->release();
    fr_qap_release();
}

// 1/2 in Fr, Montgomery form ((r + 1) / 2 * 2^256 mod r): FieldT(2).inverse() of step_radix2_domain.tcc:127
static const uint32_t FR_HALF[8] = {0x1ffffffeu, 0x783c14d8u, 0x0c8d1eddu, 0xaf982f6fu, 0xfcfd4f45u, 0x8f5f7492u, 0x3d9cbfacu, 0x1f37631au};

// table[i] = g^i (inverse == false) or g^-i, i < m, from the host element g; cst: 8 Fr of device scratch
static void coset_table_device(Device &D, cudaStream_t st, const uint64_t *g, bool inverse, size_t m, Fr *cst, Fr *table)
{
    CK(cudaMemcpyAsync(cst + 6, g, sizeof(Fr), cudaMemcpyHostToDevice, st));
    LAUNCH(D, k_fr_domain_consts, 1, 32, 0, st, 1u, (const Fr *)(cst + 6), cst);  // cst[3] = g, cst[4] = g^-1
    LAUNCH(D, k_fr_pow_table, cdiv(cdiv(m, 16), 128), 128, 0, st, (const Fr *)(cst + (inverse ? 4 : 3)), (const Fr *)nullptr, m, table);
}

// One step_radix2_domain transform of the m = 2^log_big + 2^log_small values at A (device, in place) on `st`:
// inverse == false: FFT (gp != nullptr: cosetFFT, gp[i] = g^i); inverse == true: iFFT (gp != nullptr: icosetFFT, gp[i] = g^-i).
static int step_transform_device(Device &D, cudaStream_t st, Fr *A, size_t log_big, size_t log_small, bool inverse, const Fr *gp)
{
    const size_t big = (size_t)1 << log_big, small = (size_t)1 << log_small, compr = big / small;
    g_st_c.ensure(big * sizeof(Fr));
    g_st_d.ensure(big * sizeof(Fr));
    g_st_e.ensure(small * sizeof(Fr));
    g_st_half.ensure(sizeof(Fr));
    const uint32_t J = (uint32_t)std::max<size_t>(1, std::min<size_t>(compr, std::max<size_t>(1, ((size_t)1 << 16) / small)));
    g_st_part.ensure((size_t)J * small * sizeof(Fr));
    // omega^i / omega^-i, i < big: the twiddle tables of the radix-2 domain of size 2 * big (omega = its root of unity, :31)
    FrDomain &dm2 = domain_for(D, 0, (uint32_t)log_big + 1, nullptr, st);
    const Fr *ow = dm2.tw_fwd.as<Fr>(), *owi = dm2.tw_inv.as<Fr>();
    Fr *C = g_st_c.as<Fr>(), *Dd = g_st_d.as<Fr>(), *E = g_st_e.as<Fr>(), *part = g_st_part.as<Fr>();
    if (!inverse) {
        LAUNCH(D, k_step_fwd_pre, cdiv(big, 256), 256, 0, st, (const Fr *)A, gp, ow, big, small, C, Dd);
        LAUNCH(D, k_step_strided_partial, cdiv(small * J, 256), 256, 0, st, (const Fr *)Dd, (const Fr *)nullptr, small, compr, 0u, J, part);
        LAUNCH(D, k_step_strided_final, cdiv(small, 256), 256, 0, st, (const Fr *)part, small, J, E);
        int rc = fr_fft(nullptr, C, log_big, 0, nullptr, st);
        if (rc == B200_OK && log_small >= 1) rc = fr_fft(nullptr, E, log_small, 0, nullptr, st);
        if (rc != B200_OK) return rc;
        CK(cudaMemcpyAsync(A, C, big * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(A + big, E, small * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
    } else {
        int rc = fr_fft(nullptr, A, log_big, 1, nullptr, st);
        if (rc == B200_OK && log_small >= 1) rc = fr_fft(nullptr, A + big, log_small, 1, nullptr, st);
        if (rc != B200_OK) return rc;
        CK(cudaMemcpyAsync(g_st_half.p, FR_HALF, sizeof(Fr), cudaMemcpyHostToDevice, st));
        LAUNCH(D, k_step_strided_partial, cdiv(small * J, 256), 256, 0, st, (const Fr *)A, ow, small, compr, 1u, J, part);
        LAUNCH(D, k_step_strided_final, cdiv(small, 256), 256, 0, st, (const Fr *)part, small, J, E);
        LAUNCH(D, k_step_inv_post, cdiv(big, 256), 256, 0, st, (const Fr *)A, (const Fr *)(A + big), (const Fr *)E, owi, gp,
               (const Fr *)g_st_half.as<Fr>(), big, small, A);
    }
    return B200_OK;
}

int fr_step_fft(uint64_t *a, size_t log_big, size_t log_small, int mode, const uint64_t *g)
{
    if (!g_init) return fail(B200_ERR_NOT_INIT, "b200_init has not been called (no CUDA device => no result: there is no CPU fallback)");
    if (!a || mode < 0 || mode > 3) return fail(B200_ERR_ARG, "bad argument");
    if (log_small >= log_big || log_big + 1 > FR_TWO_ADICITY) return fail(B200_ERR_ARG, "step_radix2: expected small_m < big_m and 2 big_m | 2^28");
    if (mode >= 2 && !g) return fail(B200_ERR_ARG, "coset transforms need the shift g");
    try {
        Device &D = g_devs[0];
        CK(cudaSetDevice(D.id));
        cudaStream_t st = D.stream;
        const size_t m = ((size_t)1 << log_big) + ((size_t)1 << log_small);
        const bool inverse = mode == 1 || mode == 3, coset = mode >= 2;
        g_st_a.ensure(m * sizeof(Fr));
        g_st_const.ensure(8 * sizeof(Fr));
        h2d(D, g_st_a.p, a, m * sizeof(Fr), st);
        const Fr *gp = nullptr;
        if (coset) {  // g^i (cosetFFT) or g^-i (icosetFFT), i < m
            g_st_gp.ensure(m * sizeof(Fr));
            coset_table_device(D, st, g, inverse, m, g_st_const.as<Fr>(), g_st_gp.as<Fr>());
            gp = g_st_gp.as<Fr>();
        }
        const int rc = step_transform_device(D, st, g_st_a.as<Fr>(), log_big, log_small, inverse, gp);
        if (rc != B200_OK) return rc;
        d2h(D, a, g_st_a.p, m * sizeof(Fr), st);
        CK(cudaStreamSynchronize(st));
        g_stats = b200_stats_t{};
        g_stats.n = m;
        g_stats.h2d_bytes = g_stats.d2h_bytes = (double)m * sizeof(Fr);
        return B200_OK;
    } catch (const CudaError &e) {
        return fail(B200_ERR_CUDA, "%s", e.msg.c_str());
    } catch (const std::exception &ex) {  // nothing may unwind through the extern "C" boundary
        return fail(B200_ERR_CUDA, "host error: %s", ex.what());
    } catch (...) {
        return fail(B200_ERR_CUDA, "unknown host error");
    }
}

// ------------------------------------------------------------------------------
// The vector part of r1cs_to_qap_witness_map (SNK/reductions/r1cs_to_qap/r1cs_to_qap.tcc:232-311) for d1 = d2 = d3 = 0
// (what every Groth16 / LegoGroth prover passes): from the evaluations aA, aB, aC of the constraint polynomials on the
// domain S to the coefficients of H = (A B - C) / Z,
//     A, B, C <- iFFT;  A, B, C <- cosetFFT(g);  T = A B - C;  T <- T / Z on the coset;  H <- icosetFFT(T, g).
// Seven transforms with nothing leaving the device in between (the reference runs them one by one on host vectors; through
// the per-transform shims each of them was an upload and a download).  log_small == QAP_BASIC: basic radix-2 domain of
// 2^log_big points, div = {Z^-1}; otherwise the step domain of 2^log_big + 2^log_small points, div = {c1, ratio, c0, Z1^-1}
// (the constants of step_radix2_domain::divide_by_Z_on_coset, formed by the caller).
// ------------------------------------------------------------------------------
static DevBuf g_q[3], g_q_gp, g_q_gip;
void fr_qap_release()
{
    for (auto &b : g_q) b.release();
    g_q_gp.release();
    g_q_gip.release();
}

int fr_qap_h(const uint64_t *aA, const uint64_t *aB, const uint64_t *aC, size_t log_big, size_t log_small, const uint64_t *g,
             const uint64_t *div, uint64_t *H)
{
    if (!g_init) return fail(B200_ERR_NOT_INIT, "b200_init has not been called (no CUDA device => no result: there is no CPU fallback)");
    if (!aA || !aB || !aC || !g || !div || !H) return fail(B200_ERR_ARG, "null argument");
    const bool basic = log_small == (size_t)-1;
    if (log_big < 1 || log_big + (basic ? 0 : 1) > FR_TWO_ADICITY || (!basic && log_small >= log_big)) return fail(B200_ERR_ARG, "bad domain");
    try {
        Device &D = g_devs[0];
        CK(cudaSetDevice(D.id));
        D.launches = 0;
        cudaStream_t st = D.stream;
        const size_t big = (size_t)1 << log_big, small = basic ? 0 : (size_t)1 << log_small, m = big + small;
        const uint64_t *src[3] = {aA, aB, aC};
        for (int k = 0; k < 3; k++) {
            g_q[k].ensure(m * sizeof(Fr));
            h2d(D, g_q[k].p, src[k], m * sizeof(Fr), st);
        }
        g_st_const.ensure(16 * sizeof(Fr));
        Fr *cst = g_st_const.as<Fr>();
        Fr *X = g_q[0].as<Fr>(), *Y = g_q[1].as<Fr>(), *Z = g_q[2].as<Fr>();
        if (basic) {
            for (Fr *v : {X, Y, Z}) {
                int rc = fr_fft(nullptr, v, log_big, 1, nullptr, st);
                if (rc == B200_OK) rc = fr_fft(nullptr, v, log_big, 2, g, st);
                if (rc != B200_OK) return rc;
            }
        } else {
            g_q_gp.ensure(m * sizeof(Fr));
            g_q_gip.ensure(m * sizeof(Fr));
            coset_table_device(D, st, g, false, m, cst, g_q_gp.as<Fr>());
            coset_table_device(D, st, g, true, m, cst, g_q_gip.as<Fr>());
            for (Fr *v : {X, Y, Z}) {
                int rc = step_transform_device(D, st, v, log_big, log_small, true, nullptr);
                if (rc == B200_OK) rc = step_transform_device(D, st, v, log_big, log_small, false, g_q_gp.as<Fr>());
                if (rc != B200_OK) return rc;
            }
        }
        LAUNCH(D, k_fr_mul_sub, cdiv(m, 256), 256, 0, st, X, (const Fr *)Y, (const Fr *)Z, m);
        CK(cudaMemcpyAsync(cst + 8, div, (basic ? 1 : 4) * sizeof(Fr), cudaMemcpyHostToDevice, st));
        if (basic) {
            LAUNCH(D, k_fr_scale, cdiv(m, 256), 256, 0, st, X, m, (const Fr *)(cst + 8));
            const int rc = fr_fft(nullptr, X, log_big, 3, g, st);
            if (rc != B200_OK) return rc;
        } else {
            LAUNCH(D, k_fr_scale_inv_geometric, cdiv(cdiv(big, INVG_RUN), 128), 128, 0, st, X, big, (const Fr *)(cst + 8));
            LAUNCH(D, k_fr_scale, cdiv(small, 256), 256, 0, st, X + big, small, (const Fr *)(cst + 11));
            const int rc = step_transform_device(D, st, X, log_big, log_small, true, g_q_gip.as<Fr>());
            if (rc != B200_OK) return rc;
        }
        d2h(D, H, X, m * sizeof(Fr), st);
        CK(cudaStreamSynchronize(st));
        g_stats = b200_stats_t{};
        g_stats.n = m;
        g_stats.h2d_bytes = 3.0 * m * sizeof(Fr);
        g_stats.d2h_bytes = (double)m * sizeof(Fr);
        return B200_OK;
    } catch (const CudaError &e) {
        return fail(B200_ERR_CUDA, "%s", e.msg.c_str());
    } catch (const std::exception &ex) {  // nothing may unwind through the extern "C" boundary
        return fail(B200_ERR_CUDA, "host error: %s", ex.what());
    } catch (...) {
        return fail(B200_ERR_CUDA, "unknown host error");
    }
}

}  // namespace eng
}  // namespace b200
