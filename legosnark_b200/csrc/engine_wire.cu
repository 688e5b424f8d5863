// engine_wire.cu — host side of point (de)compression (wire_kernels.cuh).
#include "engine_common.hpp"
#include "host_copy.hpp"
#include "msm_kernels.cuh"
#include "wire_kernels.cuh"

namespace b200 {
namespace eng {

template <class F>
static int compress_points(const uint64_t *pts, size_t n, int fl, uint64_t *x_out, uint8_t *flags)
{
    if (!g_init) return fail(B200_ERR_NOT_INIT, "b200_init has not been called (no CUDA device => no result: there is no CPU fallback)");
    if (fl < 0 || fl > 2) return fail(B200_ERR_ARG, "bad flavour");
    if (n == 0) return B200_OK;
    if (!pts || !x_out || !flags) return fail(B200_ERR_ARG, "null argument");
    try {
        Device &D = g_devs[0];
        D.launches = 0;
        CK(cudaSetDevice(D.id));
        cudaStream_t st = D.stream;
        D.bases_jac.ensure(n * sizeof(Jacobian<F>));
        D.bases_aff.ensure(n * sizeof(Affine<F>));
        D.flags.ensure(n);
        D.prefix.ensure(n * sizeof(F));
        D.out_norm.ensure(n * sizeof(F));
        D.out_jac.ensure(n);
        h2d(D, D.bases_jac.p, pts, n * sizeof(Jacobian<F>), st);
        // batch_to_special's one-inversion normalisation (k_ingest), then the per-point conversion
        const uint32_t blocks = std::max<uint32_t>(1, std::min<uint32_t>(cdiv(n, 128 * 16), (uint32_t)D.sms * 8));
        LAUNCH(D, (k_ingest<F, false>), blocks, 128, 0, st, (const Jacobian<F> *)D.bases_jac.as<Jacobian<F>>(), D.bases_aff.p,
               D.flags.as<uint8_t>(), D.prefix.as<F>(), n);
        LAUNCH(D, (k_compress<F>), cdiv(n, 128), 128, 0, st, (const Affine<F> *)D.bases_aff.as<Affine<F>>(), (const uint8_t *)D.flags.as<uint8_t>(),
               n, fl, D.out_norm.as<F>(), D.out_jac.as<uint8_t>());
        d2h(D, x_out, D.out_norm.p, n * sizeof(F), st);
        CK(cudaMemcpyAsync(flags, D.out_jac.p, n, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        g_stats = b200_stats_t{};
        g_stats.n = n;
        g_stats.kernel_launches = D.launches;
        g_stats.h2d_bytes = (double)n * sizeof(Jacobian<F>);
        g_stats.d2h_bytes = (double)n * (sizeof(F) + 1);
        return B200_OK;
    } catch (const CudaError &e) {
        return fail(B200_ERR_CUDA, "%s", e.msg.c_str());
    } catch (const std::exception &ex) {  // nothing may unwind through the extern "C" boundary
        return fail(B200_ERR_CUDA, "host error: %s", ex.what());
    } catch (...) {
        return fail(B200_ERR_CUDA, "unknown host error");
    }
}

template <class F>
static int decompress_points(const uint64_t *x, const uint8_t *flags, size_t n, int fl, uint64_t *pts_out, uint8_t *bad)
{
    if (!g_init) return fail(B200_ERR_NOT_INIT, "b200_init has not been called (no CUDA device => no result: there is no CPU fallback)");
    if (fl < 0 || fl > 2) return fail(B200_ERR_ARG, "bad flavour");
    if (n == 0) return B200_OK;
    if (!x || !flags || !pts_out) return fail(B200_ERR_ARG, "null argument");
    try {
        Device &D = g_devs[0];
        D.launches = 0;
        CK(cudaSetDevice(D.id));
        cudaStream_t st = D.stream;
        D.out_norm.ensure(n * sizeof(F));
        D.flags.ensure(n);
        D.out_jac.ensure(n * sizeof(Jacobian<F>));
        D.coeff.ensure(n);
        h2d(D, D.out_norm.p, x, n * sizeof(F), st);
        CK(cudaMemcpyAsync(D.flags.p, flags, n, cudaMemcpyHostToDevice, st));
        LAUNCH(D, (k_decompress<F>), cdiv(n, 128), 128, 0, st, (const F *)D.out_norm.as<F>(), (const uint8_t *)D.flags.as<uint8_t>(), n, fl,
               D.out_jac.as<Jacobian<F>>(), D.coeff.as<uint8_t>());
        d2h(D, pts_out, D.out_jac.p, n * sizeof(Jacobian<F>), st);
        std::vector<uint8_t> hb(n);
        CK(cudaMemcpyAsync(hb.data(), D.coeff.p, n, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        size_t nbad = 0;
        for (size_t i = 0; i < n; i++) nbad += hb[i];
        if (bad) memcpy(bad, hb.data(), n);
        g_stats = b200_stats_t{};
        g_stats.n = n;
        g_stats.kernel_launches = D.launches;
        g_stats.h2d_bytes = (double)n * (sizeof(F) + 1);
        g_stats.d2h_bytes = (double)n * (sizeof(Jacobian<F>) + 1);
        // the reference's sqrt does not terminate on a non-square (fp2.tcc:170); here the call reports it
        if (nbad && !bad) return fail(B200_ERR_ARG, "%zu of %zu compressed points are not on the curve (X^3 + b is not a square)", nbad, n);
        return B200_OK;
    } catch (const CudaError &e) {
        return fail(B200_ERR_CUDA, "%s", e.msg.c_str());
    } catch (const std::exception &ex) {  // nothing may unwind through the extern "C" boundary
        return fail(B200_ERR_CUDA, "host error: %s", ex.what());
    } catch (...) {
        return fail(B200_ERR_CUDA, "unknown host error");
    }
}

int compress_g1(const uint64_t *pts, size_t n, int fl, uint64_t *x, uint8_t *flags) { return compress_points<Fq>(pts, n, fl, x, flags); }
int compress_g2(const uint64_t *pts, size_t n, int fl, uint64_t *x, uint8_t *flags) { return compress_points<Fq2>(pts, n, fl, x, flags); }
int decompress_g1(const uint64_t *x, const uint8_t *flags, size_t n, int fl, uint64_t *pts, uint8_t *bad)
{
    return decompress_points<Fq>(x, flags, n, fl, pts, bad);
}
int decompress_g2(const uint64_t *x, const uint8_t *flags, size_t n, int fl, uint64_t *pts, uint8_t *bad)
{
    return decompress_points<Fq2>(x, flags, n, fl, pts, bad);
}

}  // namespace eng
}  // namespace b200
