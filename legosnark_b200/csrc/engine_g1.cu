// engine_g1.cu — G1 (coordinates in Fq) instantiation of the engine, plus the
// prime-field / Fq2 element-wise parity hooks.
#include "engine_impl.cuh"

namespace b200 {
namespace eng {

B200_INSTANTIATE_GROUP(Fq)

int test_field_op(int field, int op, const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out)
{
    if (op < 0 || op > 8 || (field == 2 && (op == 6 || op == 7))) return fail(B200_ERR_ARG, "bad op");
    switch (field) {
    case 0: return run_elementwise<Fq>(a, b, n, out, PrimeOp<Fq>{op});
    case 1: return run_elementwise<Fr>(a, b, n, out, PrimeOp<Fr>{op});
    case 2: return run_elementwise<Fq2>(a, b, n, out, Fq2Op{op});
    default: return fail(B200_ERR_ARG, "bad field");
    }
}

// CPPoly::prove (LS/gadgets/poly.h:45-91) on one device: the witness coefficients are folded on
// the device (fr_kernels.cuh) and never leave it; witness[i] = multiExpMA(g1s, w_i) is an MSM over
// the first 2^(d-i-1) bases of the resident key with device-resident scalars.
int cppoly_prove_g1(uint64_t key, const uint64_t *v, const uint64_t *r, size_t d, uint64_t *witness, uint64_t *eval)
{
    if (!g_init) return fail(B200_ERR_NOT_INIT, "b200_init has not been called (no CUDA device => no result: there is no CPU fallback)");
    if (!v || !witness || (d && !r)) return fail(B200_ERR_ARG, "null argument");
    if (d > 30) return fail(B200_ERR_ARG, "d out of range");
    auto it = g_pinned.find(key);
    if (it == g_pinned.end() || it->second->group != 0) return fail(B200_ERR_ARG, "unknown G1 bases handle");
    PinnedBases &pb = *it->second;
    if (pb.shards.size() != 1) return fail(B200_ERR_ARG, "cppoly_prove needs a single-device key");
    const size_t N = (size_t)1 << d;
    if (d && pb.n < N / 2) return fail(B200_ERR_ARG, "key shorter than 2^(d-1) bases");
    try {
        Device &D = g_devs[pb.shards[0].dev];
        CK(cudaSetDevice(D.id));
        D.launches = 0;
        const void *fin = fr_fold_device(D, v, r, d, true);
        if (eval) CK(cudaMemcpyAsync(eval, fin, 32, cudaMemcpyDeviceToHost, D.stream));
        const uint32_t fold_launches = D.launches;
        uint32_t launches = fold_launches;
        double dev_ms = 0;
        const char *w = reinterpret_cast<const char *>(D.fr_w.p);
        for (size_t i = 0; i < d; i++) {
            const size_t m = (size_t)1 << (d - i - 1), start = N - (2 * m);
            const int rc = msm_pinned<Fq>(key, 0, nullptr, w + start * 32, m, nullptr, witness + 12 * i);
            if (rc != B200_OK) return rc;
            launches += g_stats.kernel_launches;
            dev_ms += g_stats.device_ms;
        }
        CK(cudaStreamSynchronize(D.stream));
        g_stats.n = N;
        g_stats.kernel_launches = launches;
        g_stats.device_ms = dev_ms;
        g_stats.h2d_bytes = (double)(N + d) * 32;
        g_stats.d2h_bytes = (double)d * sizeof(XYZZ<Fq>) + 32;
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
