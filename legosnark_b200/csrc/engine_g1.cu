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

}  // namespace eng
}  // namespace b200
