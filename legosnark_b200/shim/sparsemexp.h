// Shadow of LegoSNARK's utils/sparsemexp.h (LS = /root/reference/src): with `legosnark_b200/shim` ahead of
// LS/utils on the include path, LS/gadgets/subspace.cc picks this file up through its own
// `#include "sparsemexp.h"` (subspace.cc:5) and no reference file is edited.
//
// What it binds: SubspaceSnark::keygen (subspace.cc:37-76) builds its CRS row P with
//     mtxmultiexp(key->P, k, rel->M)                      subspace.cc:55
// a file-local function that runs ONE tiny multi_exp per matrix column,
//     for (const ColG1 &c : m) out[i++] = simplesparsemexp(c, exps);      subspace.cc:18-25
//     -> sparsemexpG -> multi_exp<LG1, LFr, BDLO12>                        sparsemexp.h:62-90
// 2 050 columns of one or two terms for the shipped cplink example; through the libff-level shim each of them
// would be a host<->device round trip.  The overload below takes the exponent vector by NON-const reference:
// keygen passes its local `vector<IScalar> k`, for which binding to `vector<LFr> &` is a better conversion
// than the reference's `const vector<LFr> &` ([over.ics.rank]), so the call site resolves here and all
// columns go to the engine in one b200_msm_batch_g1 call.  Callers that pass a const vector keep the
// reference's definition; the group elements are the same either way (zero coefficients contribute nothing
// and the generator is a point like any other, sparsemexp.h:75-86).
#ifndef B200_SHIM_SPARSEMEXP_H_
#define B200_SHIM_SPARSEMEXP_H_

#include_next "sparsemexp.h"

#include "b200_libff.hpp"

inline void mtxmultiexp(std::vector<LG1> &out, std::vector<LFr> &exps, const std::vector<ColG1> &m)
{
    std::vector<LG1> bases;
    std::vector<LFr> scalars;
    std::vector<uint64_t> offsets(1, 0);
    size_t terms = 0;
    for (const ColG1 &c : m) terms += c.size();
    bases.reserve(terms);
    scalars.reserve(terms);
    offsets.reserve(m.size() + 1);
    for (const ColG1 &c : m) {
        for (const auto &cp : c) {
            bases.push_back(cp.val);
            scalars.push_back(exps[cp.pos]);
        }
        offsets.push_back(bases.size());
    }
    out = b200shim::msm_batch<LG1, LFr>(bases, scalars, offsets);
}

#endif  // B200_SHIM_SPARSEMEXP_H_
