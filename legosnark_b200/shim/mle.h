// Shadow of LegoSNARK's prototools/mle.h (LS = /root/reference/src): with `legosnark_b200/shim` ahead of LS/prototools on
// the include path, every translation unit that says `#include "mle.h"` (LS/gadgets/sumcheck.h:8 and through it the
// sum-check and matrix gadgets) gets the reference's header as it is, except that the name DPMatrixMle denotes the class
// below.  DPMatrixMle's constructor (mle.h:241-259) is the O(n^2) preprocessing of the matrix sum-check,
//     v[r] = sum_l A[(l << d) + r] * eqTbl[l],   eqTbl = DPBeta::compute_eq_tbl(d, rho)   (mle.h:93-105)
// a serial double loop on the host; here it is one b200_fr_matrix_mle call (the eq table is built level by level on the
// device exactly as compute_eq_tbl writes it).  Everything else (DPMle's tables, getMLEPoly, pushRandomness) is inherited
// unchanged from the reference's DPMle, so callers see the same object.
// The reference's own class stays available as DPMatrixMle_cpu.
#ifndef B200_SHIM_MLE_H_
#define B200_SHIM_MLE_H_

#define DPMatrixMle DPMatrixMle_cpu
#include_next "mle.h"
#undef DPMatrixMle

#include <stdexcept>
#include <string>

#include "b200_libff.hpp"

class DPMatrixMle : public DPMle
{
public:
  // A is a vectorized matrix of size _n x _n, with _n = 2^_d
  DPMatrixMle(size_t _d, uint64 _n, const Ins &_A, const Ins &rho) : DPMle(_d, _n)
  {
    static_assert(sizeof(In) == 32, "unexpected scalar layout");
    if (_A.size() < (size_t)_n * _n || rho.size() < _d || _n != ((uint64)1 << _d)) throw std::runtime_error("DPMatrixMle: expected a 2^d x 2^d matrix and d challenges");
    b200shim::ensure_init();
    b200shim::check(b200_fr_matrix_mle(reinterpret_cast<const uint64_t *>(_A.data()), reinterpret_cast<const uint64_t *>(rho.data()), _d,
                                       reinterpret_cast<uint64_t *>(v.data())),
                    "b200_fr_matrix_mle");
    curVTable = v;  // the reference keeps both in step: v[r] = curVTable[r] = curVTable[r] + inc (mle.h:254)
  }
};

#endif  // B200_SHIM_MLE_H_
