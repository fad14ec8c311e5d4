// optim.cuh -- the Adam element update shared by the multi-tensor kernel (optim.cu) and the data-parallel
// peer-memory kernel (dp_p2p.cu).  A = any argument struct with the hyper-parameter fields used below.
#pragma once
#include "common.cuh"

namespace sk {

// soket/optim.pyx:201-269, every operation separately rounded in the reference's order
template <typename A>
__device__ __forceinline__ void adam_one(float &p, float g, float &m, float &v, const A &a,
                                         const float bc1, const float bc2) {
  if (a.have_scale) g = __fmul_rn(g, a.grad_scale);
  if (a.have_wd) g = __fadd_rn(g, __fmul_rn(p, a.wd));  // optim.pyx:220-222
  const float gm = __fmul_rn(g, a.omb1);                 // grad * (1 - beta1)
  const float gv = __fmul_rn(g, __fmul_rn(g, a.omb2));   // grad * (grad * (1 - beta2))
  if (a.first) {  // optim.pyx:224-238: first step has no beta*state term
    m = gm;
    v = gv;
  } else {
    m = __fadd_rn(__fmul_rn(m, a.beta1), gm);
    v = __fadd_rn(__fmul_rn(v, a.beta2), gv);
  }
  const float mh = __fdiv_rn(m, bc1);  // optim.pyx:246-247
  const float vh = __fdiv_rn(v, bc2);
  // p - lr * (mh / (pow(vh, 0.5) + eps))   optim.pyx:254-263 ; quirk Q3: maximize is a no-op
  p = __fsub_rn(p, __fmul_rn(a.lr, __fdiv_rn(mh, __fadd_rn(__fsqrt_rn(vh), a.eps))));
}

}  // namespace sk
