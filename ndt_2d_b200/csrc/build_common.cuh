// build_common.cuh -- device helpers shared by build.cu and probe.cu.
#ifndef NDT2D_BUILD_COMMON_CUH_
#define NDT2D_BUILD_COMMON_CUH_

#include <cuda_runtime.h>

namespace ndt2d_dev
{

// RN(a / b) for a point count b (an integer in [1, 2^20]) from r = RN(1 / b): q0 = RN(a r) is
// within 2 ulps of a / b, the residual a - b q0 is then a multiple of ulp(q0) below 2^22 ulps and
// the FMA forms it exactly, and q0 + rem r differs from a / b by less than 2^-50 ulp -- while
// a / b (b not a power of two: a then has more trailing zeros than a tie would need; b a power
// of two: the quotient is exact) stays at least 2^-21 ulp away from every rounding boundary.
// So the result is the correctly rounded quotient, i.e. __ddiv_rn(a, b) bit for bit, with 3
// dependent operations instead of the divide's ~10; r does not depend on the running value, so
// it is computed off the recurrence's critical path.  (Values near the subnormal range, where
// the residual could lose bits, take the divide.)
__device__ __forceinline__ double div_by_count(double a, double b, double r)
{
  if (fabs(a) < 1.0e-280) {return __ddiv_rn(a, b);}
  const double q0 = __dmul_rn(a, r);
  const double rem = __fma_rn(-b, q0, a);
  return __fma_rn(rem, r, q0);
}

}  // namespace ndt2d_dev

#endif  // NDT2D_BUILD_COMMON_CUH_
