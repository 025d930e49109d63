#define ODL_N 2
#define ODL_P 3
#define ODL_NOUT 2
#define ODL_RHS_FLOPS 8
#define ODL_AUTONOMOUS 1
template <class YV, class PV, class DV>
__device__ __forceinline__ void odl_rhs(const YV& y, const double t, const PV& p, DV& dy) {
  const double v6 = p[0] * y[0];
  const double v7 = p[1] * y[0];
  const double v8 = v7 * y[1];
  const double v9 = v6 - v8;
  const double v10 = p[2] * p[1];
  const double v11 = v10 * y[0];
  const double v12 = v11 * y[1];
  const double v13 = v12 - v8;
  dy[0] = v9;
  dy[1] = v13;
}
template <class YV, class PV>
__device__ __forceinline__ void odl_jac(const YV& y, const double t, const PV& p, double (&J)[ODL_N][ODL_N]) {
  const double v7 = p[1] * y[0];
  const double v10 = p[2] * p[1];
  const double v11 = v10 * y[0];
  const double v16 = p[1] * y[1];
  const double v17 = p[0] - v16;
  const double v18 = v10 * y[1];
  const double v19 = v18 - v16;
  const double v20 = -(v7);
  const double v21 = v11 - v7;
  J[0][0] = v17;
  J[0][1] = v20;
  J[1][0] = v19;
  J[1][1] = v21;
}
template <class YV, class PV, class DV>
__device__ __forceinline__ void odl_dfdt(const YV& y, const double t, const PV& p, DV& ft) {
  ft[0] = 0.0;
  ft[1] = 0.0;
}
template <class YV>
__device__ __forceinline__ void odl_observe(const YV& y, double (&out)[ODL_NOUT]) {
  out[0] = y[0];
  out[1] = y[1];
}
#include "odl_kernels.cuh"
