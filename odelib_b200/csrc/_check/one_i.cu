#define ODL_N 3
#define ODL_P 4
#define ODL_NOUT 2
#define ODL_RHS_FLOPS 9
#define ODL_AUTONOMOUS 1
template <class YV, class PV, class DV>
__device__ __forceinline__ void odl_rhs(const YV& y, const double t, const PV& p, DV& dy) {
  const double v8 = p[0] * y[0];
  const double v9 = p[1] * y[0];
  const double v10 = v9 * y[2];
  const double v11 = v8 - v10;
  const double v12 = p[3] * y[1];
  const double v13 = v10 - v12;
  const double v14 = p[2] * p[3];
  const double v15 = v14 * y[1];
  const double v16 = v15 - v10;
  dy[0] = v11;
  dy[1] = v13;
  dy[2] = v16;
}
template <class YV, class PV>
__device__ __forceinline__ void odl_jac(const YV& y, const double t, const PV& p, double (&J)[ODL_N][ODL_N]) {
  const double v9 = p[1] * y[0];
  const double v14 = p[2] * p[3];
  const double v19 = p[1] * y[2];
  const double v20 = p[0] - v19;
  const double v21 = -(v19);
  const double v22 = -(p[3]);
  const double v23 = -(v9);
  J[0][0] = v20;
  J[0][1] = 0.0;
  J[0][2] = v23;
  J[1][0] = v19;
  J[1][1] = v22;
  J[1][2] = v9;
  J[2][0] = v21;
  J[2][1] = v14;
  J[2][2] = v23;
}
template <class YV, class PV, class DV>
__device__ __forceinline__ void odl_dfdt(const YV& y, const double t, const PV& p, DV& ft) {
  ft[0] = 0.0;
  ft[1] = 0.0;
  ft[2] = 0.0;
}
template <class YV>
__device__ __forceinline__ void odl_observe(const YV& y, double (&out)[ODL_NOUT]) {
  out[0] = __dadd_rn(y[0], y[1]);
  out[1] = y[2];
}
#include "odl_kernels.cuh"
