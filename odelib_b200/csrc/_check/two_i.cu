#define ODL_N 4
#define ODL_P 5
#define ODL_NOUT 2
#define ODL_RHS_FLOPS 11
#define ODL_AUTONOMOUS 1
template <class YV, class PV, class DV>
__device__ __forceinline__ void odl_rhs(const YV& y, const double t, const PV& p, DV& dy) {
  const double v10 = p[0] * y[0];
  const double v11 = p[1] * y[0];
  const double v12 = v11 * y[3];
  const double v13 = v10 - v12;
  const double v14 = p[4] * y[1];
  const double v15 = v12 - v14;
  const double v16 = p[3] * y[2];
  const double v17 = v14 - v16;
  const double v18 = p[2] * p[3];
  const double v19 = v18 * y[2];
  const double v20 = v19 - v12;
  dy[0] = v13;
  dy[1] = v15;
  dy[2] = v17;
  dy[3] = v20;
}
template <class YV, class PV>
__device__ __forceinline__ void odl_jac(const YV& y, const double t, const PV& p, double (&J)[ODL_N][ODL_N]) {
  const double v11 = p[1] * y[0];
  const double v18 = p[2] * p[3];
  const double v23 = p[1] * y[3];
  const double v24 = p[0] - v23;
  const double v25 = -(v23);
  const double v26 = -(p[4]);
  const double v27 = -(p[3]);
  const double v28 = -(v11);
  J[0][0] = v24;
  J[0][1] = 0.0;
  J[0][2] = 0.0;
  J[0][3] = v28;
  J[1][0] = v23;
  J[1][1] = v26;
  J[1][2] = 0.0;
  J[1][3] = v11;
  J[2][0] = 0.0;
  J[2][1] = p[4];
  J[2][2] = v27;
  J[2][3] = 0.0;
  J[3][0] = v25;
  J[3][1] = 0.0;
  J[3][2] = v18;
  J[3][3] = v28;
}
template <class YV, class PV, class DV>
__device__ __forceinline__ void odl_dfdt(const YV& y, const double t, const PV& p, DV& ft) {
  ft[0] = 0.0;
  ft[1] = 0.0;
  ft[2] = 0.0;
  ft[3] = 0.0;
}
template <class YV>
__device__ __forceinline__ void odl_observe(const YV& y, double (&out)[ODL_NOUT]) {
  out[0] = __dadd_rn(__dadd_rn(y[0], y[1]), y[2]);
  out[1] = y[3];
}
#include "odl_kernels.cuh"
