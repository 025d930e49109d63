// Host harness (TEST INFRASTRUCTURE ONLY): compiles the device steppers of odl_kernels.cuh with g++ so the
// DOPRI5 / ROS23 / Radau5 step logic can be exercised against scipy on a GPU-less box.  Not part of the
// product: nothing in odelib_b200/ loads this; the kernels and warp-collective code are compiled out.
#include <math.h>
#include <stdint.h>
#include <string.h>

#define ODL_HOST_HARNESS 1
#define __device__
#define __forceinline__ inline
#define __global__
#define __launch_bounds__(...)
#define __constant__ static const
#define __log2f(x) log2f(x)
static inline float __frcp_rn(float x) { return 1.0f / x; }
#define __powf(a, b) powf((a), (b))
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
static inline unsigned int __umulhi(unsigned int a, unsigned int b) { return (unsigned int)(((unsigned long long)a * b) >> 32); }
static inline void sincospi(double x, double* s, double* c) { *s = sin(M_PI * x); *c = cos(M_PI * x); }

#include ODL_MODEL_HEADER
#include "odl_kernels.cuh"

struct HostSink {
  static constexpr bool kObservedOnly = false;
  double* out;
  inline void put(int, const double (&)[ODL_NOUT]) {}
  inline void operator()(int slot, const double (&yi)[ODL_N]) {
    for (int i = 0; i < ODL_N; ++i) out[(long long)slot * ODL_N + i] = yi[i];
  }
};

// LU self-test hooks: solve A x = b (real) and (Ar + i Ai) x = b (complex) with the steppers' in-register LU
extern "C" void harness_lu(const double* A, const double* b, double* x) {
  OdlLU F;
  for (int i = 0; i < ODL_N; ++i) for (int j = 0; j < ODL_N; ++j) F.a[i][j] = A[i * ODL_N + j];
  odl_lu_factor(F);
  double v[ODL_N];
  for (int i = 0; i < ODL_N; ++i) v[i] = b[i];
  odl_lu_solve(F, v);
  for (int i = 0; i < ODL_N; ++i) x[i] = v[i];
}
extern "C" void harness_clu(const double* Ar, const double* Ai, const double* br, const double* bi, double* xr, double* xi) {
  OdlCLU F;
  for (int i = 0; i < ODL_N; ++i) for (int j = 0; j < ODL_N; ++j) { F.ar[i][j] = Ar[i * ODL_N + j]; F.ai[i][j] = Ai[i * ODL_N + j]; }
  odl_clu_factor(F);
  double vr[ODL_N], vi[ODL_N];
  for (int i = 0; i < ODL_N; ++i) { vr[i] = br[i]; vi[i] = bi[i]; }
  odl_clu_solve(F, vr, vi);
  for (int i = 0; i < ODL_N; ++i) { xr[i] = vr[i]; xi[i] = vi[i]; }
}
extern "C" void harness_dbg(long long* out) { for (int i = 0; i < 8; ++i) { out[i] = odl_dbg[i]; odl_dbg[i] = 0; } }
extern "C" int harness_dims(int* n, int* p) { *n = ODL_N; *p = ODL_P; return 0; }

// solver: 0 DOPRI5, 1 ROS23, 3 Radau5, 4 BDF.  out: [n_slot][ODL_N].  Returns the status word.
extern "C" int harness_solve(int solver, const double* theta, const double* slot_t, int n_slot, const double* y0,
                             double t0, double rtol, double atol, int max_steps, double* out, int* nsteps) {
  int y0p[ODL_N];
  for (int i = 0; i < ODL_N; ++i) y0p[i] = -1;
  OdlData D;
  memset(&D, 0, sizeof D);
  D.slot_t = slot_t; D.y0 = y0; D.y0_from_param = y0p; D.n_slot = n_slot; D.t0 = t0;
  OdlOpts O;
  memset(&O, 0, sizeof O);
  O.rtol = rtol; O.atol = atol; O.max_steps = max_steps; O.stiff_min_steps = 2000;
  OdlShared S;
  memset(&S, 0, sizeof S);
  S.slot_t = const_cast<double*>(slot_t);
  double p[ODL_P];
  for (int q = 0; q < ODL_P; ++q) p[q] = theta[q];
  OdlStepper st;
  HostSink sink{out};
  odl_init_system(st, p, D, O, nullptr, false);
  odl_emit_initial_slots(st, S, D, sink);
  OdlRadauAux ax;
  ax.reset();
  OdlBdfAux bx;
  bx.reset();
  while (st.slot < D.n_slot && st.status == ODL_OK) {
    if (solver == 0) odl_dopri5_attempt(st, p, S, D, O, sink);
    else if (solver == 1) odl_ros23_attempt(st, p, S, D, O, sink);
    else if (solver == 4) odl_bdf_attempt(st, bx, p, S, D, O, sink);
    else odl_radau5_attempt(st, ax, p, S, D, O, sink);
  }
  *nsteps = st.nsteps;
  return st.status;
}

// DOPRI5 with the sink of the sweep / chain kernels' kind: observed columns only, interpolated in OBSERVED space (the
// kObservedOnly path of odl_dopri5_attempt).  out: [n_slot][ODL_NOUT].
struct HostObsSink {
  static constexpr bool kObservedOnly = true;
  double* out;
  inline void operator()(int slot, const double (&yi)[ODL_N]) {
    double o[ODL_NOUT];
    odl_observe(yi, o);
    put(slot, o);
  }
  inline void put(int slot, const double (&o)[ODL_NOUT]) {
    for (int c = 0; c < ODL_NOUT; ++c) out[(long long)slot * ODL_NOUT + c] = o[c];
  }
};
extern "C" int harness_nout(void) { return ODL_NOUT; }
extern "C" int harness_solve_observed(const double* theta, const double* slot_t, int n_slot, const double* y0, double t0,
                                      double rtol, double atol, int max_steps, double* out, int* nsteps) {
  int y0p[ODL_N];
  for (int i = 0; i < ODL_N; ++i) y0p[i] = -1;
  OdlData D;
  memset(&D, 0, sizeof D);
  D.slot_t = slot_t; D.y0 = y0; D.y0_from_param = y0p; D.n_slot = n_slot; D.t0 = t0;
  OdlOpts O;
  memset(&O, 0, sizeof O);
  O.rtol = rtol; O.atol = atol; O.max_steps = max_steps; O.stiff_min_steps = 2000;
  OdlShared S;
  memset(&S, 0, sizeof S);
  S.slot_t = const_cast<double*>(slot_t);
  double p[ODL_P];
  for (int q = 0; q < ODL_P; ++q) p[q] = theta[q];
  OdlStepper st;
  HostObsSink sink{out};
  odl_init_system(st, p, D, O, nullptr, false);
  odl_emit_initial_slots(st, S, D, sink);
  while (st.slot < D.n_slot && st.status == ODL_OK) odl_dopri5_attempt(st, p, S, D, O, sink);
  *nsteps = st.nsteps;
  return st.status;
}

// The AUTO sweep's hand-over on the host: DOPRI5 until it gives a row up (step cap / projection check), `idle` further
// attempts of the stopped solve (what a lane of the sweep kernel does while it waits for its warp's write-back: they must
// leave the state exactly as it is), then the BDF stepper continuing from that state.  out: [n_slot][ODL_N];
// info[0] = DOPRI5 attempts, info[1] = BDF steps, info[2] = slot at the hand-over; t_hand = time at the hand-over.
extern "C" int harness_solve_handover(const double* theta, const double* slot_t, int n_slot, const double* y0, double t0,
                                      double rtol, double atol, int cap, int early, int idle, double* out, int* info,
                                      double* t_hand) {
  int y0p[ODL_N];
  for (int i = 0; i < ODL_N; ++i) y0p[i] = -1;
  OdlData D;
  memset(&D, 0, sizeof D);
  D.slot_t = slot_t; D.y0 = y0; D.y0_from_param = y0p; D.n_slot = n_slot; D.t0 = t0;
  OdlOpts O;
  memset(&O, 0, sizeof O);
  O.rtol = rtol; O.atol = atol; O.max_steps = cap; O.early_check_steps = early; O.stiff_check = 1; O.stiff_min_steps = 2000;
  OdlShared S;
  memset(&S, 0, sizeof S);
  S.slot_t = const_cast<double*>(slot_t);
  double p[ODL_P];
  for (int q = 0; q < ODL_P; ++q) p[q] = theta[q];
  OdlStepper st;
  HostSink sink{out};
  odl_init_system(st, p, D, O, nullptr, false);
  odl_emit_initial_slots(st, S, D, sink);
  while (st.slot < D.n_slot && st.status == ODL_OK) odl_dopri5_attempt(st, p, S, D, O, sink);
  info[0] = st.nsteps; info[1] = 0; info[2] = st.slot; *t_hand = st.t;
  if (st.status != ODL_MAXSTEPS && st.status != ODL_STIFF) return st.status;       // finished (or failed otherwise)
  const double t_stop = st.t;
  double y_stop[ODL_N];
  for (int i = 0; i < ODL_N; ++i) y_stop[i] = st.y[i];
  const int slot_stop = st.slot;
  for (int k = 0; k < idle; ++k) odl_dopri5_attempt(st, p, S, D, O, sink);
  if (st.t != t_stop || st.slot != slot_stop || st.nsteps != info[0]) return -100;  // a stopped solve must stand still
  for (int i = 0; i < ODL_N; ++i) if (memcmp(&st.y[i], &y_stop[i], 8) != 0) return -101;
  // the stiff pass takes over (odl_take_over): t, y, slot; k1 = f(t, y); a fresh stepper of order 1
  OdlOpts Ob = O;
  Ob.max_steps = 2000000; Ob.early_check_steps = 0; Ob.stiff_check = 0;
  st.status = ODL_OK; st.nsteps = 0;
  odl_rhs(st.y, st.t, p, st.k1);
  OdlBdfAux bx;
  bx.reset();
  while (st.slot < D.n_slot && st.status == ODL_OK) odl_bdf_attempt(st, bx, p, S, D, Ob, sink);
  info[1] = st.nsteps;
  return st.status;
}

// DOPRI5 progress profile (dev tool: how well does early progress predict the total step count?):
// marks[k] = time reached after checkpoints[k] attempted steps (NaN if the solve ended earlier), hs[k] = step size then
extern "C" int harness_progress(const double* theta, const double* slot_t, int n_slot, const double* y0, double t0,
                                double rtol, double atol, int max_steps, const int* checkpoints, int n_check,
                                double* marks, double* hs, int* nsteps) {
  int y0p[ODL_N];
  for (int i = 0; i < ODL_N; ++i) y0p[i] = -1;
  OdlData D;
  memset(&D, 0, sizeof D);
  D.slot_t = slot_t; D.y0 = y0; D.y0_from_param = y0p; D.n_slot = n_slot; D.t0 = t0;
  OdlOpts O;
  memset(&O, 0, sizeof O);
  O.rtol = rtol; O.atol = atol; O.max_steps = max_steps; O.stiff_min_steps = 2000;
  OdlShared S;
  memset(&S, 0, sizeof S);
  S.slot_t = const_cast<double*>(slot_t);
  double p[ODL_P];
  for (int q = 0; q < ODL_P; ++q) p[q] = theta[q];
  static double scratch[4096];
  OdlStepper st;
  HostSink sink{scratch};
  odl_init_system(st, p, D, O, nullptr, false);
  odl_emit_initial_slots(st, S, D, sink);
  for (int k = 0; k < n_check; ++k) { marks[k] = NAN; hs[k] = NAN; }
  while (st.slot < D.n_slot && st.status == ODL_OK) {
    odl_dopri5_attempt(st, p, S, D, O, sink);
    for (int k = 0; k < n_check; ++k) if (st.nsteps == checkpoints[k]) { marks[k] = st.t; hs[k] = st.h; }
  }
  *nsteps = st.nsteps;
  return st.status;
}
