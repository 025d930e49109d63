// odl_kernels.cuh -- device code of the ODElib hot path for sm_100a (B200).
//
// Compiled per model by NVRTC at run time (odl_capi.cu prepends the traced model) and by nvcc at
// build time for the demo models (register/spill checks).  The including translation unit must
// define, before this file:
//     ODL_N, ODL_P, ODL_NOUT            state / parameter / observed-column counts
//     odl_rhs, odl_jac, odl_dfdt, odl_observe   (emitted by odelib_b200/tracer.py)
// Optional: ODL_BLOCK (threads per CTA), ODL_MINBLOCKS, ODL_DENSE (1 = dense output, 0 = land on slots).
//
// What replaces what (reference file:line -> here):
//   Framework.py:656   odeint(func, y0, times, args)       -> Dopri5 stepper, one system per thread
//   Framework.py:659-664, :677-682  summation + obs pick   -> odl_observe on the dense output at the
//                                                             19 distinct observation grid times
//   Framework.py:685-697 + stats.py:41   masked chi        -> odl_score (warp-cooperative)
//   stats.py:49-56  R^2                                    -> odl_score
//   Framework.py:41-48   _Fit_worker loop                  -> odl_sweep_kernel
//   Samplers.py:53-174   MetropolisHastings                -> odl_mcmc_kernel (one chain per thread)
//   (stiff parameter regions, LSODA's BDF branch)         -> variable-order BDF: odl_*_bdf_kernel, odl_mcmc_auto_kernel
//                                                             (DOPRI5, then BDF, per solve); ROS23: odl_*_ros23_kernel
//
// Execution model: every thread owns one ODE system at a time and keeps its state, the 7 stage
// derivatives and the parameter vector in registers.  The main loop is a flat state machine:
//     [finished lanes: warp-cooperative chi/R^2, write results, fetch next work, re-init]
//     [one adaptive step attempt for all lanes that have work]
//     [lanes whose accepted step crossed observation times: dense output -> shared staging]
// so a lane that finishes refills immediately (sweep: next parameter set from a global counter;
// MCMC: accept/reject and propose the chain's next point) while its warp mates keep stepping; a warp
// leaves when a vote says no lane has work.  Trajectories never reach HBM: predictions at the
// observation slots are staged per thread in shared memory and consumed by the fused scorer.

#ifndef ODL_BLOCK
#define ODL_BLOCK 128
#endif
#ifndef ODL_MINBLOCKS
#define ODL_MINBLOCKS 4
#endif
#ifndef ODL_DENSE
#define ODL_DENSE 1
#endif
#ifndef ODL_Y0P
#define ODL_Y0P 1
#endif
#ifndef ODL_INNER
#define ODL_INNER 8
#endif
#ifndef ODL_MINBLOCKS_MCMC
#define ODL_MINBLOCKS_MCMC ODL_MINBLOCKS
#endif
// Compilation units.  The library compiles ONE kernel (the small ordering trio counts as one) per NVRTC
// program, on demand and in parallel host threads, so that a new model is ready when the kernels a call needs are
// (odl_capi.cu: Unit): -DODL_UNIT=<k> keeps only that unit's __global__ functions, 0 (default) keeps everything.
//   1 sweep   2 traj   3 mcmc   4 sweep_bdf   5 mcmc_bdf   6 sweep_ros23   7 mcmc_ros23   8 mcmc_auto
//   9 sweep_radau5   10 mcmc_radau5   11 sweep_coop   12 mcmc_coop   13 order_key / order_scan / order_scatter
#ifndef ODL_UNIT
#define ODL_UNIT 0
#endif
#define ODL_HAS(u) (ODL_UNIT == 0 || ODL_UNIT == (u))
#ifndef ODL_TIMELINE
#define ODL_TIMELINE 0     // 1: development build that timestamps the feed of the stiff pass (OdlSweepArgs.timeline)
#endif
#ifndef ODL_MINBLOCKS_ROS
#define ODL_MINBLOCKS_ROS (ODL_MINBLOCKS > 2 ? ODL_MINBLOCKS - 1 : ODL_MINBLOCKS)
#endif

// Small systems (n <= 8): every loop over states / parameters is unrolled and the arrays live in registers.
// Larger systems: rolled loops, arrays in (L1-resident) local memory -- functional, not yet tuned.
#if ODL_N <= 8
#define ODL_UNROLL _Pragma("unroll")
#define ODL_SMALL 1
#else
#define ODL_UNROLL _Pragma("unroll 1")
#define ODL_SMALL 0
#endif
#define ODL_FULL 0xffffffffu
#define ODL_DBL_MIN 2.2250738585072014e-308
#define ODL_DBL_MAX 1.7976931348623157e308

#include "odl_abi.h"

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool odl_finite(double x) { return fabs(x) <= ODL_DBL_MAX; }

#ifndef ODL_HOST_HARNESS
__device__ __forceinline__ double odl_shfl_xor(double v, int m) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_xor_sync(ODL_FULL, lo, m);
  hi = __shfl_xor_sync(ODL_FULL, hi, m);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double odl_warp_sum(double v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += odl_shfl_xor(v, m);
  return v;
}

#endif  // ODL_HOST_HARNESS

// Philox4x32-10 (Salmon et al., SC'11) -- counter-based, one independent stream per (chain, iteration)
struct OdlPhilox { unsigned int x, y, z, w; };
__device__ __forceinline__ OdlPhilox odl_philox(unsigned int c0, unsigned int c1, unsigned int c2, unsigned int c3,
                                                 unsigned int k0, unsigned int k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned int hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const unsigned int hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const unsigned int n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  OdlPhilox o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
  return o;
}
// 53-bit uniform in [0,1) from two words; (0,1] variant for the logarithm of Box-Muller
__device__ __forceinline__ double odl_u53(unsigned int a, unsigned int b) {
  const unsigned long long v = (((unsigned long long)a) << 21) ^ ((unsigned long long)b >> 11);
  return (double)(v & ((1ull << 53) - 1)) * 1.1102230246251565e-16;  // 2^-53
}

// ------------------------------------------------------------------------------------------------
// shared-memory view of the data tables + per-thread staging
// ------------------------------------------------------------------------------------------------
// ODL_LOGTAB intervals of [1, 2) for odl_log below: (invc, -ln invc) per interval, 4 KB, filled by every CTA at start
struct OdlShared {
  double* slot_t; double* lnO; double* w; double* lin; int* src; double* stage; double* logtab;
};
__device__ __forceinline__ OdlShared odl_carve(double* base, const OdlData& D) {
  OdlShared S;
  S.logtab = base;               // first: 16-byte aligned pairs
  base += 2 * ODL_LOGTAB;
  S.slot_t = base;
  S.lnO = S.slot_t + D.n_slot;
  S.w = S.lnO + D.n_obs;
  S.lin = S.w + D.n_obs;
  S.src = (int*)(S.lin + D.n_obs);
  S.stage = S.lin + D.n_obs + (D.n_obs + 1) / 2;
  return S;
}
#ifndef ODL_HOST_HARNESS
extern __shared__ double odl_smem[];   // dynamic shared memory of every kernel: tables, then staging
__device__ __forceinline__ void odl_load_tables(const OdlShared& S, const OdlData& D) {
  for (int i = threadIdx.x; i < ODL_LOGTAB; i += blockDim.x) {
    const double invc = 1.0 / (1.0 + ((double)i + 0.5) * (1.0 / ODL_LOGTAB));
    S.logtab[2 * i] = invc; S.logtab[2 * i + 1] = -log(invc);
  }
  for (int i = threadIdx.x; i < D.n_slot; i += blockDim.x) S.slot_t[i] = D.slot_t[i];
  for (int i = threadIdx.x; i < D.n_obs; i += blockDim.x) {
    S.lnO[i] = D.obs_lnO[i]; S.w[i] = D.obs_w[i]; S.lin[i] = D.obs_lin[i]; S.src[i] = D.obs_src[i];
  }
  __syncthreads();
}

// ln(x) for the scorer (np.log at Framework.py:692).  CUDA's log() is ~80 instructions and ran 37 times per solve: 7.9 %
// of the sweep kernel's instructions (profiles/r1f).  Table-driven instead (Tang / glibc style), ~20 instructions: x =
// 2^k m, m in [1,2); interval i = top 8 mantissa bits, invc_i ~ 1/centre_i; r = m invc_i - 1 is EXACT up to one rounding
// of a number below 2^-9 (FMA), and ln m = -ln(invc_i) + log1p(r) holds exactly for whatever invc_i is -- so the only
// table error is the rounding of -ln(invc_i); log1p(r) = r - r^2/2 + r^3/3 - r^4/4 + r^5/5 (next term < 1e-17).
// k ln2 in two pieces as in fdlibm (k ln2_hi is exact).  Error <= ~1 ulp of the result, like log() itself; arguments
// that are not positive normal numbers (<= 0, denormal, inf, NaN: masked terms) take log().
#ifndef ODL_FASTLOG
#define ODL_FASTLOG 1
#endif
__device__ __forceinline__ double odl_log(const OdlShared& S, double x) {
#if ODL_FASTLOG
  const int hi = __double2hiint(x), lo = __double2loint(x);
  if ((unsigned int)(hi - 0x00100000) < 0x7fe00000u) {
    const double kd = (double)((hi >> 20) - 1023);
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    const double2 tc = reinterpret_cast<const double2*>(S.logtab)[(hi >> 12) & (ODL_LOGTAB - 1)];
    const double r = fma(m, tc.x, -1.0);
    double t = fma(r, 0.2, -0.25);
    t = fma(r, t, 0.3333333333333333);
    t = fma(r, t, -0.5);
    const double head = fma(kd, 6.93147180369123816490e-01, tc.y);
    double tail = fma(r * r, t, r);
    tail = fma(kd, 1.90821492927058770002e-10, tail);
    return head + tail;
  }
#endif
  return log(x);
}

// Warp-cooperative chi + R^2 of the system staged by lane `leader` (stats.py:41, :49-56).
// All 32 lanes must call.  Every lane returns the reduced values.
__device__ __forceinline__ void odl_score(const OdlShared& S, const OdlData& D, const double* stage_leader, int lane,
                                          double* pred_out, double& chi, double& ssres, int& nvalid) {
  double c = 0.0, s = 0.0;
  int k = 0;
  for (int o = lane; o < D.n_obs; o += 32) {
    const double pred = stage_leader[S.src[o]];
    if (pred_out) pred_out[o] = pred;
    const double d = __dadd_rn(S.lnO[o], -odl_log(S, pred));   // masked_invalid(O) - C
    const double dd = __dmul_rn(d, d);                          // (...)**2   : masked when not finite
    // / (2*S**2), as a multiplication by the reciprocal the host tabulated (no division per observation: 1.9 % of the
    // sweep kernel's instructions in profiles/r2f, plus the special case a zero numerator needed).  np.ma masks the
    // term when the quotient is not finite or the denominator vanishes against the numerator: sigma = 0 (or a
    // denormal 2 sigma^2) has the reciprocal inf -> inf or NaN here, masked as well.
    const double term = __dmul_rn(dd, S.w[o]);
    const bool ok = odl_finite(dd) && odl_finite(term);
    if (ok) { c += term; ++k; }
    const double r = __dadd_rn(pred, -S.lin[o]);
    const double rr = __dmul_rn(r, r);
    if (rr == rr) s += rr;                                      // np.nansum
  }
  chi = odl_warp_sum(c);
  ssres = odl_warp_sum(s);
  nvalid = __reduce_add_sync(ODL_FULL, k);
}


// Same sums, one system per lane (the MCMC kernel finalises all lanes of a warp together).
__device__ __forceinline__ void odl_score_self(const OdlShared& S, const OdlData& D, const double* stage, double& chi,
                                               double& ssres, int& nvalid) {
  double c = 0.0, s = 0.0;
  int k = 0;
  for (int o = 0; o < D.n_obs; ++o) {
    const double pred = stage[S.src[o]];
    const double d = __dadd_rn(S.lnO[o], -odl_log(S, pred));
    const double dd = __dmul_rn(d, d);
    const double term = __dmul_rn(dd, S.w[o]);
    const bool ok = odl_finite(dd) && odl_finite(term);
    if (ok) { c += term; ++k; }
    const double r = __dadd_rn(pred, -S.lin[o]);
    const double rr = __dmul_rn(r, r);
    if (rr == rr) s += rr;
  }
  chi = c; ssres = s; nvalid = k;
}
#endif  // ODL_HOST_HARNESS

// ------------------------------------------------------------------------------------------------
// Dormand-Prince 5(4), coefficients of Hairer/Norsett/Wanner (dopri5.f); step controller: see ODL_PI_BETA.
// ------------------------------------------------------------------------------------------------
struct OdlStepper {
  double y[ODL_N];       // state at t
  double k1[ODL_N];      // f(t, y)  (FSAL)
  double t, h, tend;
  float lgfac;           // log2 of the PI controller's previous error (facold of dopri5.f), >= ODL_LG_FACMIN
  int nsteps;            // attempted steps
  int slot;              // next observation slot to produce
  int status;
  int iasti, nonsti;     // stiffness detection counters
  bool last_rejected;
};

// Approximate fp64 reciprocal (one MUFU.RCP64H, ~20 bits): only step-size control sees it -- the scaled
// error norms need 2-3 digits -- so no 20-instruction IEEE division and no fp64<->fp32 conversions.
#ifndef ODL_HOST_HARNESS
__device__ __forceinline__ double odl_rcp_approx(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  return r;
}
#else
__device__ __forceinline__ double odl_rcp_approx(double x) { return 1.0 / x; }
#endif
__device__ __forceinline__ float odl_err_ratio(double e, double sk) { return (float)(e * odl_rcp_approx(sk)); }
// fp32 square root for step-size heuristics: one MUFU instead of the IEEE sequence (8 instructions)
#ifndef ODL_HOST_HARNESS
__device__ __forceinline__ float odl_sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
#else
__device__ __forceinline__ float odl_sqrt_approx(float x) { return sqrtf(x); }
#endif

// 2^x on the SFU without exp2f()'s denormal rescaling (step-size factors only), and 1/x to full precision from the
// 20-bit SFU seed + two Newton steps (5 instructions instead of the IEEE division's ~14; used for the dense-output
// abscissa, where an ulp of theta is 1e-16 of a step)
#ifndef ODL_HOST_HARNESS
__device__ __forceinline__ float odl_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ double odl_rcp(double x) {
  double r = odl_rcp_approx(x);
  r = fma(r, fma(-x, r, 1.0), r);
  r = fma(r, fma(-x, r, 1.0), r);
  return r;
}
#else
__device__ __forceinline__ float odl_ex2(float x) { return exp2f(x); }
__device__ __forceinline__ double odl_rcp(double x) { return 1.0 / x; }
#endif
// Step-size controller of the DOPRI5 kernels: h_new = h * SAFETY * err^-(0.2 - 0.75 beta) * err_old^beta, factors in
// [0.2, 10], no growth right after a rejection.  Hairer's dopri5.f stabilises with beta = 0.04 (the default here);
// beta = 0 is the controller of scipy.integrate.RK45 (SAFETY 0.9, plain I-controller).  On 12,000 two_i prior draws
// (host harness, rows that finish within 704 attempts, errors against the 1e-12 solution at the observation times):
//   beta 0.04: 1,046k attempted steps, error median 4.0e-9, max 9.3e-8      beta 0.02: 1,002k, 5.0e-9, 1.1e-7
//   beta 0   :   972k (-7.2 %),                   5.9e-9,     1.2e-7      beta 0.08: 1,296k, 1.2e-9, 3.9e-8
// (safety 0.95 at beta 0.04: 980k, 5.9e-9, 1.4e-7): one curve of work against accuracy -- the reference's own LSODA solve
// at the same rtol = atol = 1.49e-8 is 4e-8 .. 1.3e-7 from the tight solution on the golden rows.  Measured on B200 with
// beta = 0 (-DODL_PI_BETA=0.0f through ODL_KERNEL_DEFINES): 4096 chains 71 -> 75 M chain-steps/s, the sweep unchanged
// (3.2 ms: its DOPRI5 pass ends 3 % sooner and then waits for the stiff pass) -- not adopted.
#ifndef ODL_PI_BETA
#define ODL_PI_BETA 0.04f
#endif
#ifndef ODL_PI_SAFETY
#define ODL_PI_SAFETY 0.9f
#endif
#define ODL_LG_FACMIN -13.287712f              /* log2(1e-4): floor of the PI controller's memory (Hairer: facold >= 1e-4) */

// Butcher tableau of DOPRI5 (Hairer/Norsett/Wanner, dopri5.f) in constant memory: DFMA takes c[bank][offset]
// operands directly, whereas literals are re-materialised with two UMOVs per use (13 % of the issued
// instructions in the first profile, profiles/r1_sweep_bulk_ncu.txt).
__constant__ double ODL_TAB[36] = {
    0.2,                                                                    // 0  a21
    3.0 / 40.0, 9.0 / 40.0,                                                 // 1  a31 a32
    44.0 / 45.0, -56.0 / 15.0, 32.0 / 9.0,                                  // 3  a41 a42 a43
    19372.0 / 6561.0, -25360.0 / 2187.0, 64448.0 / 6561.0, -212.0 / 729.0,  // 6  a51..a54
    9017.0 / 3168.0, -355.0 / 33.0, 46732.0 / 5247.0, 49.0 / 176.0, -5103.0 / 18656.0,   // 10 a61..a65
    35.0 / 384.0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0,          // 15 a71 a73 a74 a75 a76
    71.0 / 57600.0, -71.0 / 16695.0, 71.0 / 1920.0, -17253.0 / 339200.0, 22.0 / 525.0, -1.0 / 40.0,  // 20 e1 e3 e4 e5 e6 e7
    -12715105075.0 / 11282082432.0, 87487479700.0 / 32700410799.0, -10690763975.0 / 1880347072.0,
    701980252875.0 / 199316789632.0, -1453857185.0 / 822651844.0, 69997945.0 / 29380423.0,            // 26 d1 d3 d4 d5 d6 d7
    0.3, 0.8, 8.0 / 9.0, 0.0};                                                                        // 32 c3 c4 c5
#define ODL_T(i) ODL_TAB[i]

// y0_from_params: the '<state>0' convention.  The reference copies such a parameter into the state's initial value
// in ONE place only -- the proposal loop of MetropolisHastings (Samplers.py:110-114; restored on reject :139-143) --
// so it holds for the solves of PROPOSALS.  integrate(), the survey (_Fit_worker, Framework.py:41-48) and the chain's
// a-priori solve (Samplers.py:88) start from the model's istates (Framework.py:647-650) whatever the parameter says.
__device__ __forceinline__ void odl_init_system(OdlStepper& st, const double (&p)[ODL_P], const OdlData& D,
                                                const OdlOpts& O, const double* y0_override, bool y0_from_params) {
ODL_UNROLL
  for (int i = 0; i < ODL_N; ++i) {
    double v = y0_override ? y0_override[i] : D.y0[i];
#if ODL_Y0P
    if (y0_from_params) {
      const int src = D.y0_from_param[i];
ODL_UNROLL
      for (int q = 0; q < ODL_P; ++q) if (src == q) v = p[q];   // Samplers.py:110-114
    }
#endif
    st.y[i] = v;
  }
  st.t = D.t0;
  st.tend = D.slot_t[D.n_slot - 1];
  st.nsteps = 0; st.slot = 0; st.status = ODL_OK; st.iasti = 0; st.nonsti = 0;
  st.lgfac = ODL_LG_FACMIN; st.last_rejected = false;
  odl_rhs(st.y, st.t, p, st.k1);
  const double span = st.tend - st.t;
  const double hmax = (O.hmax > 0.0) ? O.hmax : span;
  double h = O.h0;
  if (!(h > 0.0)) {
    // Hairer's hinit: h ~ 0.01 |y|/|f|, refined with a difference quotient of f along an Euler step.
    // A heuristic -> fp32 norms (SFU reciprocal / pow), guarded below against overflow.
    double dnf = 0.0, dny = 0.0;
    double rsk[ODL_N];
ODL_UNROLL
    for (int i = 0; i < ODL_N; ++i) {
      rsk[i] = odl_rcp_approx(O.atol + O.rtol * fabs(st.y[i]));
      const double a = st.k1[i] * rsk[i], b = st.y[i] * rsk[i];
      dnf += a * a; dny += b * b;
    }
    const float fnf = (float)dnf, fny = (float)dny;
    const float hf = (fnf <= 1e-10f || fny <= 1e-10f) ? 1e-6f : 0.01f * sqrtf(fny * __frcp_rn(fnf));
    h = fmin((double)hf, hmax);
    double y1[ODL_N], f1[ODL_N];
ODL_UNROLL
    for (int i = 0; i < ODL_N; ++i) y1[i] = st.y[i] + h * st.k1[i];
    odl_rhs(y1, st.t + h, p, f1);
    double d2 = 0.0;
ODL_UNROLL
    for (int i = 0; i < ODL_N; ++i) {
      const double a = (f1[i] - st.k1[i]) * rsk[i];
      d2 += a * a;
    }
    const float der2 = sqrtf((float)d2) * __frcp_rn((float)h);
    const float der12 = fmaxf(der2, sqrtf(fnf));
    const float h1 = (der12 <= 1e-15f) ? fmaxf(1e-6f, (float)h * 1e-3f) : __powf(0.01f * __frcp_rn(der12), 0.2f);
    h = fmin(fmin(100.0 * h, (double)h1), hmax);
  }
  if (!(h > 0.0) || !odl_finite(h)) h = 1e-6 * (span > 0.0 ? span : 1.0);
  st.h = h;
}

// dense output of the last accepted step [t0, t0+h] at time ts (Hairer's contd5)

// One step attempt.  On acceptance the lambda-like macro ODL_ON_SLOT is not used; instead the caller
// passes a functor `sink(slot, yi)` that receives the interpolated state at every crossed slot.
template <class Sink>
__device__ __forceinline__ void odl_dopri5_attempt(OdlStepper& st, const double (&p)[ODL_P], const OdlShared& S,
                                                   const OdlData& D, const OdlOpts& O, Sink& sink) {
  const double t = st.t;
  double h = st.h;
  bool last = false;
#if ODL_DENSE
  if ((t + 1.01 * h - st.tend) > 0.0) { h = st.tend - t; last = true; }
  const bool truncated = false;
  const double h_untrunc = h;
#else
  // land exactly on the next observation slot instead of interpolating
  const double ttarget = S.slot_t[st.slot];
  bool truncated = false;
  const double h_untrunc = h;
  if ((t + 1.01 * h - ttarget) > 0.0) { h = ttarget - t; truncated = true; last = (st.slot == D.n_slot - 1); }
#endif
  // attempts of a solve that is finished (slot == n_slot) or has failed are not counted: in the sweep and chain kernels a
  // lane keeps running this code while it waits for its warp, and the count it holds stays that of its solve
  st.nsteps += (st.slot < D.n_slot && st.status == ODL_OK) ? 1 : 0;
  double k2[ODL_N], k3[ODL_N], k4[ODL_N], k5[ODL_N], k6[ODL_N], k7[ODL_N], yt[ODL_N], yn[ODL_N];
ODL_UNROLL
  for (int i = 0; i < ODL_N; ++i) yt[i] = st.y[i] + h * (ODL_T(0) * st.k1[i]);
  odl_rhs(yt, t + ODL_T(0) * h, p, k2);
ODL_UNROLL
  for (int i = 0; i < ODL_N; ++i) yt[i] = st.y[i] + h * (ODL_T(1) * st.k1[i] + ODL_T(2) * k2[i]);
  odl_rhs(yt, t + ODL_T(32) * h, p, k3);
ODL_UNROLL
  for (int i = 0; i < ODL_N; ++i)
    yt[i] = st.y[i] + h * (ODL_T(3) * st.k1[i] + ODL_T(4) * k2[i] + ODL_T(5) * k3[i]);
  odl_rhs(yt, t + ODL_T(33) * h, p, k4);
ODL_UNROLL
  for (int i = 0; i < ODL_N; ++i)
    yt[i] = st.y[i] + h * (ODL_T(6) * st.k1[i] + ODL_T(7) * k2[i] + ODL_T(8) * k3[i] + ODL_T(9) * k4[i]);
  odl_rhs(yt, t + ODL_T(34) * h, p, k5);
ODL_UNROLL
  for (int i = 0; i < ODL_N; ++i)
    yt[i] = st.y[i] + h * (ODL_T(10) * st.k1[i] + ODL_T(11) * k2[i] + ODL_T(12) * k3[i] + ODL_T(13) * k4[i] +
                           ODL_T(14) * k5[i]);
  const double tph = t + h;
  odl_rhs(yt, tph, p, k6);
ODL_UNROLL
  for (int i = 0; i < ODL_N; ++i)
    yn[i] = st.y[i] + h * (ODL_T(15) * st.k1[i] + ODL_T(16) * k3[i] + ODL_T(17) * k4[i] + ODL_T(18) * k5[i] +
                           ODL_T(19) * k6[i]);
  odl_rhs(yn, tph, p, k7);

  // embedded error estimate, scaled RMS norm.  Scale: atol + rtol (|y| + |y_new|)/2 -- one DADD with |.| operand
  // modifiers instead of the compare-and-select sequence fp64 max costs (5 % of the kernel's instructions in
  // profiles/r1c); never larger than Hairer's max(|y|, |y_new|), i.e. never less strict.  Non-finite y_new shows in
  // the sum below (a finite sum that overflows only costs a rejected step).
  double errsq = 0.0, ysum = 0.0;
  const double rtol_half = 0.5 * O.rtol;
ODL_UNROLL
  for (int i = 0; i < ODL_N; ++i) {
    const double e = h * (ODL_T(20) * st.k1[i] + ODL_T(21) * k3[i] + ODL_T(22) * k4[i] + ODL_T(23) * k5[i] +
                          ODL_T(24) * k6[i] + ODL_T(25) * k7[i]);
    const double sk = O.atol + rtol_half * (fabs(st.y[i]) + fabs(yn[i]));
    const double r = e * odl_rcp_approx(sk);
    errsq += r * r;
    ysum += fabs(yn[i]);
  }
  const bool finite_all = odl_finite(ysum);
  const float err = odl_sqrt_approx((float)errsq * (1.0f / ODL_N));
  // step-size controller (ODL_PI_BETA / ODL_PI_SAFETY above): h_new = h * safety * facold^beta / err^(0.2 - 0.75 beta), one SFU exp2
  const float lg_err = __log2f(err);
  if (err <= 1.0f && finite_all) {
    // ---- accepted ----
    const float inv = ODL_PI_SAFETY * odl_ex2(ODL_PI_BETA * st.lgfac - (0.2f - ODL_PI_BETA * 0.75f) * lg_err);
    float fac = fminf(10.0f, fmaxf(0.2f, inv));                 // growth <= 10, shrink <= 5
    if (!(err > 0.f)) fac = 10.0f;
    if (st.last_rejected) fac = fminf(fac, 1.0f);               // no growth right after a rejection (fp32: no fp64 min)
    double hnew = h * (double)fac;
    st.lgfac = fmaxf(lg_err, ODL_LG_FACMIN);
    if (O.stiff_check && ((st.nsteps % 10) == 0 || st.iasti > 0)) {
      // Hairer's test: h * |k7 - k6| / |ynew - y6| approximates h * |lambda_max|
      double num = 0.0, den = 0.0;
ODL_UNROLL
      for (int i = 0; i < ODL_N; ++i) {
        const double a = k7[i] - k6[i], b = yn[i] - yt[i];
        num += a * a; den += b * b;
      }
      if (den > 0.0) {
        // h |lambda| ~ h sqrt(num/den) > 3.25, without the fp64 square root and division (the test runs on one step
        // in ten of every lane, i.e. on most warp steps for one lane or another)
        if (h * h * num > 10.5625 * den) {
          st.nonsti = 0;
          // stiff for 15 checks in a row AND still far from the end at this (stability-limited) step size:
          // hand the system to the Rosenbrock path; mildly stiff / nearly finished systems stay here
          if (++st.iasti >= 15 && (st.tend - tph) > (double)O.stiff_min_steps * h) st.status = ODL_STIFF;
        } else if (++st.nonsti == 6) st.iasti = 0;
      }
    }
#if ODL_DENSE
    const double tnew = last ? st.tend : tph;
#else
    const double tnew = truncated ? ttarget : tph;
#endif
#if ODL_DENSE
    if constexpr (Sink::kObservedOnly && (ODL_NOUT < ODL_N) && ODL_SMALL) {
      // The sink wants the OBSERVED columns only, and they are sums of state groups (odl_observe is linear): interpolate
      // the ODL_NOUT sums instead of the ODL_N states.  The continuous extension is linear in (y, y_new, k1, k3..k7), so
      // the polynomial of a sum is the sum of the polynomials: for the two_i model (4 states, 2 columns) half the
      // coefficient and Horner work of this block -- which a warp issues on nearly every step for the 9 of its 32 lanes
      // that crossed an observation time -- for 16 additions (profiles/r2u_*: dense output 15 % of the instructions).
      if (st.slot < D.n_slot && S.slot_t[st.slot] <= tnew) {
        double oy[ODL_NOUT], on[ODL_NOUT], o1[ODL_NOUT], o3[ODL_NOUT], o4[ODL_NOUT], o5[ODL_NOUT], o6[ODL_NOUT], o7[ODL_NOUT];
        odl_observe(st.y, oy); odl_observe(yn, on); odl_observe(st.k1, o1); odl_observe(k3, o3);
        odl_observe(k4, o4); odl_observe(k5, o5); odl_observe(k6, o6); odl_observe(k7, o7);
        double rc2[ODL_NOUT], rc3[ODL_NOUT], rc4[ODL_NOUT], rc5[ODL_NOUT];
#pragma unroll
        for (int c = 0; c < ODL_NOUT; ++c) {
          rc2[c] = on[c] - oy[c];
          rc3[c] = h * o1[c] - rc2[c];
          rc4[c] = rc2[c] - h * o7[c] - rc3[c];
          rc5[c] = h * (ODL_T(26) * o1[c] + ODL_T(27) * o3[c] + ODL_T(28) * o4[c] + ODL_T(29) * o5[c] +
                        ODL_T(30) * o6[c] + ODL_T(31) * o7[c]);
        }
        const double rh = odl_rcp(h);
        do {
          const double th = (S.slot_t[st.slot] - t) * rh, th1 = 1.0 - th;
          double out[ODL_NOUT];
#pragma unroll
          for (int c = 0; c < ODL_NOUT; ++c)
            out[c] = oy[c] + th * (rc2[c] + th1 * (rc3[c] + th * (rc4[c] + th1 * rc5[c])));
          sink.put(st.slot, out);
          ++st.slot;
        } while (st.slot < D.n_slot && S.slot_t[st.slot] <= tnew);
      }
    } else
    if (st.slot < D.n_slot && S.slot_t[st.slot] <= tnew) {
      double rc2[ODL_N], rc3[ODL_N], rc4[ODL_N], rc5[ODL_N];
ODL_UNROLL
      for (int i = 0; i < ODL_N; ++i) {
        rc2[i] = yn[i] - st.y[i];
        rc3[i] = h * st.k1[i] - rc2[i];
        rc4[i] = rc2[i] - h * k7[i] - rc3[i];
        rc5[i] = h * (ODL_T(26) * st.k1[i] + ODL_T(27) * k3[i] + ODL_T(28) * k4[i] + ODL_T(29) * k5[i] +
                      ODL_T(30) * k6[i] + ODL_T(31) * k7[i]);
      }
      const double rh = odl_rcp(h);
      do {
        const double th = (S.slot_t[st.slot] - t) * rh, th1 = 1.0 - th;
        double yi[ODL_N];
ODL_UNROLL
        for (int i = 0; i < ODL_N; ++i)
          yi[i] = st.y[i] + th * (rc2[i] + th1 * (rc3[i] + th * (rc4[i] + th1 * rc5[i])));
        sink(st.slot, yi);
        ++st.slot;
      } while (st.slot < D.n_slot && S.slot_t[st.slot] <= tnew);
    }
#else
    if (truncated) {
      // several slots may share a time only if the host deduplicated badly; loop is defensive
      do { sink(st.slot, yn); ++st.slot; } while (st.slot < D.n_slot && S.slot_t[st.slot] <= tnew);
      hnew = fmax(hnew, h_untrunc);
    }
#endif
ODL_UNROLL
    for (int i = 0; i < ODL_N; ++i) { st.y[i] = yn[i]; st.k1[i] = k7[i]; }
    st.t = tnew;
    st.last_rejected = false;
    st.h = hnew;
    (void)truncated; (void)h_untrunc;
  } else {
    // ---- rejected ----
    double hnew;
    if (err == err && finite_all && err < 3.0e38f)
      hnew = h * (double)fmaxf(0.2f, ODL_PI_SAFETY * odl_ex2(-(0.2f - ODL_PI_BETA * 0.75f) * lg_err));
    else hnew = 0.2 * h;                                        // NaN / overflow inside the step
    st.last_rejected = true;
    st.h = hnew;
    if (!(fabs(hnew) > 4.0 * 2.220446049250313e-16 * fmax(fabs(t), fabs(st.tend)))) st.status = ODL_HUNDERFLOW;
  }
  // Step budget.  One compare per attempt until the first of the two checks is due: the cap itself, and -- capped pass
  // of the cohort sweep -- the projection check: a system whose progress after early_check_steps attempts projects to
  // more than max_steps in total leaves now instead of burning the rest of its budget (its warp waits for it).
  // (Earlier checks against multiples of the cap -- 64 / 128 / 256 attempts against 8x / 4x / 2x -- were measured: the
  // hopeless systems reach the stiff pass 0.4 ms sooner, but 40 % more rows do (slow starters), and that pass, which
  // runs beside this one on a quarter of the SMs, is bound by its throughput: 3.88 against 3.71 ms per 1M rows.)
  const int first_check = (O.early_check_steps > 0 && O.early_check_steps < O.max_steps) ? O.early_check_steps : O.max_steps;
  if (st.nsteps >= first_check && st.slot < D.n_slot && st.status == ODL_OK) {
    if (st.nsteps >= O.max_steps) st.status = ODL_MAXSTEPS;
    else if (st.nsteps == O.early_check_steps && (double)st.nsteps * (st.tend - D.t0) > (double)O.max_steps * (st.t - D.t0))
      st.status = ODL_MAXSTEPS;
  }
  // A solve that has stopped stands still: in the sweep kernel its lane keeps running this code until its warp's next
  // visit of the write-back, and with a step of zero every further attempt leaves (t, y, k1, slot) exactly as they are --
  // what the stiff pass takes over (OdlSweepArgs.handover) must not depend on how long the lane waited.
  if (st.status != ODL_OK) st.h = 0.0;
}

// slots at (or before) the start time take the initial state (odeint returns y0 at times[0])
template <class Sink>
__device__ __forceinline__ void odl_emit_initial_slots(OdlStepper& st, const OdlShared& S, const OdlData& D, Sink& sink) {
  while (st.slot < D.n_slot && S.slot_t[st.slot] <= st.t) { sink(st.slot, st.y); ++st.slot; }
}

struct OdlStageSink {            // observation columns -> per-thread shared staging
  static constexpr bool kObservedOnly = true;                    // see the dense output of odl_dopri5_attempt
  double* stage;
  __device__ __forceinline__ void operator()(int slot, const double (&yi)[ODL_N]) {
    double out[ODL_NOUT];
    odl_observe(yi, out);
    put(slot, out);
  }
  __device__ __forceinline__ void put(int slot, const double (&out)[ODL_NOUT]) {
ODL_UNROLL
    for (int c = 0; c < ODL_NOUT; ++c) stage[slot * ODL_NOUT + c] = out[c];
  }
};
struct OdlTrajSink {             // raw states -> global trajectory
  static constexpr bool kObservedOnly = false;
  double* traj;
  __device__ __forceinline__ void put(int, const double (&)[ODL_NOUT]) {}
  __device__ __forceinline__ void operator()(int slot, const double (&yi)[ODL_N]) {
ODL_UNROLL
    for (int i = 0; i < ODL_N; ++i) traj[(long long)slot * ODL_N + i] = yi[i];
  }
};


// ------------------------------------------------------------------------------------------------
// Rosenbrock-W 2(3) of Shampine & Reichelt ("The MATLAB ODE suite", SIAM J. Sci. Comput. 18, 1997: ode23s)
// for the stiff parameter regions.  Analytic Jacobian from the tracer, W = I - h d J factorised in
// registers (fully unrolled LU with partial pivoting; swap decisions kept in a bit mask and replayed on
// the three right-hand sides), L-stable second-order advance with third-order error estimate, FSAL on f.
// ------------------------------------------------------------------------------------------------
#define ODL_ROS_D 0.29289321881345254      /* 1/(2+sqrt 2) */
#define ODL_ROS_E32 7.414213562373095      /* 6+sqrt 2 */

#if ODL_SMALL
struct OdlLU {
  double a[ODL_N][ODL_N];
  double inv[ODL_N];
  unsigned long long swaps;
};
__device__ __forceinline__ void odl_lu_factor(OdlLU& F) {
  unsigned long long sw = 0ull;
  int bit = 0;
ODL_UNROLL
  for (int k = 0; k < ODL_N; ++k) {
ODL_UNROLL
    for (int i = k + 1; i < ODL_N; ++i) {
      const bool s = fabs(F.a[i][k]) > fabs(F.a[k][k]);
      if (s) sw |= (1ull << (bit & 63));
      ++bit;
      // rows are exchanged from column k on only (LINPACK dgefa/dgesl convention): the multipliers already
      // stored in columns < k stay put, matching the interleaved swap-then-eliminate order of odl_lu_solve
ODL_UNROLL
      for (int j = k; j < ODL_N; ++j) {
        const double u = F.a[k][j], v = F.a[i][j];
        F.a[k][j] = s ? v : u;
        F.a[i][j] = s ? u : v;
      }
    }
    F.inv[k] = 1.0 / F.a[k][k];
ODL_UNROLL
    for (int i = k + 1; i < ODL_N; ++i) {
      const double l = F.a[i][k] * F.inv[k];
      F.a[i][k] = l;
ODL_UNROLL
      for (int j = k + 1; j < ODL_N; ++j) F.a[i][j] -= l * F.a[k][j];
    }
  }
  F.swaps = sw;
}
__device__ __forceinline__ void odl_lu_solve(const OdlLU& F, double (&b)[ODL_N]) {
  if (F.swaps == 0ull) {
    // no row was exchanged (the usual case for I - c J): skip the select pairs that replay the exchanges
ODL_UNROLL
    for (int k = 0; k < ODL_N; ++k) {
ODL_UNROLL
      for (int i = k + 1; i < ODL_N; ++i) b[i] -= F.a[i][k] * b[k];
    }
  } else {
    int bit = 0;
ODL_UNROLL
    for (int k = 0; k < ODL_N; ++k) {
ODL_UNROLL
      for (int i = k + 1; i < ODL_N; ++i) {
        const bool s = (F.swaps >> (bit & 63)) & 1ull;
        ++bit;
        const double u = b[k], v = b[i];
        b[k] = s ? v : u;
        b[i] = s ? u : v;
      }
ODL_UNROLL
      for (int i = k + 1; i < ODL_N; ++i) b[i] -= F.a[i][k] * b[k];
    }
  }
ODL_UNROLL
  for (int k = ODL_N - 1; k >= 0; --k) {
    double acc = b[k];
ODL_UNROLL
    for (int j = k + 1; j < ODL_N; ++j) acc -= F.a[k][j] * b[j];
    b[k] = acc * F.inv[k];
  }
}

#else
struct OdlLU {
  double a[ODL_N][ODL_N];
  double inv[ODL_N];
  int piv[ODL_N];
};
__device__ __forceinline__ void odl_lu_factor(OdlLU& F) {
  for (int k = 0; k < ODL_N; ++k) {
    int best = k;
    double big = fabs(F.a[k][k]);
    for (int i = k + 1; i < ODL_N; ++i) { const double v = fabs(F.a[i][k]); if (v > big) { big = v; best = i; } }
    F.piv[k] = best;
    if (best != k) for (int j = k; j < ODL_N; ++j) { const double u = F.a[k][j]; F.a[k][j] = F.a[best][j]; F.a[best][j] = u; }
    F.inv[k] = 1.0 / F.a[k][k];
    for (int i = k + 1; i < ODL_N; ++i) {
      const double l = F.a[i][k] * F.inv[k];
      F.a[i][k] = l;
      if (l != 0.0) for (int j = k + 1; j < ODL_N; ++j) F.a[i][j] -= l * F.a[k][j];
    }
  }
}
__device__ __forceinline__ void odl_lu_solve(const OdlLU& F, double (&b)[ODL_N]) {
  for (int k = 0; k < ODL_N; ++k) {
    const int q = F.piv[k];
    const double u = b[k]; b[k] = b[q]; b[q] = u;
    const double bk = b[k];
    if (bk != 0.0) for (int i = k + 1; i < ODL_N; ++i) b[i] -= F.a[i][k] * bk;
  }
  for (int k = ODL_N - 1; k >= 0; --k) {
    double acc = b[k];
    for (int j = k + 1; j < ODL_N; ++j) acc -= F.a[k][j] * b[j];
    b[k] = acc * F.inv[k];
  }
}

#endif  // ODL_SMALL

template <class Sink>
__device__ __forceinline__ void odl_ros23_attempt(OdlStepper& st, const double (&p)[ODL_P], const OdlShared& S,
                                                  const OdlData& D, const OdlOpts& O, Sink& sink) {
  const double t = st.t;
  double h = st.h;
  bool last = false;
  if ((t + 1.01 * h - st.tend) > 0.0) { h = st.tend - t; last = true; }
  ++st.nsteps;
  OdlLU F;
  odl_jac(st.y, t, p, F.a);
ODL_UNROLL
  for (int i = 0; i < ODL_N; ++i)
ODL_UNROLL
    for (int j = 0; j < ODL_N; ++j) F.a[i][j] = ((i == j) ? 1.0 : 0.0) - (h * ODL_ROS_D) * F.a[i][j];
  odl_lu_factor(F);
  double hdT[ODL_N];
#if ODL_AUTONOMOUS
ODL_UNROLL
  for (int i = 0; i < ODL_N; ++i) hdT[i] = 0.0;
#else
  odl_dfdt(st.y, t, p, hdT);
ODL_UNROLL
  for (int i = 0; i < ODL_N; ++i) hdT[i] *= h * ODL_ROS_D;
#endif
  double k1[ODL_N], k2[ODL_N], k3[ODL_N], F1[ODL_N], F2[ODL_N], yt[ODL_N], yn[ODL_N];
ODL_UNROLL
  for (int i = 0; i < ODL_N; ++i) k1[i] = st.k1[i] + hdT[i];
  odl_lu_solve(F, k1);
ODL_UNROLL
  for (int i = 0; i < ODL_N; ++i) yt[i] = st.y[i] + (0.5 * h) * k1[i];
  odl_rhs(yt, t + 0.5 * h, p, F1);
ODL_UNROLL
  for (int i = 0; i < ODL_N; ++i) k2[i] = F1[i] - k1[i];
  odl_lu_solve(F, k2);
ODL_UNROLL
  for (int i = 0; i < ODL_N; ++i) { k2[i] += k1[i]; yn[i] = st.y[i] + h * k2[i]; }
  const double tph = t + h;
  odl_rhs(yn, tph, p, F2);
ODL_UNROLL
  for (int i = 0; i < ODL_N; ++i) k3[i] = F2[i] - ODL_ROS_E32 * (k2[i] - F1[i]) - 2.0 * (k1[i] - st.k1[i]) + hdT[i];
  odl_lu_solve(F, k3);
  float errsq = 0.f;
  bool finite_all = true;
ODL_UNROLL
  for (int i = 0; i < ODL_N; ++i) {
    const double e = (h * (1.0 / 6.0)) * (k1[i] - 2.0 * k2[i] + k3[i]);
    const double sk = O.atol + O.rtol * fmax(fabs(st.y[i]), fabs(yn[i]));
    const float r = odl_err_ratio(e, sk);
    errsq += r * r;
    finite_all = finite_all && odl_finite(yn[i]);
  }
  const float err = sqrtf(errsq * (1.0f / ODL_N));
  const float fac = 0.8f * __powf(err, -1.0f / 3.0f);           // h_new = h * clamp(0.8 err^(-1/3), 0.2, 5)
  if (err <= 1.0f && finite_all) {
    double hnew = (err > 0.f) ? h * (double)fminf(5.0f, fmaxf(0.2f, fac)) : 5.0 * h;
    const double tnew = last ? st.tend : tph;
    if (st.slot < D.n_slot && S.slot_t[st.slot] <= tnew) {
      do {   // ode23s interpolant: y + h (k1 s(1-s)/(1-2d) + k2 s(s-2d)/(1-2d))
        const double s1 = (S.slot_t[st.slot] - t) / h;
        const double c1 = s1 * (1.0 - s1) * (1.0 / (1.0 - 2.0 * ODL_ROS_D));
        const double c2 = s1 * (s1 - 2.0 * ODL_ROS_D) * (1.0 / (1.0 - 2.0 * ODL_ROS_D));
        double yi[ODL_N];
ODL_UNROLL
        for (int i = 0; i < ODL_N; ++i) yi[i] = st.y[i] + h * (c1 * k1[i] + c2 * k2[i]);
        sink(st.slot, yi);
        ++st.slot;
      } while (st.slot < D.n_slot && S.slot_t[st.slot] <= tnew);
    }
ODL_UNROLL
    for (int i = 0; i < ODL_N; ++i) { st.y[i] = yn[i]; st.k1[i] = F2[i]; }
    st.t = tnew;
    if (st.last_rejected) hnew = fmin(hnew, h);
    st.last_rejected = false;
    st.h = hnew;
  } else {
    double hnew;
    if (err == err && finite_all && err < 3.0e38f) hnew = h * (double)fmaxf(0.2f, fminf(0.9f, fac));
    else hnew = 0.2 * h;
    st.last_rejected = true;
    st.h = hnew;
    if (!(fabs(hnew) > 4.0 * 2.220446049250313e-16 * fmax(fabs(t), fabs(st.tend)))) st.status = ODL_HUNDERFLOW;
  }
  if (st.nsteps >= O.max_steps && st.slot < D.n_slot && st.status == ODL_OK) st.status = ODL_MAXSTEPS;
}


// ------------------------------------------------------------------------------------------------
// Radau IIA, 3 stages, order 5 (Hairer & Wanner, Solving ODEs II, ch. IV.8: RADAU5) -- the high-order
// stiff stepper.  ROS23 is second order and needs 1e4-1e5 steps at rtol 1.5e-8 on the stiff corner of
// the demo priors where LSODA's BDF needs a few hundred; Radau5 needs about as few as LSODA.
// Simplified Newton on the transformed collocation system (one real and one complex n x n LU per
// step, factorised in registers), analytic Jacobian refreshed every step, embedded third-order error
// estimate, Gustafsson's predictive controller, collocation polynomial as dense output and as the
// starting guess of the next step.
// ------------------------------------------------------------------------------------------------
#define ODL_RAD_MU 3.637834252744496               /* real eigenvalue of A^-1 */
#define ODL_RAD_ALPHA 2.6810828736277523           /* complex pair alpha -+ i beta */
#define ODL_RAD_BETA 3.050430199247411
#define ODL_RAD_C0 0.15505102572168222             /* (4 - sqrt 6)/10 */
#define ODL_RAD_C1 0.6449489742783178              /* (4 + sqrt 6)/10 */
#define ODL_RAD_NEWTON 6

#if ODL_SMALL
struct OdlCLU {                  // complex LU in split storage
  double ar[ODL_N][ODL_N], ai[ODL_N][ODL_N];
  double ir[ODL_N], ii[ODL_N];   // reciprocal pivots
  unsigned long long swaps;
};
__device__ __forceinline__ void odl_clu_factor(OdlCLU& F) {
  unsigned long long sw = 0ull;
  int bit = 0;
ODL_UNROLL
  for (int k = 0; k < ODL_N; ++k) {
ODL_UNROLL
    for (int i = k + 1; i < ODL_N; ++i) {
      const bool s = (fabs(F.ar[i][k]) + fabs(F.ai[i][k])) > (fabs(F.ar[k][k]) + fabs(F.ai[k][k]));
      if (s) sw |= (1ull << (bit & 63));
      ++bit;
ODL_UNROLL
      for (int j = k; j < ODL_N; ++j) {
        const double ur = F.ar[k][j], vr = F.ar[i][j], ui = F.ai[k][j], vi = F.ai[i][j];
        F.ar[k][j] = s ? vr : ur; F.ar[i][j] = s ? ur : vr;
        F.ai[k][j] = s ? vi : ui; F.ai[i][j] = s ? ui : vi;
      }
    }
    const double pr = F.ar[k][k], pi = F.ai[k][k];
    const double den = 1.0 / (pr * pr + pi * pi);
    F.ir[k] = pr * den; F.ii[k] = -pi * den;
ODL_UNROLL
    for (int i = k + 1; i < ODL_N; ++i) {
      const double lr = F.ar[i][k] * F.ir[k] - F.ai[i][k] * F.ii[k];
      const double li = F.ar[i][k] * F.ii[k] + F.ai[i][k] * F.ir[k];
      F.ar[i][k] = lr; F.ai[i][k] = li;
ODL_UNROLL
      for (int j = k + 1; j < ODL_N; ++j) {
        F.ar[i][j] -= lr * F.ar[k][j] - li * F.ai[k][j];
        F.ai[i][j] -= lr * F.ai[k][j] + li * F.ar[k][j];
      }
    }
  }
  F.swaps = sw;
}
__device__ __forceinline__ void odl_clu_solve(const OdlCLU& F, double (&br)[ODL_N], double (&bi)[ODL_N]) {
  int bit = 0;
ODL_UNROLL
  for (int k = 0; k < ODL_N; ++k) {
ODL_UNROLL
    for (int i = k + 1; i < ODL_N; ++i) {
      const bool s = (F.swaps >> (bit & 63)) & 1ull;
      ++bit;
      const double ur = br[k], vr = br[i], ui = bi[k], vi = bi[i];
      br[k] = s ? vr : ur; br[i] = s ? ur : vr;
      bi[k] = s ? vi : ui; bi[i] = s ? ui : vi;
    }
ODL_UNROLL
    for (int i = k + 1; i < ODL_N; ++i) {
      br[i] -= F.ar[i][k] * br[k] - F.ai[i][k] * bi[k];
      bi[i] -= F.ar[i][k] * bi[k] + F.ai[i][k] * br[k];
    }
  }
ODL_UNROLL
  for (int k = ODL_N - 1; k >= 0; --k) {
    double xr = br[k], xi = bi[k];
ODL_UNROLL
    for (int j = k + 1; j < ODL_N; ++j) {
      xr -= F.ar[k][j] * br[j] - F.ai[k][j] * bi[j];
      xi -= F.ar[k][j] * bi[j] + F.ai[k][j] * br[j];
    }
    br[k] = xr * F.ir[k] - xi * F.ii[k];
    bi[k] = xr * F.ii[k] + xi * F.ir[k];
  }
}

#else
struct OdlCLU {
  double ar[ODL_N][ODL_N], ai[ODL_N][ODL_N];
  double ir[ODL_N], ii[ODL_N];
  int piv[ODL_N];
};
__device__ __forceinline__ void odl_clu_factor(OdlCLU& F) {
  for (int k = 0; k < ODL_N; ++k) {
    int best = k;
    double big = fabs(F.ar[k][k]) + fabs(F.ai[k][k]);
    for (int i = k + 1; i < ODL_N; ++i) { const double v = fabs(F.ar[i][k]) + fabs(F.ai[i][k]); if (v > big) { big = v; best = i; } }
    F.piv[k] = best;
    if (best != k) for (int j = k; j < ODL_N; ++j) {
      double u = F.ar[k][j]; F.ar[k][j] = F.ar[best][j]; F.ar[best][j] = u;
      u = F.ai[k][j]; F.ai[k][j] = F.ai[best][j]; F.ai[best][j] = u;
    }
    const double pr = F.ar[k][k], pi = F.ai[k][k];
    const double den = 1.0 / (pr * pr + pi * pi);
    F.ir[k] = pr * den; F.ii[k] = -pi * den;
    for (int i = k + 1; i < ODL_N; ++i) {
      const double lr = F.ar[i][k] * F.ir[k] - F.ai[i][k] * F.ii[k];
      const double li = F.ar[i][k] * F.ii[k] + F.ai[i][k] * F.ir[k];
      F.ar[i][k] = lr; F.ai[i][k] = li;
      if (lr != 0.0 || li != 0.0) for (int j = k + 1; j < ODL_N; ++j) {
        F.ar[i][j] -= lr * F.ar[k][j] - li * F.ai[k][j];
        F.ai[i][j] -= lr * F.ai[k][j] + li * F.ar[k][j];
      }
    }
  }
}
__device__ __forceinline__ void odl_clu_solve(const OdlCLU& F, double (&br)[ODL_N], double (&bi)[ODL_N]) {
  for (int k = 0; k < ODL_N; ++k) {
    const int q = F.piv[k];
    double u = br[k]; br[k] = br[q]; br[q] = u;
    u = bi[k]; bi[k] = bi[q]; bi[q] = u;
    for (int i = k + 1; i < ODL_N; ++i) {
      br[i] -= F.ar[i][k] * br[k] - F.ai[i][k] * bi[k];
      bi[i] -= F.ar[i][k] * bi[k] + F.ai[i][k] * br[k];
    }
  }
  for (int k = ODL_N - 1; k >= 0; --k) {
    double xr = br[k], xi = bi[k];
    for (int j = k + 1; j < ODL_N; ++j) {
      xr -= F.ar[k][j] * br[j] - F.ai[k][j] * bi[j];
      xi -= F.ar[k][j] * bi[j] + F.ai[k][j] * br[j];
    }
    br[k] = xr * F.ir[k] - xi * F.ii[k];
    bi[k] = xr * F.ii[k] + xi * F.ir[k];
  }
}

#endif  // ODL_SMALL

#ifdef ODL_HOST_HARNESS
static long long odl_dbg[8];
#endif
struct OdlRadauAux {
  double Q[3][ODL_N];            // collocation polynomial of the last accepted step: y(t0 + x h) = y0 + Q0 x + Q1 x^2 + Q2 x^3
  double h_old;
  float err_old;
  bool have_sol, rejected;
  __device__ __forceinline__ void reset() { have_sol = false; rejected = false; h_old = 0.0; err_old = 0.f; }
};
struct OdlNoAux { __device__ __forceinline__ void reset() {} };

template <class Sink>
__device__ __forceinline__ void odl_radau5_attempt(OdlStepper& st, OdlRadauAux& ax, const double (&p)[ODL_P],
                                                   const OdlShared& S, const OdlData& D, const OdlOpts& O, Sink& sink) {
  // transformation matrices of RADAU5 (eigen-decomposition of A^-1), TI = T^-1
  const double T00 = 0.09443876248897524, T01 = -0.1412552950209542, T02 = 0.03002919410514742;
  const double T10 = 0.2502131229653333, T11 = 0.20412935229379994, T12 = -0.3829421127572619;
  // T2x = (1, 1, 0)
  const double I00 = 4.178718591551904, I01 = 0.32768282076106237, I02 = 0.5233764454994495;
  const double I10 = -4.178718591551904, I11 = -0.32768282076106237, I12 = 0.47662355450055044;
  const double I20 = 0.5028726349457868, I21 = -2.571926949855605, I22 = 0.5960392048282249;
  const double E0 = -10.048809399827414, E1 = 1.382142733160748, E2 = -0.3333333333333333;   // (-13 -+ 7 sqrt6)/3, -1/3
  const double P00 = 10.048809399827414, P01 = -25.62959144707664, P02 = 15.580782047249224;
  const double P10 = -1.382142733160748, P11 = 10.296258113743303, P12 = -8.914115380582556;
  const double P20 = 0.3333333333333333, P21 = -2.6666666666666665, P22 = 3.3333333333333335;

  const double t = st.t;
  double h = st.h;
  bool last = false;
  if ((t + 1.01 * h - st.tend) > 0.0) { h = st.tend - t; last = true; }
  ++st.nsteps;
  const double mr = ODL_RAD_MU / h, ca = ODL_RAD_ALPHA / h, cb = -ODL_RAD_BETA / h;   // M_complex = ca + i cb
  OdlLU R;
  OdlCLU Cx;
  {
    double J[ODL_N][ODL_N];
    odl_jac(st.y, t, p, J);
ODL_UNROLL
    for (int i = 0; i < ODL_N; ++i)
ODL_UNROLL
      for (int j = 0; j < ODL_N; ++j) {
        R.a[i][j] = ((i == j) ? mr : 0.0) - J[i][j];
        Cx.ar[i][j] = ((i == j) ? ca : 0.0) - J[i][j];
        Cx.ai[i][j] = (i == j) ? cb : 0.0;
      }
  }
  odl_lu_factor(R);
  odl_clu_factor(Cx);

  // RADAU5's tolerance transformation (radau5.f: RTOL' = 0.1 RTOL^(2/3), ATOL' = RTOL' ATOL/RTOL): the error
  // estimate belongs to the embedded third-order formula while the fifth-order solution is what advances.
  const double rtol = 0.1 * exp2(0.6666666666666666 * log2(O.rtol));
  const double atol = rtol * (O.atol / O.rtol);
  float rsc[ODL_N];
ODL_UNROLL
  for (int i = 0; i < ODL_N; ++i) rsc[i] = __frcp_rn((float)(atol + rtol * fabs(st.y[i])));
  const float newton_tol = fmaxf((float)(10.0 * 2.220446049250313e-16 / rtol), fminf(0.03f, sqrtf((float)rtol)));

  double Z0[ODL_N], Z1[ODL_N], Z2[ODL_N], W0[ODL_N], W1[ODL_N], W2[ODL_N];
  if (ax.have_sol) {
    const double r = h / ax.h_old;
    const double x0 = 1.0 + r * ODL_RAD_C0, x1 = 1.0 + r * ODL_RAD_C1, x2 = 1.0 + r;
ODL_UNROLL
    for (int j = 0; j < ODL_N; ++j) {
      const double q0 = ax.Q[0][j], q1 = ax.Q[1][j], q2 = ax.Q[2][j];
      const double one = q0 + q1 + q2;
      Z0[j] = x0 * (q0 + x0 * (q1 + x0 * q2)) - one;
      Z1[j] = x1 * (q0 + x1 * (q1 + x1 * q2)) - one;
      Z2[j] = x2 * (q0 + x2 * (q1 + x2 * q2)) - one;
    }
  } else {
ODL_UNROLL
    for (int j = 0; j < ODL_N; ++j) { Z0[j] = 0.0; Z1[j] = 0.0; Z2[j] = 0.0; }
  }
ODL_UNROLL
  for (int j = 0; j < ODL_N; ++j) {
    W0[j] = I00 * Z0[j] + I01 * Z1[j] + I02 * Z2[j];
    W1[j] = I10 * Z0[j] + I11 * Z1[j] + I12 * Z2[j];
    W2[j] = I20 * Z0[j] + I21 * Z1[j] + I22 * Z2[j];
  }
  bool converged = false;
  int n_iter = 0;
  float rate = -1.f, dwn_old = -1.f;
  for (int k = 0; k < ODL_RAD_NEWTON; ++k) {
    ++n_iter;
    double F0[ODL_N], F1[ODL_N], F2[ODL_N], yt[ODL_N];
ODL_UNROLL
    for (int j = 0; j < ODL_N; ++j) yt[j] = st.y[j] + Z0[j];
    odl_rhs(yt, t + ODL_RAD_C0 * h, p, F0);
ODL_UNROLL
    for (int j = 0; j < ODL_N; ++j) yt[j] = st.y[j] + Z1[j];
    odl_rhs(yt, t + ODL_RAD_C1 * h, p, F1);
ODL_UNROLL
    for (int j = 0; j < ODL_N; ++j) yt[j] = st.y[j] + Z2[j];
    odl_rhs(yt, t + h, p, F2);
    bool fin = true;
    double dr[ODL_N], dcr[ODL_N], dci[ODL_N];
ODL_UNROLL
    for (int j = 0; j < ODL_N; ++j) {
      fin = fin && odl_finite(F0[j]) && odl_finite(F1[j]) && odl_finite(F2[j]);
      dr[j] = I00 * F0[j] + I01 * F1[j] + I02 * F2[j] - mr * W0[j];
      dcr[j] = I10 * F0[j] + I11 * F1[j] + I12 * F2[j] - (ca * W1[j] - cb * W2[j]);
      dci[j] = I20 * F0[j] + I21 * F1[j] + I22 * F2[j] - (cb * W1[j] + ca * W2[j]);
    }
    if (!fin) break;
    odl_lu_solve(R, dr);
    odl_clu_solve(Cx, dcr, dci);
    float nrm = 0.f;
ODL_UNROLL
    for (int j = 0; j < ODL_N; ++j) {
      const float a = (float)dr[j] * rsc[j], b = (float)dcr[j] * rsc[j], c = (float)dci[j] * rsc[j];
      nrm += a * a + b * b + c * c;
    }
    const float dwn = sqrtf(nrm * (1.0f / (3 * ODL_N)));
    if (!(dwn == dwn)) break;
    if (dwn_old >= 0.f) rate = dwn / dwn_old;
    if (rate >= 0.f && (rate >= 1.f || __powf(rate, (float)(ODL_RAD_NEWTON - k)) / (1.f - rate) * dwn > newton_tol)) break;
ODL_UNROLL
    for (int j = 0; j < ODL_N; ++j) {
      W0[j] += dr[j]; W1[j] += dcr[j]; W2[j] += dci[j];
      Z0[j] = T00 * W0[j] + T01 * W1[j] + T02 * W2[j];
      Z1[j] = T10 * W0[j] + T11 * W1[j] + T12 * W2[j];
      Z2[j] = W0[j] + W1[j];
    }
    if (dwn == 0.f || (rate >= 0.f && rate / (1.f - rate) * dwn < newton_tol)) { converged = true; break; }
    dwn_old = dwn;
  }
  const double hmin = 4.0 * 2.220446049250313e-16 * fmax(fabs(t), fabs(st.tend));
#ifdef ODL_HOST_HARNESS
  odl_dbg[0] += 1; odl_dbg[1] += n_iter; if (!converged) odl_dbg[2] += 1;
#endif
  if (!converged) {
    st.h = 0.5 * h;
    if (!(st.h > hmin)) st.status = ODL_HUNDERFLOW;
    if (st.nsteps >= O.max_steps && st.status == ODL_OK) st.status = ODL_MAXSTEPS;
    return;
  }
  double yn[ODL_N], ze[ODL_N], er[ODL_N];
  const double rh = 1.0 / h;
  float errsq = 0.f;
  bool finite_all = true;
ODL_UNROLL
  for (int j = 0; j < ODL_N; ++j) {
    yn[j] = st.y[j] + Z2[j];
    ze[j] = (E0 * Z0[j] + E1 * Z1[j] + E2 * Z2[j]) * rh;
    er[j] = st.k1[j] + ze[j];
    finite_all = finite_all && odl_finite(yn[j]);
  }
  odl_lu_solve(R, er);
  float rs2[ODL_N];
ODL_UNROLL
  for (int j = 0; j < ODL_N; ++j) {
    rs2[j] = __frcp_rn((float)(atol + rtol * fmax(fabs(st.y[j]), fabs(yn[j]))));
    const float a = (float)er[j] * rs2[j];
    errsq += a * a;
  }
  float err = sqrtf(errsq * (1.0f / ODL_N));
  const float safety = 0.9f * (2 * ODL_RAD_NEWTON + 1) / (float)(2 * ODL_RAD_NEWTON + n_iter);
  if (ax.rejected && err > 1.0f) {
    double yt[ODL_N], fe[ODL_N];
ODL_UNROLL
    for (int j = 0; j < ODL_N; ++j) yt[j] = st.y[j] + er[j];
    odl_rhs(yt, t, p, fe);
ODL_UNROLL
    for (int j = 0; j < ODL_N; ++j) er[j] = fe[j] + ze[j];
    odl_lu_solve(R, er);
    errsq = 0.f;
ODL_UNROLL
    for (int j = 0; j < ODL_N; ++j) { const float a = (float)er[j] * rs2[j]; errsq += a * a; }
    err = sqrtf(errsq * (1.0f / ODL_N));
  }
  // Gustafsson: factor = min(1, h/h_old (err_old/err)^(1/4)) err^(-1/4)
  float mult = 1.f;
  if (ax.have_sol && ax.err_old > 0.f && err > 0.f) mult = (float)(h / ax.h_old) * __powf(ax.err_old / err, 0.25f);
  float factor = fminf(1.f, mult) * __powf(err, -0.25f);
  if (!(err > 0.f)) factor = 10.f;
  if (!(err <= 1.0f) || !finite_all) {
#ifdef ODL_HOST_HARNESS
    odl_dbg[3] += 1;
#endif
    double hnew = (err == err && err < 3.0e38f) ? h * (double)fmaxf(0.2f, fminf(0.95f, safety * factor)) : 0.2 * h;
    ax.rejected = true;
    st.h = hnew;
    if (!(hnew > hmin)) st.status = ODL_HUNDERFLOW;
    if (st.nsteps >= O.max_steps && st.status == ODL_OK) st.status = ODL_MAXSTEPS;
    return;
  }
  // ---- accepted ----
  const double tnew = last ? st.tend : t + h;
ODL_UNROLL
  for (int j = 0; j < ODL_N; ++j) {
    ax.Q[0][j] = P00 * Z0[j] + P10 * Z1[j] + P20 * Z2[j];
    ax.Q[1][j] = P01 * Z0[j] + P11 * Z1[j] + P21 * Z2[j];
    ax.Q[2][j] = P02 * Z0[j] + P12 * Z1[j] + P22 * Z2[j];
  }
  while (st.slot < D.n_slot && S.slot_t[st.slot] <= tnew) {
    const double x = (S.slot_t[st.slot] - t) * rh;
    double yi[ODL_N];
ODL_UNROLL
    for (int j = 0; j < ODL_N; ++j) yi[j] = st.y[j] + x * (ax.Q[0][j] + x * (ax.Q[1][j] + x * ax.Q[2][j]));
    sink(st.slot, yi);
    ++st.slot;
  }
ODL_UNROLL
  for (int j = 0; j < ODL_N; ++j) st.y[j] = yn[j];
  odl_rhs(st.y, tnew, p, st.k1);
  st.t = tnew;
  ax.h_old = h; ax.err_old = fmaxf(err, 1e-10f); ax.have_sol = true; ax.rejected = false;
  st.h = h * (double)fminf(10.f, fmaxf(0.2f, safety * factor));
  if (st.nsteps >= O.max_steps && st.slot < D.n_slot && st.status == ODL_OK) st.status = ODL_MAXSTEPS;
}

// ------------------------------------------------------------------------------------------------
// Variable-order (1..5), quasi-constant step size BDF/NDF in backward differences (Shampine & Reichelt,
// "The MATLAB ODE suite", ode15s; Byrne & Hindmarsh 1975 for the step-size change of the difference
// array).  This is the method family of LSODA's stiff branch -- what the reference itself runs on these
// systems (Framework.py:656) -- and the cheapest per step of the stiff steppers here: one real n x n LU
// that is kept while the step size stands (it changes at most every order+1 accepted steps), two or
// three simplified-Newton iterations of (one RHS + one back-substitution) per step.  Radau5 needs half as
// many steps on the hard corner of the demo priors but ~6x the instructions per step (complex LU, three
// stages), and the tail pass of the sweep is bound by the latency of its longest system.
//
// Storage: D[0] = y lives in st.y; the differences D[1..order+2] are kept REVERSED, E[k] = D[order+2-k],
// so that everything touched on every step (d = D[order+1], D[order+2], the running sums) sits at
// compile-time register indices whatever the current order is; only a step-size/order change (rare)
// re-indexes, through a small local scratch array.
// ------------------------------------------------------------------------------------------------
#define ODL_BDF_MAXORD 5
#define ODL_BDF_NEWTON 4
#ifndef ODL_BDF_NEWTON_TOL
#define ODL_BDF_NEWTON_TOL 0.03f
#endif
__constant__ double ODL_BDF_GAMMA[8] = {0.0, 1.0, 1.5, 11.0 / 6.0, 25.0 / 12.0, 137.0 / 60.0, 0.0, 0.0};   // [(q+2-k) & 7]
// alpha_q = (1 - kappa_q) gamma_q, kappa = (0, -0.1850, -1/9, -0.0823, -0.0415, 0)
__constant__ double ODL_BDF_RALPHA[ODL_BDF_MAXORD + 1] = {
    0.0, 1.0 / (1.1850 * 1.0), 1.0 / ((1.0 + 1.0 / 9.0) * 1.5), 1.0 / (1.0823 * (11.0 / 6.0)),
    1.0 / (1.0415 * (25.0 / 12.0)), 1.0 / (137.0 / 60.0)};
// error_const_q = kappa_q gamma_q + 1/(q+1)
__constant__ float ODL_BDF_ERRC[ODL_BDF_MAXORD + 1] = {
    1.0f, (float)(-0.1850 * 1.0 + 0.5), (float)(-1.5 / 9.0 + 1.0 / 3.0), (float)(-0.0823 * (11.0 / 6.0) + 0.25),
    (float)(-0.0415 * (25.0 / 12.0) + 0.2), (float)(1.0 / 6.0)};
__constant__ double ODL_BDF_INV[ODL_BDF_MAXORD + 2] = {0.0, 1.0, 0.5, 1.0 / 3.0, 0.25, 0.2, 1.0 / 6.0};

struct OdlBdfAux {
  double E[ODL_BDF_MAXORD + 2][ODL_N];   // E[k] = D[order+2-k], k = 0 .. order+1
  OdlLU lu;                              // LU of I - (h/alpha_q) J
  int order, n_equal;
  int phase;                             // calls of odl_bdf_attempt for this system (waiting calls included)
  int pend_order;                        // a step-size / order change that has been decided but not applied yet:
  double pend_factor;                    //   new order, ratio new/old step (0 = none)
  float crate;                           // Newton contraction rate carried from step to step (reset with the LU)
  bool have_lu, started;
  __device__ __forceinline__ void reset() {
    started = false; have_lu = false; order = 1; n_equal = 0; crate = 1.f; phase = 0; pend_order = 1; pend_factor = 0.0;
  }
};

// D[1..q_new] <- (R(factor) U)^T D[1..q_new]: the differences of the same interpolating polynomial on a grid
// of spacing factor*h (scipy's change_D; R[i][m] = prod_{l<=i} (l-1-factor*m)/l, U = R(1) = (-1)^m C(j,m)).
__device__ __forceinline__ void odl_bdf_change_D(OdlBdfAux& ax, int q_old, int q_new, double factor) {
  double L[ODL_BDF_MAXORD + 2][ODL_N];
#pragma unroll
  for (int k = 0; k < ODL_BDF_MAXORD + 2; ++k)
ODL_UNROLL
    for (int c = 0; c < ODL_N; ++c) L[k][c] = ax.E[k][c];
  double Dn[ODL_BDF_MAXORD + 1][ODL_N];          // natural order, [1..5]
#pragma unroll
  for (int i = 1; i <= ODL_BDF_MAXORD; ++i) {
    const int src = (i <= q_new) ? q_old + 2 - i : 0;
ODL_UNROLL
    for (int c = 0; c < ODL_N; ++c) Dn[i][c] = (i <= q_new) ? L[src][c] : 0.0;
  }
  double tmp[ODL_BDF_MAXORD + 1][ODL_N];
#pragma unroll
  for (int m = 1; m <= ODL_BDF_MAXORD; ++m) {
    double r = 1.0;
ODL_UNROLL
    for (int c = 0; c < ODL_N; ++c) tmp[m][c] = 0.0;
    const double fm = factor * (double)m;
#pragma unroll
    for (int i = 1; i <= ODL_BDF_MAXORD; ++i) {
      r *= ((double)(i - 1) - fm) * ODL_BDF_INV[i];
      // rows beyond q_new hold zeros, columns beyond q_new are never read back
ODL_UNROLL
      for (int c = 0; c < ODL_N; ++c) tmp[m][c] += r * Dn[i][c];
    }
  }
  // U^T: D'[j] = sum_{m<=j} (-1)^m C(j,m) tmp[m]
  const double U[ODL_BDF_MAXORD + 1][ODL_BDF_MAXORD + 1] = {{0, 0, 0, 0, 0, 0},  {0, -1, -2, -3, -4, -5}, {0, 0, 1, 3, 6, 10},
                                                            {0, 0, 0, -1, -4, -10}, {0, 0, 0, 0, 1, 5},  {0, 0, 0, 0, 0, -1}};
#pragma unroll
  for (int j = 1; j <= ODL_BDF_MAXORD; ++j)
ODL_UNROLL
    for (int c = 0; c < ODL_N; ++c) {
      double acc = 0.0;
#pragma unroll
      for (int m = 1; m <= j; ++m) acc += U[m][j] * tmp[m][c];
      L[j][c] = acc;                             // scratch reused: natural index j
    }
#pragma unroll
  for (int k = 0; k < ODL_BDF_MAXORD + 2; ++k) {
    const bool live = (k >= 2) && (k <= q_new + 1);
    const int src = live ? q_new + 2 - k : 1;
ODL_UNROLL
    for (int c = 0; c < ODL_N; ++c) ax.E[k][c] = live ? L[src][c] : 0.0;
  }
}

__device__ __forceinline__ float odl_bdf_rms(const double (&v)[ODL_N], const double (&rs)[ODL_N]) {
  double s = 0.0;
ODL_UNROLL
  for (int c = 0; c < ODL_N; ++c) { const double a = v[c] * rs[c]; s += a * a; }
  return odl_sqrt_approx((float)s * (1.0f / ODL_N));
}

template <class Sink>
__device__ __forceinline__ void odl_bdf_attempt(OdlStepper& st, OdlBdfAux& ax, const double (&p)[ODL_P], const OdlShared& S,
                                                const OdlData& D, const OdlOpts& O, Sink& sink) {
  const double t = st.t;
  const double hmin = 4.0 * 2.220446049250313e-16 * fmax(fabs(t), fabs(st.tend));
  if (!ax.started) {
    // order 1 start: D[1] = h f(y0); initial step as in Hairer's hinit with the exponent of a first-order method
    double rs0[ODL_N], y1[ODL_N], f1[ODL_N];
ODL_UNROLL
    for (int c = 0; c < ODL_N; ++c) rs0[c] = odl_rcp_approx(O.atol + O.rtol * fabs(st.y[c]));
    const float d0 = odl_bdf_rms(st.y, rs0), d1 = odl_bdf_rms(st.k1, rs0);
    double h0 = (d0 < 1e-5f || d1 < 1e-5f) ? 1e-6 : 0.01 * (double)(d0 / d1);
    h0 = fmin(h0, st.tend - t);
ODL_UNROLL
    for (int c = 0; c < ODL_N; ++c) y1[c] = st.y[c] + h0 * st.k1[c];
    odl_rhs(y1, t + h0, p, f1);
ODL_UNROLL
    for (int c = 0; c < ODL_N; ++c) f1[c] -= st.k1[c];
    const float d2 = odl_bdf_rms(f1, rs0) / (float)h0;
    const float dm = fmaxf(d1, d2);
    const double h1 = (dm <= 1e-15f) ? fmax(1e-6, h0 * 1e-3) : (double)sqrtf(0.01f / dm);
    double h = fmin(fmin(100.0 * h0, h1), st.tend - t);
    if (O.h0 > 0.0) h = fmin(O.h0, st.tend - t);
    if (!(h > 0.0) || !odl_finite(h)) h = 1e-6 * (st.tend - t);
    st.h = h;
    ax.order = 1; ax.n_equal = 0; ax.have_lu = false; ax.started = true;
#pragma unroll
    for (int k = 0; k < ODL_BDF_MAXORD + 2; ++k)
ODL_UNROLL
      for (int c = 0; c < ODL_N; ++c) ax.E[k][c] = (k == 2) ? h * st.k1[c] : 0.0;
  }
  // Every change of the difference array -- order / step-size selection, a rejected step, a Newton failure, the
  // clamp onto t_end -- is DECIDED where it arises (st.h already holds the new step) and APPLIED here, at the one
  // place odl_bdf_change_D is inlined, on calls whose number is a multiple of 3.  The lanes of a warp share their
  // call count, so the ~330 instructions are issued every third warp step for all lanes that need them, instead of on
  // a quarter of all warp steps for the one lane that just rejected; a lane with a pending change waits (at most two
  // calls).  The rule looks at this system only: a solve does not depend on its warp mates.
  ++ax.phase;
  bool force = false;
  if (ax.pend_factor == 0.0 && (t + 1.01 * st.h - st.tend) > 0.0 && (st.tend - t) != st.h) {
    ax.pend_factor = (st.tend - t) / st.h; ax.pend_order = ax.order; st.h = st.tend - t;      // once per solve: no waiting
    force = true;
  }
  if (ax.pend_factor != 0.0) {
    if (!force && (ax.phase % 3) != 0) return;
    odl_bdf_change_D(ax, ax.order, ax.pend_order, ax.pend_factor);
    ax.order = ax.pend_order; ax.pend_factor = 0.0; ax.n_equal = 0; ax.have_lu = false;
  }
  const int q = ax.order;
  const double h = st.h;
  const bool last = (t + 1.01 * h - st.tend) > 0.0;              // after the clamp h == t_end - t exactly
  ++st.nsteps;
  const double tn = last ? st.tend : t + h;
  const double ralpha = ODL_BDF_RALPHA[q];
  const double cc = h * ralpha;

  // predictor y_pred = sum_{i<=q} D[i], psi = sum_{i=1..q} gamma_i D[i] / alpha_q
  double yp[ODL_N], psi[ODL_N], rs[ODL_N];
ODL_UNROLL
  for (int c = 0; c < ODL_N; ++c) { yp[c] = st.y[c]; psi[c] = 0.0; }
  // rows k > q+1 of E are zero (change_D clears them, the update below never touches them): no predicate needed,
  // and the gamma of a dead row may be anything finite
#pragma unroll
  for (int k = 2; k < ODL_BDF_MAXORD + 2; ++k) {
    const double g = ODL_BDF_GAMMA[(q + 2 - k) & 7];
ODL_UNROLL
    for (int c = 0; c < ODL_N; ++c) { yp[c] += ax.E[k][c]; psi[c] += g * ax.E[k][c]; }
  }
ODL_UNROLL
  for (int c = 0; c < ODL_N; ++c) { psi[c] *= ralpha; rs[c] = odl_rcp_approx(O.atol + O.rtol * fabs(yp[c])); }

  bool fresh = false;
  if (!ax.have_lu) {
    odl_jac(yp, tn, p, ax.lu.a);
ODL_UNROLL
    for (int i = 0; i < ODL_N; ++i)
ODL_UNROLL
      for (int j = 0; j < ODL_N; ++j) ax.lu.a[i][j] = ((i == j) ? 1.0 : 0.0) - cc * ax.lu.a[i][j];
    odl_lu_factor(ax.lu);
    ax.have_lu = true; fresh = true;
    ax.crate = 1.f;
  }
  // scipy's BDF asks for min(0.03, sqrt(rtol)) here; LSODA / CVODE accept a Newton error of a few per cent of
  // the local error tolerance whatever rtol is, which halves the iterations at tight tolerances
  const float newton_tol = fmaxf((float)(10.0 * 2.220446049250313e-16 / O.rtol), ODL_BDF_NEWTON_TOL);
  double d[ODL_N], yk[ODL_N];
ODL_UNROLL
  for (int c = 0; c < ODL_N; ++c) { d[c] = 0.0; yk[c] = yp[c]; }
  bool converged = false;
  int n_iter = 0;
  float rate = -1.f, dn_old = -1.f;
  for (int k = 0; k < ODL_BDF_NEWTON; ++k) {
    ++n_iter;
    double f[ODL_N];
    odl_rhs(yk, tn, p, f);
    double fsum = 0.0;
ODL_UNROLL
    for (int c = 0; c < ODL_N; ++c) { fsum += fabs(f[c]); f[c] = cc * f[c] - psi[c] - d[c]; }
    if (!odl_finite(fsum)) break;
    odl_lu_solve(ax.lu, f);
    const float dn = odl_bdf_rms(f, rs);
    if (!(dn == dn)) break;
    // contraction rate: measured from the second iteration on; on the first one, the rate the previous steps saw
    // with this very LU (as LSODA / CVODE do).  est = rate/(1-rate) dn bounds the remaining Newton error.
    const bool measured = dn_old >= 0.f;
    if (measured) rate = dn * __frcp_rn(dn_old);
    const float r_use = measured ? rate : ((ax.crate < 0.3f) ? ax.crate : -1.f);
    if (measured && !(rate < 0.9f)) break;                      // diverging or too slow: new Jacobian / smaller step
ODL_UNROLL
    for (int c = 0; c < ODL_N; ++c) { yk[c] += f[c]; d[c] += f[c]; }
    if (dn == 0.f || (r_use >= 0.f && r_use * dn < newton_tol * (1.f - r_use))) { converged = true; break; }
    dn_old = dn;
  }
  if (rate >= 0.f) ax.crate = fmaxf(0.2f * ax.crate, rate);
#ifdef ODL_HOST_HARNESS
  odl_dbg[0] += 1; odl_dbg[1] += n_iter; if (!converged) odl_dbg[2] += 1; if (fresh) odl_dbg[4] += 1;
#endif
  if (!converged) {
    ax.have_lu = false;
    if (fresh) {                                                // a current Jacobian did not help: halve the step
      ax.pend_factor = 0.5; ax.pend_order = q;
      ax.n_equal = 0;
      st.h = 0.5 * h;
      if (!(st.h > hmin)) st.status = ODL_HUNDERFLOW;
    }
    if (st.nsteps >= O.max_steps && st.status == ODL_OK) st.status = ODL_MAXSTEPS;
    return;
  }
  const float safety = 0.9f * (2 * ODL_BDF_NEWTON + 1) / (float)(2 * ODL_BDF_NEWTON + n_iter);
  double ysum = 0.0;
ODL_UNROLL
  for (int c = 0; c < ODL_N; ++c) { ysum += fabs(yk[c]); rs[c] = odl_rcp_approx(O.atol + O.rtol * fabs(yk[c])); }
  const bool finite_all = odl_finite(ysum);
  const float err = ODL_BDF_ERRC[q] * odl_bdf_rms(d, rs);
  if (!(err <= 1.0f) || !finite_all) {
#ifdef ODL_HOST_HARNESS
    odl_dbg[3] += 1;
#endif
    const float fac = (err == err && err < 3.0e38f && finite_all)
                          ? fmaxf(0.2f, safety * exp2f(-__log2f(err) * (float)ODL_BDF_INV[q + 1])) : 0.2f;
    ax.pend_factor = (double)fac; ax.pend_order = q;
    ax.n_equal = 0; ax.have_lu = false;
    st.h = h * (double)fac;
    if (!(st.h > hmin)) st.status = ODL_HUNDERFLOW;
    if (st.nsteps >= O.max_steps && st.status == ODL_OK) st.status = ODL_MAXSTEPS;
    return;
  }
  // ---- accepted: D[q+2] = d - D[q+1], D[q+1] = d, D[i] += D[i+1] (i = q .. 0) ----
  ++ax.n_equal;
ODL_UNROLL
  for (int c = 0; c < ODL_N; ++c) { ax.E[0][c] = d[c] - ax.E[1][c]; ax.E[1][c] = d[c]; }
#pragma unroll
  for (int k = 2; k < ODL_BDF_MAXORD + 2; ++k)
    if (k <= q + 1) {
ODL_UNROLL
      for (int c = 0; c < ODL_N; ++c) ax.E[k][c] += ax.E[k - 1][c];
    }
  // dense output: y(ts) = y_new + sum_{j=1..q} D[j] prod_{i<j} (ts - (tn - i h)) / (h (1+i))
  if (st.slot < D.n_slot && S.slot_t[st.slot] <= tn) {
    const double rh = 1.0 / h;
    do {
      const double x0 = (S.slot_t[st.slot] - tn) * rh;
      double yi[ODL_N];
ODL_UNROLL
      for (int c = 0; c < ODL_N; ++c) yi[c] = yk[c];
      double pr = 1.0;
#pragma unroll
      for (int k = ODL_BDF_MAXORD + 1; k >= 2; --k)
        if (k <= q + 1) {
          const int i = q + 1 - k;
          pr *= (x0 + (double)i) * ODL_BDF_INV[i + 1];
ODL_UNROLL
          for (int c = 0; c < ODL_N; ++c) yi[c] += pr * ax.E[k][c];
        }
      sink(st.slot, yi);
      ++st.slot;
    } while (st.slot < D.n_slot && S.slot_t[st.slot] <= tn);
  }
ODL_UNROLL
  for (int c = 0; c < ODL_N; ++c) st.y[c] = yk[c];
  st.t = tn;
  if (st.nsteps >= O.max_steps && st.slot < D.n_slot && st.status == ODL_OK) st.status = ODL_MAXSTEPS;
  // The order/step-size selection is taken on calls whose successor is a multiple of 3 (orders 1-2) or 6 (orders
  // 3-5), so that its change of the difference array is applied on an aligned call (see the top of this function).
  if (ax.n_equal < q + 1 || st.slot >= D.n_slot || ((ax.phase + 1) % (q >= 3 ? 6 : 3)) != 0) return;
  // ---- order / step-size selection (every q+1 equal steps) ----
  float f_m = 0.f, f_p = 0.f;
  const float f_0 = (err > 0.f) ? exp2f(-__log2f(err) * (float)ODL_BDF_INV[q + 1]) : 3.0e38f;
  if (q > 1) {
    const float e = ODL_BDF_ERRC[q - 1] * odl_bdf_rms(ax.E[2], rs);
    f_m = (e > 0.f) ? exp2f(-__log2f(e) * (float)ODL_BDF_INV[q]) : 3.0e38f;
  }
  if (q < ODL_BDF_MAXORD) {
    const float e = ODL_BDF_ERRC[q + 1] * odl_bdf_rms(ax.E[0], rs);
    f_p = (e > 0.f) ? exp2f(-__log2f(e) * (float)ODL_BDF_INV[q + 2]) : 3.0e38f;
  }
  int qn = q - 1;                                               // np.argmax: first maximum of (f_m, f_0, f_p)
  float fbest = f_m;                                            // (0 when q == 1, and f_0 >= 1 on an accepted step)
  if (f_0 > fbest) { fbest = f_0; qn = q; }
  if (f_p > fbest) { fbest = f_p; qn = q + 1; }
  const float fac = fminf(10.f, safety * fbest);
  if (qn == q && fac >= 0.9f && fac < 1.2f) {
    // same order, about the same step: keep the grid and the LU (LSODA / CVODE have the same dead band) and look
    // again in a few steps
    ax.n_equal = (q > 2) ? q - 2 : 0;
    return;
  }
  double hn = h * (double)fac;
  if ((st.t + 1.01 * hn - st.tend) > 0.0) hn = st.tend - st.t;  // the clamp onto t_end, folded into this change
  ax.pend_factor = hn / h; ax.pend_order = qn;                  // applied at the top of the next call (an aligned one)
  ax.n_equal = 0;
  st.h = hn;
}

template <int SOLVER> struct OdlAuxOf { typedef OdlNoAux type; };
template <> struct OdlAuxOf<2> { typedef OdlBdfAux type; };
template <> struct OdlAuxOf<3> { typedef OdlRadauAux type; };
template <> struct OdlAuxOf<4> { typedef OdlBdfAux type; };

// solver dispatch: 0 = DOPRI5, 1 = ROS23, 2 = per-solve choice (DOPRI5 until it gives up -- step budget or Hairer's
// stiffness test --, then the same solve again on BDF: what LSODA's method switch does for the reference), 3 = Radau5,
// 4 = BDF
template <int SOLVER, class Sink>
__device__ __forceinline__ void odl_attempt(OdlStepper& st, typename OdlAuxOf<SOLVER>::type& ax, const double (&p)[ODL_P],
                                            const OdlShared& S, const OdlData& D, const OdlOpts& O, Sink& sink, bool use_alt) {
  if constexpr (SOLVER == 0) odl_dopri5_attempt(st, p, S, D, O, sink);
  else if constexpr (SOLVER == 1) odl_ros23_attempt(st, p, S, D, O, sink);
  else if constexpr (SOLVER == 3) odl_radau5_attempt(st, ax, p, S, D, O, sink);
  else if constexpr (SOLVER == 4) odl_bdf_attempt(st, ax, p, S, D, O, sink);
  else { if (use_alt) odl_bdf_attempt(st, ax, p, S, D, O, sink); else odl_dopri5_attempt(st, p, S, D, O, sink); }
}

#ifndef ODL_HOST_HARNESS
__device__ __forceinline__ long long odl_globaltimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// fetch `want` lanes' worth of indices from a global counter with one atomic per warp
__device__ __forceinline__ long long odl_fetch(unsigned long long* counter, bool want, int lane) {
  const unsigned m = __ballot_sync(ODL_FULL, want);
  if (m == 0) return -1;
  const int leader = __ffs(m) - 1;
  unsigned long long base = 0;
  if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(m));
  base = ((unsigned long long)__shfl_sync(ODL_FULL, (int)(base >> 32), leader) << 32) |
         (unsigned int)__shfl_sync(ODL_FULL, (int)(base & 0xffffffffu), leader);
  return (long long)(base + __popc(m & ((1u << lane) - 1)));
}

// loads that must observe what other kernels / CTAs published with atomics + __threadfence (L2, in program order)
__device__ __forceinline__ int odl_ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long odl_ld_acquire(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// ------------------------------------------------------------------------------------------------
// Cost ordering of a sweep (see OdlOrderArgs): histogram of the keys, start offsets, scatter of the row numbers.
// The order inside a bin is whatever the atomics give; it only decides when a system runs, never its result.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int odl_order_bin(const OdlData& D, const double* theta_row) {
  double p[ODL_P], y[ODL_N];
ODL_UNROLL
  for (int q = 0; q < ODL_P; ++q) p[q] = theta_row[q];
ODL_UNROLL
  for (int i = 0; i < ODL_N; ++i) y[i] = D.y0[i];             // a sweep starts from istates (see odl_init_system)
  double J[ODL_N][ODL_N];
  odl_jac(y, D.t0, p, J);
  double nrm = 0.0;
ODL_UNROLL
  for (int i = 0; i < ODL_N; ++i) {
    double r = 0.0;
ODL_UNROLL
    for (int j = 0; j < ODL_N; ++j) r += fabs(J[i][j]);
    nrm = fmax(nrm, r);
  }
  const float key = (float)(nrm * (D.slot_t[D.n_slot - 1] - D.t0));
  if (!(key == key) || key > 3.0e38f) return ODL_ORDER_BINS - 1;                 // not finite: fail fast, first
  if (!(key > 0.f)) return 0;
  const int b = (int)floorf((__log2f(key) + 32.0f) * 4.0f);                      // quarter octaves over 2^-32 .. 2^32
  return b < 0 ? 0 : (b > ODL_ORDER_BINS - 1 ? ODL_ORDER_BINS - 1 : b);
}
#if ODL_HAS(13)
extern "C" __global__ void __launch_bounds__(256)
odl_order_key_kernel(const OdlData D, const OdlOrderArgs A) {
  __shared__ int h[ODL_ORDER_BINS];
  for (int i = threadIdx.x; i < ODL_ORDER_BINS; i += blockDim.x) h[i] = 0;
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < A.n; i += (long long)gridDim.x * blockDim.x) {
    const int b = odl_order_bin(D, A.theta + i * ODL_P);
    A.bins[i] = (unsigned char)b;
    atomicAdd(&h[b], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ODL_ORDER_BINS; i += blockDim.x) if (h[i]) atomicAdd(&A.hist[i], h[i]);
}
extern "C" __global__ void __launch_bounds__(ODL_ORDER_BINS)
odl_order_scan_kernel(const OdlOrderArgs A) {
  // exclusive prefix over the bins taken from the highest down: cursor[b] = number of rows in bins above b
  __shared__ int s[ODL_ORDER_BINS];
  const int b = threadIdx.x;
  s[b] = A.hist[ODL_ORDER_BINS - 1 - b];
  __syncthreads();
  for (int off = 1; off < ODL_ORDER_BINS; off <<= 1) {
    const int v = (b >= off) ? s[b - off] : 0;
    __syncthreads();
    s[b] += v;
    __syncthreads();
  }
  A.cursor[ODL_ORDER_BINS - 1 - b] = s[b] - A.hist[ODL_ORDER_BINS - 1 - b];
}
extern "C" __global__ void __launch_bounds__(256)
odl_order_scatter_kernel(const OdlOrderArgs A) {
  // one tile of rows per CTA: rank inside the tile from shared atomics, one global reservation per (tile, bin)
  __shared__ int h[ODL_ORDER_BINS];
  __shared__ int base[ODL_ORDER_BINS];
  const long long tile = (long long)blockDim.x * 8;
  for (long long t0 = (long long)blockIdx.x * tile; t0 < A.n; t0 += (long long)gridDim.x * tile) {
    for (int i = threadIdx.x; i < ODL_ORDER_BINS; i += blockDim.x) h[i] = 0;
    __syncthreads();
    int myb[8], myr[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const long long i = t0 + (long long)k * blockDim.x + threadIdx.x;
      myb[k] = -1; myr[k] = 0;
      if (i < A.n) { myb[k] = A.bins[i]; myr[k] = atomicAdd(&h[myb[k]], 1); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ODL_ORDER_BINS; i += blockDim.x) base[i] = h[i] ? atomicAdd(&A.cursor[i], h[i]) : 0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const long long i = t0 + (long long)k * blockDim.x + threadIdx.x;
      if (myb[k] >= 0) A.index[base[myb[k]] + myr[k]] = (int)i + A.row_base;
    }
    __syncthreads();
  }
}
// follows the last bulk launch of an AUTO sweep in its stream: the feed list of the stiff pass is complete
extern "C" __global__ void odl_feed_done_kernel(int* flag) { atomicExch(flag, 1); }
// precedes the first bulk launch in its stream: holds it back until the CTAs of the stiff pass (launched on another
// stream) are resident on SMs of their own -- a bulk grid that arrives first spreads over every SM and leaves none
// free for them until it ends.  Gives up after `spins` polls (~0.5 us each): late consumers cost time, not results.
extern "C" __global__ void odl_gate_kernel(const int* resident, int want, int spins) {
  for (int i = 0; i < spins; ++i) {
    if (odl_ld_acquire(resident) >= want) return;
    __nanosleep(500);
  }
}
#endif  // unit 13

// ------------------------------------------------------------------------------------------------
// Forward sweep: Framework.py:41-48 (_Fit_worker) for n parameter sets
// ------------------------------------------------------------------------------------------------
// Stiff pass of an AUTO sweep: continue where the DOPRI5 pass stopped (OdlSweepArgs.handover: t, next slot, y and the
// observation columns already staged) instead of integrating the row again from t0 -- at the projection check the rows
// have covered a third of the interval on average (35 % on the two_i priors), and the pass is bound by its throughput.
// The stepper starts as it does at t0: order 1, its own first step from f(t, y).  -> false when there is nothing to take.
template <int SOLVER>
__device__ __forceinline__ bool odl_take_over(OdlStepper& st, const double (&p)[ODL_P], const OdlSweepArgs& A, long long row,
                                              double* my_stage) {
  if constexpr (SOLVER == 0) return false;
  else {
    if (!A.handover || !A.index) return false;                  // only rows that come through the feed list have a record
    const double* rec = A.handover + (size_t)row * A.handover_stride;
    st.t = rec[0];
    st.slot = (int)rec[1];
ODL_UNROLL
    for (int i = 0; i < ODL_N; ++i) st.y[i] = rec[2 + i];
    odl_rhs(st.y, st.t, p, st.k1);
    for (int j = 0; j < st.slot * ODL_NOUT; ++j) my_stage[j] = rec[2 + ODL_N + j];
    return true;
  }
}

template <int SOLVER>
__device__ __forceinline__ void odl_sweep_body(const OdlData& D, const OdlOpts& O, const OdlSweepArgs& A) {
  const OdlShared S = odl_carve(odl_smem, D);
  odl_load_tables(S, D);
  const int lane = threadIdx.x & 31;
  double* my_stage = S.stage + (size_t)threadIdx.x * D.stage_stride;
  OdlStageSink sink; sink.stage = my_stage;
  const long long n = A.index_count ? (long long)(*A.index_count) : A.n;

  OdlStepper st;
  typename OdlAuxOf<SOLVER>::type ax;
  double p[ODL_P];
  long long sys = -1;            // slot in the work list
  long long row = -1;            // row of theta / outputs
  bool active = false, done = false;
  int fin_status = ODL_OK, fin_nsteps = 0;                       // of the system this lane finished (kept for (A))
  // a defined state for lanes without a system: in the DOPRI5 kernel every lane runs the step code (see (C)), and a
  // lane with slot == n_slot writes nothing
  st.status = ODL_OK; st.nsteps = 0; st.slot = D.n_slot; st.iasti = 0; st.nonsti = 0;
  st.t = 0.0; st.h = 0.0; st.tend = 0.0; st.lgfac = 0.f; st.last_rejected = false;
ODL_UNROLL
  for (int i = 0; i < ODL_N; ++i) { st.y[i] = 0.0; st.k1[i] = 0.0; }
ODL_UNROLL
  for (int q = 0; q < ODL_P; ++q) p[q] = 0.0;

  // first fetch.  O.lanes < 32 (latency-bound tail passes): only the first O.lanes lanes of a warp take systems, so
  // that the few long systems are spread over more warps -- a warp pays for the union of its lanes' branches.
  // O.lanes < 0 (the stiff pass AFTER the bulk pass: the feed is complete when this kernel starts): the entries are
  // spread evenly over all warps of the grid, each warp using as few lanes as that takes.
  int lanes_used = O.lanes;
  if (O.lanes < 0) {
    int count = A.index_count ? *A.index_count : 32 * (int)(gridDim.x * (blockDim.x >> 5));
    // a consumer launched behind the bulk pass (the feed is complete): what the first consumer has not taken yet
    if (A.feed_ticket) count = max(0, count - (int)min((unsigned long long)count, *(volatile unsigned long long*)A.feed_ticket));
    const int warps = (int)(gridDim.x * (blockDim.x >> 5));
    lanes_used = min(32, max(1, (count + warps - 1) / warps));
  }
  const bool lane_on = (lanes_used <= 0) || (lane < lanes_used);
  bool want = lane_on;
  // consumer of a feed another kernel is still writing (the stiff pass beside the bulk pass): a lane takes a ticket
  // and waits, `pending`, until that entry of index[] has landed or the producer is known to have finished short of it
  const bool consumer = A.feed_ticket != nullptr;
  bool pending = false;
  long long ticket = -1;
  unsigned int idle_spins = 0;
  if (consumer && A.resident && threadIdx.x == 0) atomicAdd(A.resident, 1);      // this CTA has its SM
#if ODL_TIMELINE
  if (consumer && A.timeline && threadIdx.x == 0) {                               // which SM, at the far end of the buffer
    unsigned int smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    A.timeline[3 * (A.n - 1 - blockIdx.x)] = (long long)smid + 1;
    A.timeline[3 * (A.n - 1 - blockIdx.x) + 1] = odl_globaltimer();
  }
#endif
  for (;;) {
    // ---- (A) finished lanes: cooperative score, write-back ----
    const bool fin = active && done;
    unsigned m = __ballot_sync(ODL_FULL, fin);
    double w_chi = 0.0, w_ss = 0.0;
    int w_nv = 0;
    while (m) {
      const int L = __ffs(m) - 1;
      m &= m - 1;
      const long long rowL = ((long long)__shfl_sync(ODL_FULL, (int)(row >> 32), L) << 32) |
                             (unsigned int)__shfl_sync(ODL_FULL, (int)(row & 0xffffffff), L);
      const int statL = __shfl_sync(ODL_FULL, fin_status, L);
      double chi, ss; int nv;
      double* pred_out = (A.pred && statL == ODL_OK) ? A.pred + rowL * D.n_obs : nullptr;
      // (unfinished solves are scored too: a branch around the inlined scorer costs more -- +2 % -- than their 2.5 %)
      odl_score(S, D, S.stage + (size_t)((threadIdx.x & ~31) + L) * D.stage_stride, lane, pred_out, chi, ss, nv);
      if (lane == L) { w_chi = chi; w_ss = ss; w_nv = nv; }     // kept for the write-back below, all finished lanes at once
    }
    if (fin) {
      int status = fin_status;
      double chi = w_chi;
      double r2 = fma(-w_ss, D.inv_sstot, 1.0);                   // 1 - ssres/sstot (stats.py:56), reciprocal from the host
      if (status != ODL_OK) { chi = __longlong_as_double(0x7ff8000000000000LL); r2 = chi; }
      else if (w_nv == 0) { chi = __longlong_as_double(0x7ff8000000000000LL); status |= ODL_ALLMASKED; }
      A.chi[row] = chi;                                          // r2 / status / nsteps only when the caller asked for them
      if (A.r2) A.r2[row] = r2;
      if (A.status) A.status[row] = status;
      if (A.nsteps) A.nsteps[row] = fin_nsteps;
#if ODL_TIMELINE
      if ((fin_status == ODL_MAXSTEPS || fin_status == ODL_STIFF) && A.defer_list[0]) {
        const int pos = atomicAdd(A.defer_count[0], 1);
        if (A.timeline) A.timeline[3 * pos] = odl_globaltimer();
        A.defer_list[0][pos] = (int)row;
      }
      if (consumer && A.timeline) A.timeline[3 * sys + 2] = odl_globaltimer();
#else
      if constexpr (SOLVER == 0) {
        if ((fin_status == ODL_MAXSTEPS || fin_status == ODL_STIFF) && A.handover && A.defer_list[0]) {
          // where this solve stands (frozen since it stopped, see odl_dopri5_attempt), published before the feed entry
          double* rec = A.handover + (size_t)row * A.handover_stride;
          rec[0] = st.t; rec[1] = (double)st.slot;
ODL_UNROLL
          for (int i = 0; i < ODL_N; ++i) rec[2 + i] = st.y[i];
          for (int j = 0; j < st.slot * ODL_NOUT; ++j) rec[2 + ODL_N + j] = my_stage[j];
          __threadfence();
        }
      }
      if (fin_status == ODL_MAXSTEPS && A.defer_list[0]) A.defer_list[0][atomicAdd(A.defer_count[0], 1)] = (int)row;
      if (fin_status == ODL_STIFF && A.defer_list[1]) A.defer_list[1][atomicAdd(A.defer_count[1], 1)] = (int)row;
#endif
    }
    if (fin) { active = false; done = false; want = lane_on; }
    // ---- (B) refill ----
    if (!consumer) {
      const long long got = odl_fetch(A.counter, want, lane);
      if (want) {
        want = false;
        long long r_ = -1;
        if (got >= 0 && got < n) r_ = A.index ? (long long)A.index[got] : got;
        // a feed entry below zero was finished by the pass that ran beside the bulk pass: take the next one
        if (got >= 0 && got < n && r_ < 0) want = lane_on;
        if (r_ >= 0) {
          sys = got;
          row = r_;
ODL_UNROLL
          for (int q = 0; q < ODL_P; ++q) p[q] = A.theta[row * ODL_P + q];
          odl_init_system(st, p, D, O, nullptr, false);
          ax.reset();
          if (!odl_take_over<SOLVER>(st, p, A, row, my_stage)) odl_emit_initial_slots(st, S, D, sink);
          active = true;
          done = (st.slot >= D.n_slot);
          if (done) { fin_status = st.status; fin_nsteps = st.nsteps; }
        }
      }
      if (!__any_sync(ODL_FULL, active)) {
        if (__any_sync(ODL_FULL, want)) continue;                // only skipped entries this time: fetch again
        break;
      }
    } else {
      const long long got = odl_fetch(A.feed_ticket, want, lane);
      if (want) { want = false; pending = true; ticket = got; }
      if (__any_sync(ODL_FULL, pending)) {
        // *feed_done is set by a one-thread kernel that follows the last bulk launch in its stream: once it reads 1,
        // the count read AFTER it is final
        int landed = 0, complete = 0;
        if (lane == 0) {
          complete = odl_ld_acquire(A.feed_done);
          landed = odl_ld_acquire(A.index_count);
        }
        landed = __shfl_sync(ODL_FULL, landed, 0);
        complete = __shfl_sync(ODL_FULL, complete, 0);
        if (pending) {
          if (ticket < (long long)landed) {
            int r;
            do { r = odl_ld_acquire(A.index + ticket); } while (r < 0);       // written right after the count moved
            // taken: a lane that starts a system finishes it (a warp leaves only with no lane at work), so the pass
            // that follows this one skips the entry
            const_cast<int*>(A.index)[ticket] = -2;
            sys = ticket; row = r;
#if ODL_TIMELINE
            if (A.timeline) A.timeline[3 * ticket + 1] = odl_globaltimer();
#endif
ODL_UNROLL
            for (int q = 0; q < ODL_P; ++q) p[q] = A.theta[row * ODL_P + q];
            odl_init_system(st, p, D, O, nullptr, false);
            ax.reset();
            if (!odl_take_over<SOLVER>(st, p, A, row, my_stage)) odl_emit_initial_slots(st, S, D, sink);
            active = true; pending = false;
            done = (st.slot >= D.n_slot);
            if (done) { fin_status = st.status; fin_nsteps = st.nsteps; }
          } else if (complete) {
            pending = false;                                                   // the feed ended before this ticket
          }
        }
      }
      if (!__any_sync(ODL_FULL, active || pending)) break;
      if (!__any_sync(ODL_FULL, active)) {                                      // nothing to integrate yet
        // watchdog: a producer that never reports completion (a bug, a killed launch) must not hang the device
        if (++idle_spins > (unsigned int)O.watchdog_spins) { if (lane == 0 && A.watchdog) atomicAdd(A.watchdog, 1); break; }
        __nanosleep(400);
        continue;
      }
      idle_spins = 0;
    }
    // ---- (C) ODL_INNER step attempts between visits of (A)/(B): the ballots, shuffles and the refill logic cost
    //      about a fifth of a step; a finished lane idles for at most ODL_INNER-1 attempts (systems take ~90) ----
#pragma unroll 1          // unrolling by 2 measured 7 % slower (instruction cache)
    for (int r = 0; r < ODL_INNER; ++r) {
      if constexpr (SOLVER == 0) {
        // DOPRI5: EVERY lane runs the step code, with or without a system.  Wrapping the inlined step in
        // `if (active && !done)` made the compiler reconcile ~40 registers on the path around it -- 7 % of the
        // kernel's instructions, issued for the lanes that were NOT stepping (profiles/r1e).  A lane without work
        // steps a defined dummy state and writes nothing (slot == n_slot); a finished lane keeps the status and
        // step count of its solve in fin_*.
        odl_attempt<SOLVER>(st, ax, p, S, D, O, sink, false);
        if (active && !done) {
          done = (st.slot >= D.n_slot) || (st.status != ODL_OK);
          if (done) { fin_status = st.status; fin_nsteps = st.nsteps; }
        }
      } else {
        if (active && !done) {
          odl_attempt<SOLVER>(st, ax, p, S, D, O, sink, false);
          done = (st.slot >= D.n_slot) || (st.status != ODL_OK);
          if (done) { fin_status = st.status; fin_nsteps = st.nsteps; }
        }
      }
      if (!__any_sync(ODL_FULL, active && !done)) break;
    }
  }
}
#if ODL_HAS(1)
extern "C" __global__ void __launch_bounds__(ODL_BLOCK, ODL_MINBLOCKS)
odl_sweep_kernel(const OdlData D, const OdlOpts O, const OdlSweepArgs A) { odl_sweep_body<0>(D, O, A); }
#endif
#if ODL_HAS(6)
extern "C" __global__ void __launch_bounds__(ODL_BLOCK, ODL_MINBLOCKS_ROS)
odl_sweep_ros23_kernel(const OdlData D, const OdlOpts O, const OdlSweepArgs A) { odl_sweep_body<1>(D, O, A); }
#endif
#if ODL_HAS(9)
extern "C" __global__ void __launch_bounds__(256, 1)
odl_sweep_radau5_kernel(const OdlData D, const OdlOpts O, const OdlSweepArgs A) { odl_sweep_body<3>(D, O, A); }
#endif
// (register caps for more resident warps spill: 168 registers -> 3.1 ms against 1.6 ms at 226, tools/variant_ab.py).
// CTAs of one warp when the pass runs after the bulk pass, of 8 warps when it runs beside it on SMs of its own.
#ifndef ODL_BDF_THREADS
#define ODL_BDF_THREADS 256
#endif
#if ODL_HAS(4)
extern "C" __global__ void __launch_bounds__(ODL_BDF_THREADS, 1)
odl_sweep_bdf_kernel(const OdlData D, const OdlOpts O, const OdlSweepArgs A) { odl_sweep_body<4>(D, O, A); }
#endif

// ------------------------------------------------------------------------------------------------
// Full trajectories on the output grid (ModelFramework.integrate, Framework.py:622-683)
// ------------------------------------------------------------------------------------------------
#if ODL_HAS(2)
extern "C" __global__ void __launch_bounds__(ODL_BLOCK, ODL_MINBLOCKS)
odl_traj_kernel(const OdlData D, const OdlOpts O, const OdlTrajArgs A) {
  OdlShared S;
  S.slot_t = odl_smem; S.lnO = S.w = S.lin = nullptr; S.src = nullptr; S.stage = nullptr;
  for (int i = threadIdx.x; i < D.n_slot; i += blockDim.x) S.slot_t[i] = D.slot_t[i];
  __syncthreads();
  const long long sys = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (sys >= A.n) return;
  // (fully unrolled whatever the state count: with `unroll 1` -- ODL_UNROLL for n > 8 -- NVRTC 12.9 treated p[0] as never
  // written and folded dy[0] to a NaN constant: every trajectory of a model with more than 8 states failed at once)
  double p[ODL_P];
  const double* theta_row = A.theta + sys * ODL_P;
#pragma unroll
  for (int q = 0; q < ODL_P; ++q) p[q] = theta_row[q];
  OdlStepper st;
  OdlTrajSink sink; sink.traj = A.traj + sys * (long long)D.n_slot * ODL_N;
  odl_init_system(st, p, D, O, A.y0 ? A.y0 + sys * ODL_N : nullptr, false);
  odl_emit_initial_slots(st, S, D, sink);
  while (st.slot < D.n_slot && st.status == ODL_OK) odl_dopri5_attempt(st, p, S, D, O, sink);
  if (st.status != ODL_OK) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int s = st.slot; s < D.n_slot; ++s)
      for (int i = 0; i < ODL_N; ++i) sink.traj[(long long)s * ODL_N + i] = nan;
  }
  A.status[sys] = st.status; A.nsteps[sys] = st.nsteps;
}
#endif  // unit 2

// ------------------------------------------------------------------------------------------------
// Metropolis-Hastings: Samplers.py:53-174, one chain per thread, device resident
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void odl_propose(double (&p)[ODL_P], const OdlMcmcArgs& A, int chain_local, int it) {
  // theta' = exp(log(theta_old) + N(0, step_sd))   for every walking parameter (Framework.py:119)
  const double* cur = A.theta_cur + (size_t)chain_local * ODL_P;
ODL_UNROLL
  for (int q = 0; q < ODL_P; ++q) p[q] = cur[q];
  const long long k = (long long)chain_local * A.n_iter_total + (it - 1);
  if (A.rng_mode == 2) {
ODL_UNROLL
    for (int q = 0; q < ODL_P; ++q) p[q] = A.forced[k * ODL_P + q];
    return;
  }
  const unsigned long long gchain = A.chain_ids ? (unsigned long long)A.chain_ids[chain_local] : (unsigned long long)(A.chain_offset + chain_local);
  for (int j = 0; j < A.n_walk; j += 2) {
    double z0, z1 = 0.0;
    if (A.rng_mode == 1) {
      z0 = A.z[k * A.n_walk + j];
      if (j + 1 < A.n_walk) z1 = A.z[k * A.n_walk + j + 1];
    } else {
      const OdlPhilox r = odl_philox((unsigned int)it, (unsigned int)(1 + (j >> 1)), (unsigned int)gchain,
                                     (unsigned int)(gchain >> 32), (unsigned int)A.seed, (unsigned int)(A.seed >> 32));
      const double u1 = 1.0 - odl_u53(r.x, r.y);                // (0,1]
      const double u2 = odl_u53(r.z, r.w);
      const double rad = sqrt(-2.0 * log(u1));
      double sn, cs;
      sincospi(2.0 * u2, &sn, &cs);
      z0 = A.step_sd * (rad * cs);
      z1 = A.step_sd * (rad * sn);
    }
#if ODL_SMALL
ODL_UNROLL
    for (int q = 0; q < ODL_P; ++q) {
      if (A.walk[j] == q) p[q] = exp(log(p[q]) + z0);
      if (j + 1 < A.n_walk && A.walk[j + 1] == q) p[q] = exp(log(p[q]) + z1);
    }
#else
    p[A.walk[j]] = exp(log(p[A.walk[j]]) + z0);
    if (j + 1 < A.n_walk) p[A.walk[j + 1]] = exp(log(p[A.walk[j + 1]]) + z1);
#endif
  }
}

// log prior density of one parameter value (scipy.stats parameterisation; table row = kind, a, b, c):
// 1 lognorm(s = a, loc = b, scale = c), 2 norm(loc = b, scale = c), 3 uniform(loc = b, scale = c), else flat
__device__ __forceinline__ double odl_log_prior_term(const double* T, double th) {
  const int kind = (int)T[0];
  const double a = T[1], b = T[2], c = T[3];
  const double ninf = __longlong_as_double(0xfff0000000000000LL);
  if (kind == 1) {
    const double x = th - b;
    if (!(x > 0.0)) return ninf;
    const double lx = log(x / c);
    return -log(x * a * 2.5066282746310002) - lx * lx / (2.0 * a * a);
  }
  if (kind == 2) {
    const double zz = (th - b) / c;
    return -log(c * 2.5066282746310002) - 0.5 * zz * zz;
  }
  if (kind == 3) return (th >= b && th <= b + c) ? -log(c) : ninf;
  return 0.0;
}
// log prior of a parameter vector, and the Hastings term sum ln(theta'/theta) of the multiplicative walk against `cur`
// (static parameters contribute ln 1 = 0)
template <class PV>
__device__ __forceinline__ void odl_log_prior(const OdlMcmcArgs& A, const PV& p, const double* cur, double& lp, double& hastings) {
  lp = 0.0; hastings = 0.0;
ODL_UNROLL
  for (int q = 0; q < ODL_P; ++q) {
    lp += odl_log_prior_term(A.prior + 4 * q, p[q]);
    if (cur) hastings += log(p[q]) - log(cur[q]);
  }
}

__device__ __forceinline__ double odl_mh_uniform(const OdlMcmcArgs& A, int chain_local, int it) {
  if (A.rng_mode != 0) return A.u[(long long)chain_local * A.n_iter_total + (it - 1)];
  const unsigned long long gchain = A.chain_ids ? (unsigned long long)A.chain_ids[chain_local] : (unsigned long long)(A.chain_offset + chain_local);
  const OdlPhilox r = odl_philox((unsigned int)it, 0u, (unsigned int)gchain, (unsigned int)(gchain >> 32),
                                 (unsigned int)A.seed, (unsigned int)(A.seed >> 32));
  return odl_u53(r.x, r.y);
}

// Prefetching Metropolis-Hastings.  A chain is sequential, and with few chains (BASELINE config 3: 4096) the GPU
// is latency-bound: one warp per SM.  A.spec = K lanes per chain (a power of two, chosen by the host from the
// chain count) evaluate the next K iterations AT ONCE along the all-rejected path: lane j proposes
// theta_cur * exp(z_{it+j}) -- what iteration it+j proposes if iterations it..it+j-1 reject, which is what
// happens ~74 % of the time at the demo's acceptance rate -- and the group then consumes iterations up to and
// including the first acceptance, discarding the rest.  Proposals and uniforms are keyed by (chain, iteration),
// so the chain is the SAME chain, decision by decision and bit by bit, for every K (K = 1 is the plain loop);
// expected iterations consumed per round (1 - (1-a)^K)/a: 2.7 for K = 4, 3.5 for K = 8 at a = 0.26.
// One kept row (theta.., chi, rsquared, aic, iteration, acceptance_ratio) to HBM with 16-byte stores: rows are
// contiguous 8 (P+5)-byte records, so a row starts on a 16-byte boundary or 8 bytes past one.
__device__ __forceinline__ void odl_store_row(double* row, const double (&v)[ODL_P + 5]) {
  constexpr int LEN = ODL_P + 5;
  if ((((unsigned long long)row) & 15ull) == 0ull) {
#pragma unroll
    for (int i = 0; i + 1 < LEN; i += 2) *reinterpret_cast<double2*>(row + i) = make_double2(v[i], v[i + 1]);
    if (LEN & 1) row[LEN - 1] = v[LEN - 1];
  } else {
    row[0] = v[0];
#pragma unroll
    for (int i = 1; i + 1 < LEN; i += 2) *reinterpret_cast<double2*>(row + i) = make_double2(v[i], v[i + 1]);
    if (!(LEN & 1)) row[LEN - 1] = v[LEN - 1];
  }
}

template <int SOLVER>
__device__ __forceinline__ void odl_mcmc_body(const OdlData& D, const OdlOpts& O, const OdlMcmcArgs& A) {
  const OdlShared S = odl_carve(odl_smem, D);
  odl_load_tables(S, D);
  const int lane = threadIdx.x & 31;
  double* my_stage = S.stage + (size_t)threadIdx.x * D.stage_stride;
  OdlStageSink sink; sink.stage = my_stage;
  const int K = (A.spec >= 1 && A.spec <= 32) ? A.spec : 1;
  const int sub = lane & (K - 1), gbase = lane & ~(K - 1);
  const unsigned gbits = (K == 32) ? 0xffffffffu : ((1u << K) - 1u);
  const long long gthread = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int chain = (int)(gthread / K);                         // local chain index
  const bool has_chain = gthread / K < (long long)A.n_chain;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);

  OdlStepper st;
  typename OdlAuxOf<SOLVER>::type ax;
  ax.reset();
  double p[ODL_P];
  st.status = ODL_OK; st.nsteps = 0; st.slot = 0;
  // Chain state (chi_cur, r2_cur, accepts, best_chi, best_iteration) lives in A.chain_state, not in registers: it is
  // touched once per round by a few lanes, and registers held across the integration loop are what the DOPRI5 stepper
  // is short of (128 per thread at 4 CTAs/SM).  `it` = next iteration of Samplers.py:104 to decide, the same on the K
  // lanes of a chain; apriori marks the solve of the starting point of a fresh chain (Samplers.py:88-90).
  int it = A.it_begin;
  bool apriori = A.it_begin == 1;                               // warp-uniform: the branches below hold collectives
  bool active = false, done = false, use_alt = false;
  bool dead = false;                                             // stop_failed: this chain met a solve that failed
  // SOLVER 2: the explicit attempt of a solve runs under its own step budget
  OdlOpts Oe = O;
  if (SOLVER == 2 && O.explicit_cap > 0) Oe.max_steps = min(O.explicit_cap, O.max_steps);
  double* cs = A.chain_state + (size_t)(has_chain ? chain : 0) * ODL_CHAIN_STATE;
  bool more = has_chain && (apriori || it < A.it_end);

  while (__any_sync(ODL_FULL, more)) {
    // ---- start a round: lane `sub` takes iteration it+sub ----
    const bool valid = more && (apriori ? (sub == 0) : (it + sub < A.it_end));
    active = false; done = false; use_alt = false;
    if (valid) {
      if (apriori) {
ODL_UNROLL
        for (int q = 0; q < ODL_P; ++q) p[q] = A.theta_cur[(size_t)chain * ODL_P + q];
      } else {
        odl_propose(p, A, chain, it + sub);
      }
      odl_init_system(st, p, D, O, nullptr, !apriori);          // proposals: '<state>0' parameters set y0 (Samplers.py:110-114)
      ax.reset();
      odl_emit_initial_slots(st, S, D, sink);
      active = true;
      done = (st.slot >= D.n_slot);
    }
    // ---- integrate: the lanes of a warp step together until the last one is done.  (Finalising lane by lane as
    //      lanes finish left half of all issued instructions with one lane active -- profiles/r1a_mcmc_*; proposals
    //      of neighbouring chains, and the K proposals of one chain, cost nearly the same number of steps.) ----
    for (;;) {
      if constexpr (SOLVER == 0) {
        // DOPRI5: every lane that has a solve this round runs the step code, finished or not (as in odl_sweep_body: the
        // path AROUND the inlined step reconciles ~40 registers, and the lanes that wait for the slowest of the warp paid
        // it on every attempt; now only a lane without a solve -- past the end of its chain -- takes it).  A finished
        // lane writes nothing (slot == n_slot), its step count stands still and status words only ever leave OK.
        if (active) {
          odl_attempt<SOLVER>(st, ax, p, S, D, O, sink, false);
          done = (st.slot >= D.n_slot) || (st.status != ODL_OK);
        }
      } else if (active && !done) {
        odl_attempt<SOLVER>(st, ax, p, S, D, (SOLVER == 2 && !use_alt) ? Oe : O, sink, use_alt);
        done = (st.slot >= D.n_slot) || (st.status != ODL_OK);
      }
      if (__ballot_sync(ODL_FULL, active && !done)) continue;
      if (SOLVER == 2) {
        // DOPRI5 gave up on this proposal (step budget spent, or Hairer's test called it stiff): redo this very solve
        // with the BDF stepper.  Which stepper finishes a solve depends on that solve alone.
        const bool restart = active && (st.status == ODL_STIFF || st.status == ODL_MAXSTEPS) && !use_alt;
        if (__ballot_sync(ODL_FULL, restart)) {
          if (restart) {
            if (A.step_count) atomicAdd((unsigned long long*)&A.step_count[chain], (unsigned long long)st.nsteps);
            use_alt = true;
            odl_init_system(st, p, D, O, nullptr, !apriori);
            ax.reset();
            odl_emit_initial_slots(st, S, D, sink);
            done = (st.slot >= D.n_slot);
          }
          continue;
        }
      }
      break;
    }
    // ---- score every lane's own solve ----
    double my_chi = nan, my_r2 = nan;
    if (valid) {
      double chi, ss; int nv;
      odl_score_self(S, D, my_stage, chi, ss, nv);
      if (st.status == ODL_OK) { my_chi = (nv > 0) ? chi : nan; my_r2 = fma(-ss, D.inv_sstot, 1.0); }   // nv == 0: np.ma.masked
    }
    if (apriori) {
      // the starting point's own chi: no decision, no iteration consumed
      if (valid) {
        cs[0] = my_chi; cs[1] = my_r2;
        if (A.prior) { double lp0, h0; odl_log_prior(A, p, nullptr, lp0, h0); cs[5] = lp0; }
        if (A.step_count) atomicAdd((unsigned long long*)&A.step_count[chain], (unsigned long long)st.nsteps);
        if (A.fail_count && st.status != ODL_OK) atomicAdd(&A.fail_count[chain], 1);
      }
      if (A.stop_failed && ((__ballot_sync(ODL_FULL, valid && st.status != ODL_OK) >> gbase) & gbits)) dead = true;
      apriori = false;
      __syncwarp();
    } else {
      const double chi_cur = cs[0], r2_cur = cs[1];
      const int accepts = (int)cs[2];
      // ---- decisions along the all-rejected path: acc = exp(chi - chinew) > u (Samplers.py:124-127; NaN rejects) ----
      bool acc = false;
      double lp_new = 0.0;
      double* cur = A.theta_cur + (size_t)chain * ODL_P;
      if (valid) {
        const double u = odl_mh_uniform(A, chain, it + sub);
        if (!A.prior) {
          acc = exp(chi_cur - my_chi) > u;
        } else {                                                 // posterior ratio: prior log-densities + Hastings term
          double hast;
          odl_log_prior(A, p, cur, lp_new, hast);
          acc = exp((chi_cur - my_chi) + (lp_new - cs[5]) + hast) > u;
        }
      }
      const unsigned gacc = (__ballot_sync(ODL_FULL, acc) >> gbase) & gbits;
      const int nvalid = __popc((__ballot_sync(ODL_FULL, valid) >> gbase) & gbits);     // valid lanes are a prefix
      const int jstar = gacc ? (__ffs(gacc) - 1) : -1;                                    // first acceptance
      const int adv = (jstar >= 0) ? jstar + 1 : nvalid;                                  // iterations consumed
      const bool consumed = valid && sub < adv;
      const bool is_acc = consumed && sub == jstar;
      if (A.stop_failed && ((__ballot_sync(ODL_FULL, consumed && st.status != ODL_OK) >> gbase) & gbits)) dead = true;
      if (consumed) {
        // counters cover CONSUMED solves only (speculative work that was discarded is not counted)
        if (A.step_count) atomicAdd((unsigned long long*)&A.step_count[chain], (unsigned long long)st.nsteps);
        if (A.fail_count && st.status != ODL_OK) atomicAdd(&A.fail_count[chain], 1);
        const int iter = it + sub;
        const long long k = (long long)chain * A.n_iter_total + (iter - 1);
        if (A.trace_chinew) A.trace_chinew[k] = my_chi;
        if (A.trace_accept) A.trace_accept[k] = is_acc ? 1 : 0;
      }
      // ---- kept rows (Samplers.py:147-153) -> HBM.  Row (chain, rowi) lives at samples[chain * chain_pitch + rowi *
      //      row_pitch]: chain-major (the reference frame's order) or iteration-major.  Iteration-major with one lane per
      //      chain -- the throughput regime, where the stream matters -- puts the 32 rows a warp keeps in one round next
      //      to each other: they go through the warp's (free) staging rows in shared memory and leave as 32-lane
      //      contiguous stores, every instruction writing whole sectors.  Otherwise one 16-byte-vector row per lane.
      {
        const int iter = it + sub;
        const int rowi = iter - A.burnin - 1;
        const bool keep = consumed && rowi >= 0 && A.samples && rowi < A.n_keep;
        constexpr int LEN = ODL_P + 5;
        const bool together = K == 1 && A.smp_chain_pitch == (long long)LEN && A.row_stride == LEN && D.stage_stride >= LEN &&
                              __all_sync(ODL_FULL, keep);
        const double c = is_acc ? my_chi : chi_cur;
        if (together) {
ODL_UNROLL
          for (int q = 0; q < ODL_P; ++q) my_stage[q] = is_acc ? p[q] : cur[q];
          my_stage[ODL_P + 0] = c;
          my_stage[ODL_P + 1] = is_acc ? my_r2 : r2_cur;
          my_stage[ODL_P + 2] = 2.0 * c + 2.0 * (double)A.pnum;   // stats.py:46
          my_stage[ODL_P + 3] = (double)iter;
          my_stage[ODL_P + 4] = (double)(accepts + (is_acc ? 1 : 0)) / (double)iter;   // Samplers.py:153
          __syncwarp();
          const double* wst = S.stage + (size_t)(threadIdx.x & ~31) * D.stage_stride;
          double* base = A.samples + (long long)rowi * A.smp_row_pitch + (long long)(chain - lane) * LEN;
          if ((LEN & 1) == 0 && (((unsigned long long)base) & 15ull) == 0ull) {
#pragma unroll
            for (int j = 0; j < LEN / 2; ++j) {                   // 16 bytes per lane, 512 contiguous bytes per instruction
              const int f = 2 * (j * 32 + lane), l = f / LEN, e = f - l * LEN;
              const double* src = wst + l * D.stage_stride + e;
              reinterpret_cast<double2*>(base)[j * 32 + lane] = make_double2(src[0], src[1]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < LEN; ++j) {
              const int f = j * 32 + lane, l = f / LEN, e = f - l * LEN;
              base[f] = wst[l * D.stage_stride + e];
            }
          }
          __syncwarp();                                            // the staging rows are free again for the next solve
        } else if (keep) {
          double v[ODL_P + 5];
ODL_UNROLL
          for (int q = 0; q < ODL_P; ++q) v[q] = is_acc ? p[q] : cur[q];
          v[ODL_P + 0] = c;
          v[ODL_P + 1] = is_acc ? my_r2 : r2_cur;
          v[ODL_P + 2] = 2.0 * c + 2.0 * (double)A.pnum;
          v[ODL_P + 3] = (double)iter;
          v[ODL_P + 4] = (double)(accepts + (is_acc ? 1 : 0)) / (double)iter;
          odl_store_row(A.samples + (long long)chain * A.smp_chain_pitch + (long long)rowi * A.smp_row_pitch, v);
        }
      }
      // the accepted proposal, on every lane of the group (for the summaries; lane jstar holds it in p) -- fetched
      // only now: nothing above needs it, and the row stores are short of registers
      double pacc[ODL_P];
      const int src = gbase + (jstar >= 0 ? jstar : 0);
ODL_UNROLL
      for (int q = 0; q < ODL_P; ++q) pacc[q] = __shfl_sync(ODL_FULL, p[q], src);
      const double chi_acc = __shfl_sync(ODL_FULL, my_chi, src), r2_acc = __shfl_sync(ODL_FULL, my_r2, src);
      // Welford over ln(theta) of the kept rows for R-hat: sequential in the iteration (the same recurrence, in the
      // same order, whatever K is), by the group's first lane
      if (A.summaries && has_chain && sub == 0 && it + adv - 1 > A.burnin) {
        // one parameter at a time (all of them at once kept 4 P doubles live and spilled at 128 registers)
        double* sm = A.summaries + (size_t)chain * (1 + 2 * ODL_P);
        const double cnt0 = sm[0];
        double cnt = cnt0;
ODL_UNROLL
        for (int q = 0; q < ODL_P; ++q) {
          double mean = sm[1 + q], m2 = sm[1 + ODL_P + q];
          const double xr = (jstar == 0) ? 0.0 : log(cur[q]);    // jstar == 0: the only consumed row is the accepted one
          const double xa = (jstar >= 0) ? log(pacc[q]) : 0.0;
          cnt = cnt0;
          for (int i = 0; i < adv; ++i) {
            if (it + i <= A.burnin) continue;
            cnt += 1.0;
            const double x = (i == jstar) ? xa : xr;
            const double dlt = x - mean;
            mean += dlt / cnt;
            m2 += dlt * (x - mean);
          }
          sm[1 + q] = mean; sm[1 + ODL_P + q] = m2;
        }
        sm[0] = cnt;
      }
      // Best kept row (set_best_params' idxmin, Framework.py:725-731): the first kept row carries the current point;
      // after that only an accepted proposal can be a new minimum (rejected rows repeat a chi already seen)
      if (has_chain && sub == 0) {
        const int first_kept = A.burnin + 1;
        double best_chi = cs[3], best_it = cs[4];
        if (it <= first_kept && first_kept < it + adv && first_kept - it != jstar) {
          best_chi = chi_cur; best_it = (double)first_kept;
          if (A.best_theta) {
ODL_UNROLL
            for (int q = 0; q < ODL_P; ++q) A.best_theta[(size_t)chain * ODL_P + q] = cur[q];
          }
        }
        if (jstar >= 0 && it + jstar > A.burnin && (chi_acc < best_chi || (best_chi != best_chi && chi_acc == chi_acc) || best_it == 0.0)) {
          best_chi = chi_acc; best_it = (double)(it + jstar);
          if (A.best_theta) {
ODL_UNROLL
            for (int q = 0; q < ODL_P; ++q) A.best_theta[(size_t)chain * ODL_P + q] = pacc[q];
          }
        }
        cs[3] = best_chi; cs[4] = best_it;
      }
      __syncwarp();                                              // everything above read cur[] / cs[] before they change
      if (is_acc) {
        cs[0] = my_chi; cs[1] = my_r2; cs[2] = (double)(accepts + 1);
        if (A.prior) cs[5] = lp_new;
ODL_UNROLL
        for (int q = 0; q < ODL_P; ++q) cur[q] = p[q];
      }
      it += adv;
      __syncwarp();                                              // the next round's proposals read cur[]
    }
    more = has_chain && !dead && it < A.it_end;
  }
}
#if ODL_HAS(3)
extern "C" __global__ void __launch_bounds__(ODL_BLOCK, ODL_MINBLOCKS_MCMC)
odl_mcmc_kernel(const OdlData D, const OdlOpts O, const OdlMcmcArgs A) { odl_mcmc_body<0>(D, O, A); }
#endif
#if ODL_HAS(7)
extern "C" __global__ void __launch_bounds__(ODL_BLOCK, ODL_MINBLOCKS_ROS)
odl_mcmc_ros23_kernel(const OdlData D, const OdlOpts O, const OdlMcmcArgs A) { odl_mcmc_body<1>(D, O, A); }
#endif
#if ODL_HAS(8)
extern "C" __global__ void __launch_bounds__(32, 1)
odl_mcmc_auto_kernel(const OdlData D, const OdlOpts O, const OdlMcmcArgs A) { odl_mcmc_body<2>(D, O, A); }
#endif
#if ODL_HAS(10)
extern "C" __global__ void __launch_bounds__(32, 1)
odl_mcmc_radau5_kernel(const OdlData D, const OdlOpts O, const OdlMcmcArgs A) { odl_mcmc_body<3>(D, O, A); }
#endif
#if ODL_HAS(5)
extern "C" __global__ void __launch_bounds__(32, 1)
odl_mcmc_bdf_kernel(const OdlData D, const OdlOpts O, const OdlMcmcArgs A) { odl_mcmc_body<4>(D, O, A); }
#endif
// ================================================================================================
// Cooperative kernels for larger systems (n > 8): ODL_G lanes per system.
//
// Thread-per-system keeps y, the seven stage derivatives and the parameters in registers; beyond n = 8 that is
// no longer possible and the arrays fall into local memory (35 states: 6 KB per thread, 200 KB per warp -- one warp
// fills an SM's L1 and every access of the second warp goes to L2).  Here a system is spread over ODL_G lanes:
// lane `sub` owns components sub, sub+G, sub+2G, ... (ODL_C per lane) of the state and of every stage -- static
// register indices again -- and the stage combinations, the error norm and the dense output are sliced the same
// way.  The right-hand side couples everything, so the slice owners publish the stage state to a per-system row in
// shared memory and EVERY lane evaluates the full traced RHS from it, keeping only its own components (the
// assignments to other components fold away after inlining, the arithmetic feeding them mostly does not): the lanes
// of a group run one instruction stream, no divergence, no per-lane code.  Parameters sit in shared memory too.
// Groups are independent: every collective below uses the group's own lane mask.
// ================================================================================================
#if !ODL_SMALL
#ifndef ODL_G
#define ODL_G (ODL_N <= 16 ? 4 : (ODL_N <= 64 ? 8 : (ODL_N <= 128 ? 16 : 32)))
#endif
// Which components a lane owns.  With a slice plan from the tracer (ODL_COOP_SLICED: odl_rhs_slice, ODL_SLICE_PERM)
// components are laid out class by class -- outputs of the same SHAPE next to each other -- so that the lanes of a
// round evaluate one code on different leaves; without one, lane `sub` owns sub, sub+G, ...
#ifdef ODL_COOP_SLICED
static_assert(ODL_SLICE_G == ODL_G, "the slice plan was made for another number of lanes per system");
#define ODL_C ODL_CS
#else
#define ODL_C ((ODL_N + ODL_G - 1) / ODL_G)
#endif
#ifndef ODL_COOP_BLOCK
#define ODL_COOP_BLOCK 128
#endif
#ifndef ODL_COOP_MINBLOCKS
#define ODL_COOP_MINBLOCKS 2
#endif

struct OdlSmemView {             // y[i] / p[i] of the traced code -> shared memory
  const double* b;
  __device__ __forceinline__ double operator[](int i) const { return b[i]; }
};
struct OdlSliceOut {             // dy[k] = v of the traced code: kept when component k is this lane's
  double (&mine)[ODL_C];
  int sub;
  struct Ref {
    OdlSliceOut& o; int k;
    // a SELECT, not a branch.  Written as `if (mine) slot = v` the compiler branched around every output and sank the
    // arithmetic that feeds only that output into the branch: executed once per lane-of-the-group (G times) by 32/G lanes
    // each instead of once by all of them -- 19 % of the kernel's instructions at 7 lanes, 30 % of its stall samples, half
    // of all branches divergent (profiles/r2m_mcmc_coop_network_ncu.txt)
    __device__ __forceinline__ void operator=(double v) {
      const bool own = (k % ODL_G) == o.sub;
      const int hi = own ? __double2hiint(v) : __double2hiint(o.mine[k / ODL_G]);
      const int lo = own ? __double2loint(v) : __double2loint(o.mine[k / ODL_G]);
      o.mine[k / ODL_G] = __hiloint2double(hi, lo);
    }
  };
  __device__ __forceinline__ Ref operator[](int k) { return Ref{*this, k}; }
};
struct OdlGroup {                // this lane's place in its group and the group's scratch rows
  int sub;                       // 0 .. ODL_G-1
  unsigned mask;                 // lanes of the group
  double* ysm;                   // [ODL_N]  stage state published for the RHS / the observation sums
  double* psm;                   // [ODL_P]  parameters of the group's system
  double* stage;                 // [stage_stride] predictions at the observation slots
  int comp[ODL_C];               // component this lane owns in round c (-1: none, padding)
};
__device__ __forceinline__ double odl_group_sum(double v, unsigned mask) {
#pragma unroll
  for (int m = ODL_G >> 1; m > 0; m >>= 1) {
    const int lo = __shfl_xor_sync(mask, __double2loint(v), m), hi = __shfl_xor_sync(mask, __double2hiint(v), m);
    v += __hiloint2double(hi, lo);
  }
  return v;
}
// f = rhs(t, state published from the lanes' slices `yl`), slice of this lane
__device__ __forceinline__ void odl_coop_rhs(const double (&yl)[ODL_C], double t, const OdlGroup& G, double (&f)[ODL_C]) {
#pragma unroll
  for (int c = 0; c < ODL_C; ++c) { const int i = G.comp[c]; if (i >= 0) G.ysm[i] = yl[c]; f[c] = 0.0; }
  __syncwarp(G.mask);
  const OdlSmemView yv{G.ysm}, pv{G.psm};
#ifdef ODL_COOP_SLICED
  // this lane's outputs only: one class code per round, leaves through the index table (tracer.slice_plan).  The
  // whole traced RHS on every lane (below) is G-fold redundant: 8 x 260 flops per evaluation of the 35-state network
  odl_rhs_slice(yv, t, pv, G.sub, f);
#else
  OdlSliceOut out{f, G.sub};
  odl_rhs(yv, t, pv, out);
#endif
  __syncwarp(G.mask);            // every lane has read the row before it is published again
}

struct OdlCoopStepper {
  double y[ODL_C], k1[ODL_C];
  double t, h, tend;
  float lgfac;
  int nsteps, slot, status;
  bool last_rejected;
};

// observation columns of the state slices `yi` at `slot` -> the group's staging row (lane 0 writes)
__device__ __forceinline__ void odl_coop_emit(const double (&yi)[ODL_C], int slot, const OdlGroup& G) {
#pragma unroll
  for (int c = 0; c < ODL_C; ++c) { const int i = G.comp[c]; if (i >= 0) G.ysm[i] = yi[c]; }
  __syncwarp(G.mask);
  if (G.sub == 0) {
    const OdlSmemView yv{G.ysm};
    double out[ODL_NOUT];
    odl_observe(yv, out);
#pragma unroll
    for (int c = 0; c < ODL_NOUT; ++c) G.stage[slot * ODL_NOUT + c] = out[c];
  }
  __syncwarp(G.mask);
}

__device__ __forceinline__ void odl_coop_init(OdlCoopStepper& st, const OdlGroup& G, const OdlData& D, const OdlOpts& O,
                                              bool y0_from_params) {
#pragma unroll
  for (int c = 0; c < ODL_C; ++c) {
    const int i = G.comp[c];
    double v = 0.0;
    if (i >= 0) {
      v = D.y0[i];
#if ODL_Y0P
      const int src = D.y0_from_param[i];
      if (y0_from_params && src >= 0) v = G.psm[src];              // proposals only (see odl_init_system)
#endif
    }
    st.y[c] = v;
  }
  st.t = D.t0;
  st.tend = D.slot_t[D.n_slot - 1];
  st.nsteps = 0; st.slot = 0; st.status = ODL_OK;
  st.lgfac = ODL_LG_FACMIN; st.last_rejected = false;
  odl_coop_rhs(st.y, st.t, G, st.k1);
  const double span = st.tend - st.t;
  const double hmax = (O.hmax > 0.0) ? O.hmax : span;
  double h = O.h0;
  if (!(h > 0.0)) {
    // Hairer's hinit, norms summed over the group
    double dnf = 0.0, dny = 0.0, rsk[ODL_C];
#pragma unroll
    for (int c = 0; c < ODL_C; ++c) {
      rsk[c] = odl_rcp_approx(O.atol + O.rtol * fabs(st.y[c]));
      const double a = st.k1[c] * rsk[c], b = st.y[c] * rsk[c];
      dnf += a * a; dny += b * b;
    }
    const float fnf = (float)odl_group_sum(dnf, G.mask), fny = (float)odl_group_sum(dny, G.mask);
    const float hf = (fnf <= 1e-10f || fny <= 1e-10f) ? 1e-6f : 0.01f * sqrtf(fny * __frcp_rn(fnf));
    h = fmin((double)hf, hmax);
    double y1[ODL_C], f1[ODL_C];
#pragma unroll
    for (int c = 0; c < ODL_C; ++c) y1[c] = st.y[c] + h * st.k1[c];
    odl_coop_rhs(y1, st.t + h, G, f1);
    double d2 = 0.0;
#pragma unroll
    for (int c = 0; c < ODL_C; ++c) { const double a = (f1[c] - st.k1[c]) * rsk[c]; d2 += a * a; }
    const float der2 = sqrtf((float)odl_group_sum(d2, G.mask)) * __frcp_rn((float)h);
    const float der12 = fmaxf(der2, sqrtf(fnf));
    const float h1 = (der12 <= 1e-15f) ? fmaxf(1e-6f, (float)h * 1e-3f) : __powf(0.01f * __frcp_rn(der12), 0.2f);
    h = fmin(fmin(100.0 * h, (double)h1), hmax);
  }
  if (!(h > 0.0) || !odl_finite(h)) h = 1e-6 * (span > 0.0 ? span : 1.0);
  st.h = h;
  // slots at (or before) the start time take the initial state
  while (st.slot < D.n_slot && D.slot_t[st.slot] <= st.t) { odl_coop_emit(st.y, st.slot, G); ++st.slot; }
}

// One DOPRI5 step attempt of a group; the same arithmetic as odl_dopri5_attempt, sliced.  Control flow is uniform
// inside the group (every decision comes from group-reduced numbers).
__device__ __forceinline__ void odl_coop_attempt(OdlCoopStepper& st, const OdlGroup& G, const OdlShared& S, const OdlData& D,
                                                 const OdlOpts& O) {
  const double t = st.t;
  double h = st.h;
  bool last = false;
  if ((t + 1.01 * h - st.tend) > 0.0) { h = st.tend - t; last = true; }
  ++st.nsteps;
  double k2[ODL_C], k3[ODL_C], k4[ODL_C], k5[ODL_C], k6[ODL_C], k7[ODL_C], yt[ODL_C], yn[ODL_C];
#pragma unroll
  for (int i = 0; i < ODL_C; ++i) yt[i] = st.y[i] + h * (ODL_T(0) * st.k1[i]);
  odl_coop_rhs(yt, t + ODL_T(0) * h, G, k2);
#pragma unroll
  for (int i = 0; i < ODL_C; ++i) yt[i] = st.y[i] + h * (ODL_T(1) * st.k1[i] + ODL_T(2) * k2[i]);
  odl_coop_rhs(yt, t + ODL_T(32) * h, G, k3);
#pragma unroll
  for (int i = 0; i < ODL_C; ++i) yt[i] = st.y[i] + h * (ODL_T(3) * st.k1[i] + ODL_T(4) * k2[i] + ODL_T(5) * k3[i]);
  odl_coop_rhs(yt, t + ODL_T(33) * h, G, k4);
#pragma unroll
  for (int i = 0; i < ODL_C; ++i)
    yt[i] = st.y[i] + h * (ODL_T(6) * st.k1[i] + ODL_T(7) * k2[i] + ODL_T(8) * k3[i] + ODL_T(9) * k4[i]);
  odl_coop_rhs(yt, t + ODL_T(34) * h, G, k5);
#pragma unroll
  for (int i = 0; i < ODL_C; ++i)
    yt[i] = st.y[i] + h * (ODL_T(10) * st.k1[i] + ODL_T(11) * k2[i] + ODL_T(12) * k3[i] + ODL_T(13) * k4[i] +
                           ODL_T(14) * k5[i]);
  const double tph = t + h;
  odl_coop_rhs(yt, tph, G, k6);
#pragma unroll
  for (int i = 0; i < ODL_C; ++i)
    yn[i] = st.y[i] + h * (ODL_T(15) * st.k1[i] + ODL_T(16) * k3[i] + ODL_T(17) * k4[i] + ODL_T(18) * k5[i] +
                           ODL_T(19) * k6[i]);
  odl_coop_rhs(yn, tph, G, k7);
  double errsq = 0.0, ysum = 0.0;
  const double rtol_half = 0.5 * O.rtol;
#pragma unroll
  for (int i = 0; i < ODL_C; ++i) {
    const double e = h * (ODL_T(20) * st.k1[i] + ODL_T(21) * k3[i] + ODL_T(22) * k4[i] + ODL_T(23) * k5[i] +
                          ODL_T(24) * k6[i] + ODL_T(25) * k7[i]);
    const double sk = O.atol + rtol_half * (fabs(st.y[i]) + fabs(yn[i]));     // padding components: e = 0
    const double r = e * odl_rcp_approx(sk);
    errsq += r * r;
    ysum += fabs(yn[i]);
  }
  errsq = odl_group_sum(errsq, G.mask);
  ysum = odl_group_sum(ysum, G.mask);
  const bool finite_all = odl_finite(ysum);
  const float err = odl_sqrt_approx((float)errsq * (1.0f / ODL_N));
  const float lg_err = __log2f(err);
  if (err <= 1.0f && finite_all) {
    const float inv = ODL_PI_SAFETY * odl_ex2(ODL_PI_BETA * st.lgfac - (0.2f - ODL_PI_BETA * 0.75f) * lg_err);
    float fac = fminf(10.0f, fmaxf(0.2f, inv));
    if (!(err > 0.f)) fac = 10.0f;
    if (st.last_rejected) fac = fminf(fac, 1.0f);
    double hnew = h * (double)fac;
    st.lgfac = fmaxf(lg_err, ODL_LG_FACMIN);
    const double tnew = last ? st.tend : tph;
    if (st.slot < D.n_slot && S.slot_t[st.slot] <= tnew) {
      double rc2[ODL_C], rc3[ODL_C], rc4[ODL_C], rc5[ODL_C];
#pragma unroll
      for (int i = 0; i < ODL_C; ++i) {
        rc2[i] = yn[i] - st.y[i];
        rc3[i] = h * st.k1[i] - rc2[i];
        rc4[i] = rc2[i] - h * k7[i] - rc3[i];
        rc5[i] = h * (ODL_T(26) * st.k1[i] + ODL_T(27) * k3[i] + ODL_T(28) * k4[i] + ODL_T(29) * k5[i] +
                      ODL_T(30) * k6[i] + ODL_T(31) * k7[i]);
      }
      const double rh = odl_rcp(h);
      do {
        const double th = (S.slot_t[st.slot] - t) * rh, th1 = 1.0 - th;
        double yi[ODL_C];
#pragma unroll
        for (int i = 0; i < ODL_C; ++i) yi[i] = st.y[i] + th * (rc2[i] + th1 * (rc3[i] + th * (rc4[i] + th1 * rc5[i])));
        odl_coop_emit(yi, st.slot, G);
        ++st.slot;
      } while (st.slot < D.n_slot && S.slot_t[st.slot] <= tnew);
    }
#pragma unroll
    for (int i = 0; i < ODL_C; ++i) { st.y[i] = yn[i]; st.k1[i] = k7[i]; }
    st.t = tnew;
    st.last_rejected = false;
    st.h = hnew;
  } else {
    double hnew;
    if (err == err && finite_all && err < 3.0e38f)
      hnew = h * (double)fmaxf(0.2f, ODL_PI_SAFETY * odl_ex2(-(0.2f - ODL_PI_BETA * 0.75f) * lg_err));
    else hnew = 0.2 * h;
    st.last_rejected = true;
    st.h = hnew;
    if (!(fabs(hnew) > 4.0 * 2.220446049250313e-16 * fmax(fabs(t), fabs(st.tend)))) st.status = ODL_HUNDERFLOW;
  }
  if (st.nsteps >= O.max_steps && st.slot < D.n_slot && st.status == ODL_OK) st.status = ODL_MAXSTEPS;
}

// chi / R^2 of the group's staging row: the group's lanes share the observation rows (stats.py:41, :49-56)
__device__ __forceinline__ void odl_coop_score(const OdlShared& S, const OdlData& D, const OdlGroup& G, double* pred_out,
                                               double& chi, double& ssres, int& nvalid) {
  double c = 0.0, s = 0.0, k = 0.0;
  for (int o = G.sub; o < D.n_obs; o += ODL_G) {
    const double pred = G.stage[S.src[o]];
    if (pred_out) pred_out[o] = pred;
    const double d = __dadd_rn(S.lnO[o], -odl_log(S, pred));
    const double dd = __dmul_rn(d, d);
    const double term = __dmul_rn(dd, S.w[o]);
    const bool ok = odl_finite(dd) && odl_finite(term);
    if (ok) { c += term; k += 1.0; }
    const double r = __dadd_rn(pred, -S.lin[o]);
    const double rr = __dmul_rn(r, r);
    if (rr == rr) s += rr;
  }
  chi = odl_group_sum(c, G.mask);
  ssres = odl_group_sum(s, G.mask);
  nvalid = (int)odl_group_sum(k, G.mask);
}

__device__ __forceinline__ OdlGroup odl_coop_group(const OdlShared& S, const OdlData& D) {
  // shared memory after the tables: per group  ysm[ODL_N] | psm[ODL_P] | stage[stage_stride]
  const int lane = threadIdx.x & 31;
  OdlGroup G;
  G.sub = lane & (ODL_G - 1);
  G.mask = (ODL_G == 32) ? 0xffffffffu : (((1u << ODL_G) - 1u) << (lane & ~(ODL_G - 1)));
  const int group = threadIdx.x / ODL_G;
  double* base = S.stage + (size_t)group * ODL_COOP_ROW(ODL_N + ODL_P + D.stage_stride, ODL_G);
  G.ysm = base; G.psm = base + ODL_N; G.stage = base + ODL_N + ODL_P;
#pragma unroll
  for (int c = 0; c < ODL_C; ++c) {
#ifdef ODL_COOP_SLICED
    G.comp[c] = ODL_SLICE_PERM[c * ODL_G + G.sub];
#else
    G.comp[c] = (G.sub + c * ODL_G < ODL_N) ? G.sub + c * ODL_G : -1;
#endif
  }
  return G;
}

#if ODL_HAS(11)
// ---- forward sweep, one group per system at a time, work counter refill ----
extern "C" __global__ void __launch_bounds__(ODL_COOP_BLOCK, ODL_COOP_MINBLOCKS)
odl_sweep_coop_kernel(const OdlData D, const OdlOpts O, const OdlSweepArgs A) {
  const OdlShared S = odl_carve(odl_smem, D);
  odl_load_tables(S, D);
  const OdlGroup G = odl_coop_group(S, D);
  const long long n = A.index_count ? (long long)(*A.index_count) : A.n;
  OdlCoopStepper st;
  for (;;) {
    long long sys = 0;
    if (G.sub == 0) sys = (long long)atomicAdd(A.counter, 1ull);
    sys = ((long long)__shfl_sync(G.mask, (int)(sys >> 32), 0, ODL_G) << 32) | (unsigned int)__shfl_sync(G.mask, (int)(sys & 0xffffffffLL), 0, ODL_G);
    if (sys >= n) break;
    const long long row = A.index ? (long long)A.index[sys] : sys;
    for (int q = G.sub; q < ODL_P; q += ODL_G) G.psm[q] = A.theta[row * ODL_P + q];
    __syncwarp(G.mask);
    odl_coop_init(st, G, D, O, false);
    while (st.slot < D.n_slot && st.status == ODL_OK) odl_coop_attempt(st, G, S, D, O);
    double chi, ss; int nv;
    double* pred_out = (A.pred && st.status == ODL_OK) ? A.pred + row * D.n_obs : nullptr;
    odl_coop_score(S, D, G, pred_out, chi, ss, nv);
    if (G.sub == 0) {
      int status = st.status;
      double r2 = fma(-ss, D.inv_sstot, 1.0);
      if (status != ODL_OK) { chi = __longlong_as_double(0x7ff8000000000000LL); r2 = chi; }
      else if (nv == 0) { chi = __longlong_as_double(0x7ff8000000000000LL); status |= ODL_ALLMASKED; }
      A.chi[row] = chi;
      if (A.r2) A.r2[row] = r2;
      if (A.status) A.status[row] = status;
      if (A.nsteps) A.nsteps[row] = st.nsteps;
      if (st.status == ODL_MAXSTEPS && A.defer_list[0]) A.defer_list[0][atomicAdd(A.defer_count[0], 1)] = (int)row;
    }
    __syncwarp(G.mask);
  }
}
#endif  // unit 11

// ---- Metropolis-Hastings, one group per chain (Samplers.py:53-174); same outputs as odl_mcmc_body ----
__device__ __forceinline__ void odl_coop_propose(const OdlGroup& G, const OdlMcmcArgs& A, int chain, int it) {
  const double* cur = A.theta_cur + (size_t)chain * ODL_P;
  const long long k = (long long)chain * A.n_iter_total + (it - 1);
  if (A.rng_mode == 2) {
    for (int q = G.sub; q < ODL_P; q += ODL_G) G.psm[q] = A.forced[k * ODL_P + q];
    __syncwarp(G.mask);
    return;
  }
  for (int q = G.sub; q < ODL_P; q += ODL_G) G.psm[q] = cur[q];
  __syncwarp(G.mask);
  const unsigned long long gchain = A.chain_ids ? (unsigned long long)A.chain_ids[chain] : (unsigned long long)(A.chain_offset + chain);
  for (int j = G.sub; j < A.n_walk; j += ODL_G) {
    double z;
    if (A.rng_mode == 1) {
      z = A.z[k * A.n_walk + j];
    } else {
      const OdlPhilox r = odl_philox((unsigned int)it, (unsigned int)(1 + (j >> 1)), (unsigned int)gchain,
                                     (unsigned int)(gchain >> 32), (unsigned int)A.seed, (unsigned int)(A.seed >> 32));
      const double u1 = 1.0 - odl_u53(r.x, r.y);
      const double u2 = odl_u53(r.z, r.w);
      const double rad = sqrt(-2.0 * log(u1));
      double sn, cs;
      sincospi(2.0 * u2, &sn, &cs);
      z = A.step_sd * (rad * ((j & 1) ? sn : cs));
    }
    const int q = A.walk[j];
    G.psm[q] = exp(log(cur[q]) + z);
  }
  __syncwarp(G.mask);
}

#if ODL_HAS(12)
// One chain per K groups (A.spec = K, a power of two with K * ODL_G <= 32): prefetching MH as in odl_mcmc_body, a
// group taking the place of a lane -- group kk evaluates iteration it+kk along the all-rejected path, the K groups
// consume iterations up to and including the first acceptance.  The chain does not depend on K.
extern "C" __global__ void __launch_bounds__(ODL_COOP_BLOCK, ODL_COOP_MINBLOCKS)
odl_mcmc_coop_kernel(const OdlData D, const OdlOpts O, const OdlMcmcArgs A) {
  const OdlShared S = odl_carve(odl_smem, D);
  odl_load_tables(S, D);
  const OdlGroup G = odl_coop_group(S, D);
  const int lane = threadIdx.x & 31;
  int K = (A.spec >= 1) ? A.spec : 1;
  if (K * ODL_G > 32) K = 32 / ODL_G;
  const int span = K * ODL_G;                                      // lanes of one chain
  const int kk = (lane / ODL_G) & (K - 1);                         // this group's place among the chain's groups
  const int sbase = lane & ~(span - 1);
  const unsigned smask = (span == 32) ? 0xffffffffu : (((1u << span) - 1u) << sbase);
  const long long gthread = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int chain = (int)(gthread / span);
  if (gthread / span >= (long long)A.n_chain) return;             // the lanes of a chain leave together
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  const bool lead = G.sub == 0;
  const size_t group_doubles = (size_t)ODL_COOP_ROW(ODL_N + ODL_P + D.stage_stride, ODL_G);
  double* cs = A.chain_state + (size_t)chain * ODL_CHAIN_STATE;
  double* cur = A.theta_cur + (size_t)chain * ODL_P;
  OdlCoopStepper st;
  st.status = ODL_OK; st.nsteps = 0;
  bool apriori = A.it_begin == 1;
  int it = A.it_begin;
  while (apriori || it < A.it_end) {
    const bool valid = apriori ? (kk == 0) : (it + kk < A.it_end);
    double my_chi = nan, my_r2 = nan;
    if (valid) {
      if (apriori) {
        for (int q = G.sub; q < ODL_P; q += ODL_G) G.psm[q] = cur[q];
        __syncwarp(G.mask);
      } else {
        odl_coop_propose(G, A, chain, it + kk);
      }
      odl_coop_init(st, G, D, O, !apriori);
      while (st.slot < D.n_slot && st.status == ODL_OK) odl_coop_attempt(st, G, S, D, O);
      double chi, ss; int nv;
      odl_coop_score(S, D, G, nullptr, chi, ss, nv);
      if (st.status == ODL_OK) { my_chi = (nv > 0) ? chi : nan; my_r2 = fma(-ss, D.inv_sstot, 1.0); }
    }
    if (apriori) {
      if (valid && lead) {
        cs[0] = my_chi; cs[1] = my_r2;
        if (A.prior) { double lp0, h0; const OdlSmemView pv{G.psm}; odl_log_prior(A, pv, nullptr, lp0, h0); cs[5] = lp0; }
        if (A.step_count) A.step_count[chain] += st.nsteps;
        if (A.fail_count && st.status != ODL_OK) A.fail_count[chain] += 1;
      }
      apriori = false;
      if (A.stop_failed && __ballot_sync(smask, valid && lead && st.status != ODL_OK)) break;   // see OdlMcmcArgs
      __syncwarp(smask);
      continue;
    }
    const double chi_cur = cs[0], r2_cur = cs[1];
    const int accepts = (int)cs[2];
    bool acc = false;
    double lp_new = 0.0;
    if (valid) {
      const double u = odl_mh_uniform(A, chain, it + kk);
      if (!A.prior) {
        acc = exp(chi_cur - my_chi) > u;                           // Samplers.py:124-127 (NaN rejects)
      } else {
        double hast;
        const OdlSmemView pv{G.psm};
        odl_log_prior(A, pv, cur, lp_new, hast);
        acc = exp((chi_cur - my_chi) + (lp_new - cs[5]) + hast) > u;
      }
    }
    const unsigned bacc = __ballot_sync(smask, acc && lead) >> sbase;
    const int nvalid = __popc(__ballot_sync(smask, valid && lead));               // valid groups are a prefix
    const int jstar = bacc ? (__ffs(bacc) - 1) / ODL_G : -1;                      // first acceptance
    const int adv = (jstar >= 0) ? jstar + 1 : nvalid;
    const bool consumed = valid && kk < adv;
    const bool is_acc = consumed && kk == jstar;
    const bool dead = A.stop_failed && __ballot_sync(smask, consumed && lead && st.status != ODL_OK);
    const int src = sbase + (jstar >= 0 ? jstar : 0) * ODL_G;
    const double chi_acc = __shfl_sync(smask, my_chi, src), r2_acc = __shfl_sync(smask, my_r2, src);
    const double* pacc = G.psm + (long long)((jstar >= 0 ? jstar : kk) - kk) * (long long)group_doubles;   // accepted proposal (shared)
    if (consumed && lead) {
      if (A.step_count) atomicAdd((unsigned long long*)&A.step_count[chain], (unsigned long long)st.nsteps);
      if (A.fail_count && st.status != ODL_OK) atomicAdd(&A.fail_count[chain], 1);
      const int iter = it + kk;
      const long long k = (long long)chain * A.n_iter_total + (iter - 1);
      if (A.trace_chinew) A.trace_chinew[k] = my_chi;
      if (A.trace_accept) A.trace_accept[k] = is_acc ? 1 : 0;
      if (iter > A.burnin) {
        const int rowi = iter - A.burnin - 1;
        if (A.samples && rowi < A.n_keep) {
          double* rowp = A.samples + (long long)chain * A.smp_chain_pitch + (long long)rowi * A.smp_row_pitch;
          const double c = is_acc ? my_chi : chi_cur;
          for (int q = 0; q < ODL_P; ++q) rowp[q] = is_acc ? G.psm[q] : cur[q];
          rowp[ODL_P + 0] = c; rowp[ODL_P + 1] = is_acc ? my_r2 : r2_cur;
          rowp[ODL_P + 2] = 2.0 * c + 2.0 * (double)A.pnum;
          rowp[ODL_P + 3] = (double)iter;
          rowp[ODL_P + 4] = (double)(accepts + (is_acc ? 1 : 0)) / (double)iter;
        }
      }
    }
    if (kk == 0 && lead) {
      // Welford over ln(theta) of the kept rows, sequential in the iteration; best kept row (see odl_mcmc_body)
      if (A.summaries && it + adv - 1 > A.burnin) {
        double* sm = A.summaries + (size_t)chain * (1 + 2 * ODL_P);
        double cnt = sm[0];
        for (int i = 0; i < adv; ++i) {
          if (it + i <= A.burnin) continue;
          cnt += 1.0;
          for (int q = 0; q < ODL_P; ++q) {
            const double x = log((i == jstar) ? pacc[q] : cur[q]);
            const double dlt = x - sm[1 + q];
            const double mean = sm[1 + q] + dlt / cnt;
            sm[1 + q] = mean;
            sm[1 + ODL_P + q] += dlt * (x - mean);
          }
        }
        sm[0] = cnt;
      }
      const int first_kept = A.burnin + 1;
      double best_chi = cs[3], best_it = cs[4];
      if (it <= first_kept && first_kept < it + adv && first_kept - it != jstar) {
        best_chi = chi_cur; best_it = (double)first_kept;
        if (A.best_theta) for (int q = 0; q < ODL_P; ++q) A.best_theta[(size_t)chain * ODL_P + q] = cur[q];
      }
      if (jstar >= 0 && it + jstar > A.burnin && (chi_acc < best_chi || (best_chi != best_chi && chi_acc == chi_acc) || best_it == 0.0)) {
        best_chi = chi_acc; best_it = (double)(it + jstar);
        if (A.best_theta) for (int q = 0; q < ODL_P; ++q) A.best_theta[(size_t)chain * ODL_P + q] = pacc[q];
      }
      cs[3] = best_chi; cs[4] = best_it;
    }
    __syncwarp(smask);                                             // everything above read cur[] / cs[] before they change
    if (is_acc && lead) {
      cs[0] = my_chi; cs[1] = my_r2; cs[2] = (double)(accepts + 1);
      if (A.prior) cs[5] = lp_new;
      for (int q = 0; q < ODL_P; ++q) cur[q] = G.psm[q];
    }
    it += adv;
    __syncwarp(smask);
    if (dead) break;                                               // uniform over the lanes of the chain
  }
}
#endif  // unit 12
#endif  // !ODL_SMALL
#endif  // ODL_HOST_HARNESS
