// odl_abi.h -- plain-old-data kernel argument blocks shared by the host library (odl_capi.cu) and the
// device code (odl_kernels.cuh, compiled by NVRTC).  Only fixed-size members: the layout must not depend
// on the model so that one host build can drive every NVRTC-compiled model.
#ifndef ODL_ABI_H
#define ODL_ABI_H
#define ODL_MAX_WALK 64
#define ODL_CHAIN_STATE 8
#define ODL_LOGTAB 256        // intervals of [1, 2) in the scorer's logarithm table (odl_log): 2 doubles each, per CTA

// Cooperative kernels (n > 8, g lanes per system): doubles between the shared-memory rows (state | parameters | staging)
// of two groups.  The 32/g groups of a warp load the SAME index of their own rows at once (each a broadcast inside the
// group): the rows must start in different banks.  A row length divisible by g puts them in the same ones -- the 5x5
// network's rows were 35 + 40 + 181 = 256 doubles: every shared load of the right-hand side was served in 2-4 passes
// (9.7e9 bank conflicts for 3.3e9 LDS, profiles/r2j_mcmc_coop_network_full_ncu.txt) -- so such a row gets one more.
#define ODL_COOP_ROW(base, g) ((base) + (((g) < 32 && (base) % (g) == 0) ? 1 : 0))

// status words (per system)
#define ODL_OK 0
#define ODL_MAXSTEPS 1
#define ODL_NONFINITE 2
#define ODL_HUNDERFLOW 3
#define ODL_STIFF 4
#define ODL_ALLMASKED 8   // or-ed: every chi term was invalid (reference returns np.ma.masked)

struct OdlData {                 // constant tables of one ModelFramework
  const double* slot_t;          // [n_slot] distinct observation grid times, ascending
  const double* obs_lnO;         // [n_obs] ln(abundance)              (Framework.py:326)
  const double* obs_w;           // [n_obs] 1 / (2*sigma^2)            (stats.py:41; inf for sigma = 0: term masked)
  const double* obs_lin;         // [n_obs] exp(ln O)                  (Framework.py:700)
  const int* obs_src;            // [n_obs] slot*n_out + column
  const double* y0;              // [n_state]
  const int* y0_from_param;      // [n_state] parameter index supplying the initial value, or -1
  int n_slot;
  int n_obs;
  int stage_stride;              // doubles of staging per thread (odd, >= n_slot*n_out)
  int pad_;
  double t0;
  double sstot;                  // sum_s n_s * var(O_s)               (stats.py:55)
  double inv_sstot;              // 1 / sstot: R^2 = 1 - ssres * inv_sstot without a division per solve
};

struct OdlOpts {
  double rtol, atol;
  double h0;                     // 0 = automatic initial step
  double hmax;                   // 0 = t_end - t0
  int max_steps;
  int stiff_check;               // 1 = run Hairer's stiffness test and bail out with ODL_STIFF
  int stiff_min_steps;           // bail out only if more than this many steps of the current size remain
  int early_check_steps;         // > 0: DOPRI5 stops with ODL_MAXSTEPS already after this many attempts when the
                                 //      progress so far projects to more than max_steps attempts in total
  int lanes;                     // sweep kernels: lanes per warp that take systems (0 = all 32)
  int watchdog_spins;            // consumer: idle polls (~0.4 us each) of a warp before it gives up on the producer
  int explicit_cap;              // odl_mcmc_auto_kernel: attempted DOPRI5 steps a solve may take before it is redone on
                                 //   BDF (0 = max_steps; Hairer's test alone routes)
  int pad_;
};

struct OdlSweepArgs {
  const double* theta;           // [n][n_param]
  const int* index;              // optional indirection (stiff list): system i reads theta[index[i]]
  const int* index_count;        // optional device-side count for `index` (overrides n)
  long long n;
  double* chi;                   // [n]
  double* r2;                    // [n]
  int* status;                   // [n]
  int* nsteps;                   // [n]  attempted steps
  double* pred;                  // optional [n][n_obs] predictions at the observation rows
  unsigned long long* counter;   // work counter, zeroed by the host before launch
  int* defer_list[2];            // optional: rows that stopped with [0] ODL_MAXSTEPS, [1] ODL_STIFF are appended
  int* defer_count[2];           //           here for a later pass (device-side lists, no host round trip)
  // The stiff pass BESIDE the bulk pass (ODL_SOLVER_AUTO on SMs of its own): consumer of a feed that is still growing
  unsigned long long* feed_ticket;         // consumer mode switch: next unclaimed entry of index[] (entries of index[]
                                           //   start as -1 and land while the bulk pass runs; *index_count grows; the
                                           //   consumer overwrites an entry with -2 when it takes it)
  const int* feed_done;                    // set to 1 by odl_feed_done_kernel after the last bulk launch of the sweep
  int* watchdog;                           // consumer: incremented when a warp gave up waiting for the feed
  int* resident;                           // consumer: +1 per CTA once it runs (odl_gate_kernel holds the bulk launch
                                           //   back until every consumer CTA has its SM)
  double* handover;                        // optional [n][handover_stride]: what the DOPRI5 pass had reached when it gave a
  int handover_stride;                     //   row up -- t, next slot, y[n_state], the staged observation columns of the slots
  int pad3_;                               //   behind it -- so that the stiff pass continues from there instead of from t0
  long long* timeline;                     // development (kernels built with -DODL_TIMELINE=1, else unused): per feed
                                           //   entry %globaltimer at [0] deferral, [1] start and [2] end of its stiff solve
};

// Cost ordering of a sweep (ODL_SOLVER_AUTO): key = |J(t0, y0, theta)|_inf (t_end - t0), quarter-octave bins,
// processed from the highest bin down -- the systems that need the most steps (and those the stiff pass will
// take over) start first, so nothing long is left for the end of the launch.
#define ODL_ORDER_BINS 256
struct OdlOrderArgs {
  const double* theta;           // [n][n_param]
  long long n;
  unsigned char* bins;           // [n]
  int* hist;                     // [ODL_ORDER_BINS] zeroed by the host
  int* cursor;                   // [ODL_ORDER_BINS] start of every bin in index[] (descending bins), then a cursor
  int* index;                    // [n] out: rows in processing order
  int row_base;                  // added to every row number written to index[] (theta / bins / index point at the
  int pad_;                      //   start of a chunk of a larger sweep)
};

struct OdlTrajArgs {
  const double* theta;           // [n][n_param]
  const double* y0;              // optional [n][n_state] per-system initial state
  long long n;
  double* traj;                  // [n][n_slot][n_state]  raw states on the output grid
  int* status;
  int* nsteps;
  unsigned long long* counter;
};

struct OdlMcmcArgs {
  double* theta_cur;             // [C][n_param] in: chain starts / current points; out: current points
  double* chain_state;           // [C][ODL_CHAIN_STATE] chi_cur, r2_cur, accepts, best_chi, best_iteration, 3 unused
                                 //   (persist across launches); best_* = first minimum of chi over the kept rows
  int n_chain;
  int chain_offset;              // global index of chain 0 of this launch (multi-GPU sharding / RNG key)
  int it_begin, it_end;          // iterations [it_begin, it_end) of Samplers.py:104; it_begin==1 => a-priori solve first
  int burnin;
  int n_keep;                    // rows per chain in `samples` (= nits-1-burnin)
  int row_stride;                // doubles per sample row (>= n_param + 5)
  int rng_mode;                  // 0 Philox, 1 host z/u streams, 2 teacher-forced proposals + u
  int n_walk;
  int pnum;                      // for AIC = 2 chi + 2 pnum   (stats.py:44-47)
  int walk[ODL_MAX_WALK];        // indices of walking parameters (parameter_names order)
  double step_sd;                // 0.05 (Framework.py:107)
  unsigned long long seed;
  const double* z;               // [C][n_iter_total][n_walk]   (rng_mode 1)
  const double* u;               // [C][n_iter_total]           (rng_mode 1,2)
  const double* forced;          // [C][n_iter_total][n_param]    (rng_mode 2)
  int n_iter_total;              // nits-1
  int spec;                      // lanes per chain (power of two <= 32): iterations evaluated at once along the
                                 //   all-rejected path (prefetching MH); 1 = one proposal at a time
  double* samples;               // kept rows (theta.., chi, rsquared, aic, iteration, acceptance_ratio): row (chain, r) at
  long long smp_chain_pitch;     //   samples[chain * smp_chain_pitch + r * smp_row_pitch] (doubles): chain-major
  long long smp_row_pitch;       //   [C][n_keep][row_stride] or iteration-major [n_keep][C][row_stride]
  double* summaries;             // [C][1+2*n_param]: count, mean[P], M2[P] of ln(theta) over kept rows
  double* trace_chinew;          // optional [C][n_iter_total]  chi of every proposal (parity tests)
  unsigned char* trace_accept;   // optional [C][n_iter_total]
  int* fail_count;               // optional [C] proposals whose solve failed
  long long* step_count;         // optional [C] attempted integrator steps (flop accounting)
  double* best_theta;            // optional [C][n_param] parameters of the best kept row (Framework.py:725-731)
  const long long* chain_ids;    // optional [C] global chain index of every local chain (keys the Philox stream);
                                 //   default chain_offset + local index
  const double* prior;           // optional [n_param][4] (kind, a, b, c) as odl_sample_lhs: with it the acceptance ratio
                                 //   is the posterior's, exp((chi-chinew) + (lp'-lp) + sum ln(theta'/theta)); without,
                                 //   the reference's (priors never enter, Samplers.py:118-127); lp of the current point
                                 //   lives in chain_state[5]
  int stop_failed;               // 1 = a chain stops at its first consumed solve that failed (the caller runs such chains
  int pad2_;                     //   again with another stepper: what they would still compute is thrown away)
};

#endif  // ODL_ABI_H
