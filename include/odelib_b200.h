/* odelib_b200.h -- C ABI of the B200-native ODElib hot path (libodelib_b200.so).
 *
 * Plain C, plain pointers and sizes; no torch / C++ types cross this boundary.  The reference
 * (SEpapoulis/ODElib 0.1.1) is pure Python and has NO FFI of its own; each entry point below
 * replaces one of its natural batch seams (paths are into /root/reference):
 *
 *   odl_model_create     <- the user's RHS callable handed to ModelFramework (ODElib/Framework.py:168,
 *                           :212) and evaluated by scipy odeint (Framework.py:656).  Here: CUDA source
 *                           of the traced RHS, NVRTC-compiled for sm_100a together with the integrators.
 *   odl_model_set_data   <- ModelFramework.__init__ data setup: _formatdf/_df_fitsetup
 *                           (Framework.py:281-329), times (:234), inits (:246-258, :496-510),
 *                           '<state>0' parameters (Statistics/Samplers.py:110-114).
 *   odl_model_set_grid   <- self.times, the output grid of integrate() (Framework.py:234, :241).
 *   odl_sweep            <- _Fit_worker(model, parameter_list) (Framework.py:41-48): for every
 *                           parameter set integrate(predict_obs=True) + get_chi (:622-697), plus
 *                           get_Rsqrd (:699-702).
 *   odl_trajectory       <- ModelFramework.integrate(...) full-grid result (Framework.py:656, :683).
 *   odl_mcmc             <- Samplers.MetropolisHastings (Statistics/Samplers.py:53-174) for many
 *                           chains at once; _Chain_worker / the serial loop (Framework.py:19-22,
 *                           :1025-1030).
 *   odl_select_below,    <- chain-start selection of MCMC(chain_inits=int): the threshold filter and the
 *   odl_gather_rows         resampling of acceptable survey rows (Framework.py:993-1016).
 *   odl_sample_lhs       <- Samplers.sample_lhs (Statistics/Samplers.py:6-51) under _lhs_samples
 *                           (Framework.py:589-615), for large surveys.
 *   odl_reference_streams,        <- numpy's legacy RandomState(seed) exactly as MetropolisHastings consumes it per
 *   odl_reference_streams_device     iteration (Samplers.py:70, :108, :118-121, :127; Framework.py:103, :119): fed to
 *                                    odl_mcmc (ODL_RNG_HOST_STREAMS) they make every chain the reference chain of its seed.
 *   odl_comm_*, odl_rhat <- the gather of the workers' results (Framework.py:1035-1038); R-hat itself is new.
 *   odl_fp64_peak        <- (no reference counterpart) measures the FP64 FMA roofline denominator.
 *
 * Conventions: every function returns 0 on success or an ODL_E* code; odl_last_error() gives the
 * message of the calling thread's last failure.  Numerical failure of a single system is NOT an
 * error: it is reported in that system's status word and its chi is NaN (the reference silently
 * yields NaN / masked there and the sampler rejects, Samplers.py:127).  The caller owns all buffers.
 * `mem` says where the caller's arrays live: ODL_MEM_HOST (library stages them through its own
 * device scratch; copies are inside the call) or ODL_MEM_DEVICE (CUDA device pointers, e.g. torch
 * tensor .data_ptr(); nothing is copied).  `stream` is a cudaStream_t (NULL = default stream).
 * Calls are asynchronous only for ODL_MEM_DEVICE; host-memory calls return after their results
 * have landed.  Handles are not thread-safe and allow ONE call in flight: the per-model scratch buffers, the
 * counter block, the events and the helper stream are shared by all calls on a handle, so a second call (from another
 * host thread, or an ODL_MEM_DEVICE call on another stream) must wait until the previous one's work has completed.
 * Use one handle per host thread / per GPU.  Every entry point runs on the model's device and restores the calling
 * thread's current CUDA device before it returns.
 */
#ifndef ODELIB_B200_H
#define ODELIB_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define ODL_ABI_VERSION 3

enum { ODL_SUCCESS = 0, ODL_EINVAL = 1, ODL_ECUDA = 2, ODL_ECOMPILE = 3, ODL_ENODEVICE = 4, ODL_EIO = 5 };
enum { ODL_MEM_HOST = 0, ODL_MEM_DEVICE = 1 };
enum { ODL_RHAT_LOCAL = 256 };  /* odl_rhat only, or-ed into `mem`: reduce THIS rank's chains, no collective, whatever
                                   communicator the handle has joined */
enum { ODL_SOLVER_DOPRI5 = 0, ODL_SOLVER_ROS23 = 1, ODL_SOLVER_AUTO = 2, ODL_SOLVER_RADAU5 = 3, ODL_SOLVER_BDF = 4 };
/* odl_solver_opts.auto_flags */
enum { ODL_AUTO_UNORDERED = 1,   /* process rows in input order (no cost ordering) */
       ODL_AUTO_CONCURRENT = 2,  /* run the stiff pass BESIDE the DOPRI5 pass on SMs of its own whatever the table size
                                    (default: from 262,144 rows on) */
       ODL_AUTO_ONE_PIECE = 4,   /* ODL_MEM_HOST: upload theta in one piece before anything runs (default: two pieces,
                                    the second travelling while the first is swept) */
       ODL_AUTO_SEQUENTIAL = 8,  /* run the stiff pass AFTER the DOPRI5 pass (single-warp CTAs over every SM) */
       ODL_AUTO_NO_HELPER = 16,  /* stiff pass beside the DOPRI5 pass: no second consumer on the SMs that pass frees when it
                                    ends (development: the state of the code before that helper existed) */
       ODL_AUTO_NO_HANDOVER = 32 /* the stiff pass integrates its rows again from t0 instead of continuing from where the
                                    DOPRI5 pass stopped (the rule of the code before the hand-over existed) */ };
enum { ODL_RNG_PHILOX = 0, ODL_RNG_HOST_STREAMS = 1, ODL_RNG_FORCED = 2 };
enum { ODL_SAMPLES_CHAIN_MAJOR = 0, ODL_SAMPLES_ITERATION_MAJOR = 1 };
/* per-system status words */
enum { ODL_ST_OK = 0, ODL_ST_MAXSTEPS = 1, ODL_ST_NONFINITE = 2, ODL_ST_HUNDERFLOW = 3, ODL_ST_STIFF = 4,
       ODL_ST_ALLMASKED = 8 };

typedef struct odl_model odl_model;

typedef struct odl_build_opts {
  int device;          /* CUDA ordinal; -1 = current device */
  int block_threads;   /* threads per CTA, 0 = default (128) */
  int min_blocks;      /* __launch_bounds__ min CTAs per SM, 0 = default (4) */
  int dense_output;    /* 1 = dense output at the observation times (default), 0 = land on them */
  int compile_only;    /* 1 = NVRTC-compile every kernel unit (and cache them) but do not touch a GPU; 2 = the same for
                          the units of the default paths only */
  int y0_from_param;   /* 1 = some state's initial value is a parameter ('<state>0', Samplers.py:110-114).  As in the
                          reference this holds for the solves of MCMC PROPOSALS only: odl_sweep, odl_trajectory and a
                          chain's a-priori solve start from y0 (istates, Framework.py:647-650; Samplers.py:88) */
  int coop_lanes;      /* n_state > 8: lanes per system of the cooperative kernels (2, 4, 8, 16 or 32);
                          0 = by state count, the fewest lanes with at most 9 components per lane (2 up to 18 states,
                          4 up to 36, 8 up to 72, 16 up to 144, else 32) */
  int reserved[1];
  const char* cache_dir; /* directory for compiled cubins, NULL = no cache */
} odl_build_opts;

typedef struct odl_solver_opts {
  double rtol, atol;   /* scipy odeint defaults are 1.49012e-8 (Framework.py:656) */
  double h0, hmax;     /* 0 = automatic */
  int max_steps;       /* attempted steps per solve before ODL_ST_MAXSTEPS */
  int solver;          /* ODL_SOLVER_* */
  int stiff_check;     /* DOPRI5: detect stiffness and stop with ODL_ST_STIFF */
  int stiff_min_steps; /* ... only while more than this many steps of the current size remain (0 = 2000) */
  int pass_cap0;       /* ODL_SOLVER_AUTO: step cap of the first DOPRI5 pass (odl_sweep: 0 = 704; odl_mcmc: see there) */
  int tail_warps;      /* ODL_SOLVER_AUTO: SMs set aside for the stiff pass when it runs beside the DOPRI5 pass (one CTA
                          of 8 warps each; 0 = 22 % of them) */
  int tail_solver;     /* ODL_SOLVER_AUTO: stepper of the pass over what DOPRI5 did not finish:
                          0 = default (ODL_SOLVER_BDF), or ODL_SOLVER_RADAU5 */
  int early_check_steps; /* ODL_SOLVER_AUTO: the first pass drops a system after this many attempts when its progress
                          projects beyond pass_cap0 (0 = 3/4 of pass_cap0 but at most 384, -1 = never) */
  int tail_lanes;      /* ODL_SOLVER_AUTO: lanes per warp that take systems in the stiff pass (0 = as few as spreading its
                          systems over every resident warp takes; 32 = full warps) */
  int auto_flags;      /* ODL_SOLVER_AUTO: ODL_AUTO_* bits, 0 = cost-ordered DOPRI5 pass, then the stiff pass */
} odl_solver_opts;

typedef struct odl_mcmc_opts {
  int n_chain;
  int chain_offset;    /* global index of the first chain (multi-GPU shard; keys the Philox stream) */
  int nits;            /* as Samplers.MetropolisHastings(nits): nits-1 iterations */
  int burnin;          /* rows are kept iff iteration > burnin */
  int it_begin, it_end;/* run iterations [it_begin, it_end); 0,0 = whole chain [1, nits) */
  int rng_mode;        /* ODL_RNG_* */
  int n_walk;          /* number of walking (non-static) parameters */
  const int* walk;     /* [n_walk] their indices in parameter_names order */
  int pnum;            /* _pnum for AIC (Framework.py:261-263) */
  int row_stride;      /* doubles per sample row, >= n_param+5 */
  double step_sd;      /* 0.05 (Framework.py:107) */
  unsigned long long seed;
  int speculate;       /* lanes per chain that evaluate consecutive iterations at once along the all-rejected path
                          (prefetching MH; a power of two <= 32).  The chain is the same for every value;
                          0 = automatic (fills an otherwise latency-bound GPU), 1 = one proposal at a time.
                          n_state > 8: values >= 1 select the thread-per-system kernel; 0 and negative values the
                          cooperative kernel, -K = K groups of coop_lanes lanes per chain (K * coop_lanes <= 32) */
  int sample_layout;   /* ODL_SAMPLES_CHAIN_MAJOR: samples[chain][row][row_stride] (the reference frame's order);
                          ODL_SAMPLES_ITERATION_MAJOR: samples[row][chain][row_stride] -- the rows a warp keeps in one
                          iteration are contiguous and leave as coalesced full-sector stores */
  int stop_failed_chains; /* 1 = a chain stops at the first consumed solve that failed (fail_count > 0 marks it; its other
                          outputs are then incomplete).  For callers that run such chains again with another stepper --
                          the facade's solver="auto" does -- so that they do not burn max_steps on every later proposal
                          while the chains beside them wait; the chains that never fail are not affected */
  int reserved;
} odl_mcmc_opts;

typedef struct odl_mcmc_io {
  double* theta;           /* [n_chain][n_param] in: starts (or current points), out: current points */
  double* chain_state;     /* [n_chain][8] chi, r2, accepts, best_chi, best_iteration, 3 unused; in when
                              it_begin>1, always out.  best_* = the first minimum of chi over the kept rows
                              (what set_best_params' idxmin picks, Framework.py:725-731); best_iteration 0 = none */
  double* samples;         /* [n_chain][nits-1-burnin][row_stride] (or iteration-major, see sample_layout) or NULL:
                              theta.., chi, rsquared, aic, iteration, acceptance_ratio (Samplers.py:160-165) */
  double* summaries;       /* [n_chain][1+2*n_param] count, mean, M2 of ln(theta) over kept rows, or NULL
                              (must be zeroed by the caller before the first segment) */
  const double* z;         /* [n_chain][nits-1][n_walk]  proposal increments (ODL_RNG_HOST_STREAMS) */
  const double* u;         /* [n_chain][nits-1]          acceptance uniforms (HOST_STREAMS, FORCED) */
  const double* forced;    /* [n_chain][nits-1][n_param] proposals (ODL_RNG_FORCED) */
  double* trace_chinew;    /* optional [n_chain][nits-1] */
  unsigned char* trace_accept; /* optional [n_chain][nits-1] */
  int* fail_count;         /* optional [n_chain], accumulated */
  long long* step_count;   /* optional [n_chain], accumulated attempted integrator steps */
  double* best_theta;      /* optional [n_chain][n_param] parameters of the best kept row */
  const long long* chain_ids; /* optional [n_chain] global chain index per chain (Philox key) for a batch that is not a
                              contiguous block of chains, e.g. the chains re-run with another stepper; default
                              chain_offset + local index */
  const double* prior_table; /* optional [n_param][4] = (kind, a, b, c) per parameter, kinds and parameterisation as
                              odl_sample_lhs.  NULL (default) = the reference's chain: prior densities never enter the
                              acceptance ratio (Samplers.py:118-127; the reference evaluates them and drops them).
                              Given: Metropolis-Hastings on the posterior, exp((chi-chinew) + (lp'-lp) + sum
                              ln(theta'/theta)) > u -- prior log-densities evaluated in the kernel, Hastings term of the
                              multiplicative walk included.  chain_state[5] then holds lp of the current point. */
} odl_mcmc_io;

int odl_abi_version(void);
const char* odl_last_error(void);

int odl_model_create(const char* model_cuda_src, int n_state, int n_param, int n_out,
                     const odl_build_opts* opts, odl_model** out);
int odl_model_destroy(odl_model* m);
/* compile log of the NVRTC run (warnings included); valid until the model is destroyed */
const char* odl_model_build_log(const odl_model* m);
/* resource usage of a compiled kernel ("sweep", "mcmc", "traj", "sweep_ros23", "mcmc_ros23", "mcmc_auto",
   "sweep_radau5", "mcmc_radau5", "sweep_bdf", "mcmc_bdf", and for n_state > 8 "sweep_coop", "mcmc_coop"):
   registers/thread, local (spill) bytes, resident CTAs per SM */
int odl_model_kernel_info(odl_model* m, const char* kernel, int* regs, int* local_bytes, int* max_blocks_per_sm);
/* Kernels are compiled per unit (one NVRTC program per kernel: "sweep", "traj", "mcmc", "sweep_bdf", "mcmc_bdf",
   "sweep_ros23", "mcmc_ros23", "mcmc_auto", "sweep_radau5", "mcmc_radau5", "sweep_coop", "mcmc_coop", "order"), in
   parallel host threads: odl_model_create starts the units of the default paths (ordering, DOPRI5 sweep, BDF stiff pass,
   trajectories, chains) and returns; a call waits for the units it launches and compiles any other on first use.
   *seconds = NVRTC time of the unit (cache hits: the file read), -1 when it has not been compiled; never compiles. */
int odl_model_unit_seconds(const odl_model* m, const char* unit, double* seconds, int* cache_hit_or_null);

int odl_model_set_data(odl_model* m, int n_slot, const double* slot_time, int n_obs, const int* obs_slot,
                       const int* obs_col, const double* ln_obs, const double* log_sigma, double sstot,
                       const double* y0, const int* y0_from_param, double t0);
int odl_model_set_grid(odl_model* m, int n_t, const double* times, const double* y0, const int* y0_from_param);

/* chi[n] is required; r2[n], status[n], nsteps[n], pred[n][n_obs] are optional (NULL = not wanted: neither written by
   the kernels nor copied back).  A solve that failed has chi = NaN whether or not its status word is asked for. */
int odl_sweep(odl_model* m, const odl_solver_opts* so, long long n, const double* theta, int mem,
              double* chi, double* r2, int* status, int* nsteps, double* pred_or_null, void* stream);
int odl_trajectory(odl_model* m, const odl_solver_opts* so, long long n, const double* theta,
                   const double* y0_or_null, int mem, double* traj, int* status, int* nsteps, void* stream);
/* odl_mcmc with so->solver = ODL_SOLVER_AUTO: every solve of a chain (a-priori point and proposals) is attempted with
   DOPRI5 and, when that gives up, done again with the variable-order BDF stepper -- the per-solve method switch LSODA makes
   for the reference (Framework.py:656).  "Gives up": so->pass_cap0 > 0 = that many attempted steps (Hairer's stiffness
   test only with so->stiff_check); pass_cap0 = 0 = Hairer's test (and max_steps).  Which stepper finishes a solve depends
   on that solve alone, so a chain whose solves all stay within pass_cap0 is the ODL_SOLVER_DOPRI5 chain run with
   max_steps = pass_cap0 (same decisions; chi to rounding -- the two kernels are compiled separately).  fail_count counts
   solves that neither stepper finished. */
int odl_mcmc(odl_model* m, const odl_solver_opts* so, const odl_mcmc_opts* mo, const odl_mcmc_io* io, int mem,
             void* stream);

/* Chain-start selection on the device (replaces the pandas filter of Framework.py:1004-1012).
   odl_select_below: index_dev[0..*count_host) <- the rows i of chi_dev[0..n) with chi < cut, ascending (NaN never
   qualifies); returns after *count_host is valid.  odl_gather_rows: dst_dev[r][:] = src_dev[row(r)][:] with
   row(r) = index_dev[picks_host[r]] (or picks_host[r] when index_dev is NULL); picks are drawn by the caller
   (the reference draws them with DataFrame.sample(n, replace=True), Framework.py:1012). */
int odl_select_below(odl_model* m, const double* chi_dev, long long n, double cut, int* index_dev, long long* count_host,
                     void* stream);
int odl_gather_rows(odl_model* m, const double* src_dev, int row_len, const int* index_dev_or_null,
                    const long long* picks_host, long long n_pick, double* dst_dev, void* stream);

/* Latin-hypercube sample of the priors on the device (replaces Samplers.sample_lhs, Samplers.py:6-51, for large
   surveys): theta_dev[i][j] = ppf_j((stratum_j(i) + U) / n) with stratum_j a keyed permutation of 0..n-1 per
   column.  kind[j]: 0 constant a[j]; 1 lognorm(s = a[j], loc = b[j], scale = c[j]); 2 norm(loc = b[j], scale = c[j]);
   3 uniform(loc = b[j], scale = c[j]) -- scipy.stats parameterisation.  Deterministic in (seed, n). */
int odl_sample_lhs(odl_model* m, long long n, int n_param, const int* kind, const double* a, const double* b,
                   const double* c, unsigned long long seed, double* theta_dev, void* stream);

/* The reference chain's own random numbers (host code; works without a GPU): for every chain c, numpy's legacy
   RandomState(seeds[c]) consumed as Samplers.MetropolisHastings consumes it per iteration (Samplers.py:70, :108,
   :118-121, :127): n_walk normals N(0, step_sd) -> z[c][i][:], n_prior_draws discarded standard normals (the `rvs` of
   lognorm / norm priors inside the unused pdf() calls), one uniform -> u[c][i].  z [n_chain][n_iter][n_walk],
   u [n_chain][n_iter]; feed them to odl_mcmc with ODL_RNG_HOST_STREAMS. */
int odl_reference_streams(const unsigned int* seeds, int n_chain, int n_iter, int n_walk, int n_prior_draws,
                          double step_sd, double* z, double* u);
/* The same streams generated on the device (one thread per chain, MT19937 key arrays in device memory), for runs whose
   streams would take the host generator longer than the chains take the GPU: z_dev [n_chain][n_iter][n_walk], u_dev
   [n_chain][n_iter] device buffers, seeds on the host.  Uniforms are bit-identical to numpy's; a gaussian goes through
   log(), where CUDA's and glibc's -- both within 1 ulp -- may round differently: z agrees to 1 ulp. */
int odl_reference_streams_device(odl_model* m, const unsigned int* seeds_host, int n_chain, int n_iter, int n_walk,
                                 int n_prior_draws, double step_sd, double* z_dev, double* u_dev, void* stream);

/* The one collective of the path (the reference gathers its workers' frames with pd.concat,
   Framework.py:1035-1038, and has no R-hat): Gelman-Rubin R-hat on ln(theta) over the chains of EVERY rank.
   odl_comm_unique_id: rank 0 draws an NCCL id (ODL_COMM_ID_BYTES bytes) and hands it to the other ranks by whatever
   means the launcher offers; odl_comm_init: every rank joins (ncclCommInitRank on the model's device; world 1 needs no
   NCCL); odl_rhat: ncclAllGather of the per-chain summaries (count, mean[P], M2[P] -- odl_mcmc_io.summaries; shards may
   differ in length) over NVLink, reduction on the device: rhat_host[P] = sqrt(((n-1)/n W + B/n) / W) with W = mean_j
   s_j^2, B = n var_j(mean_j) (ddof 1); pooled_host[1+2P] = count, mean[P], M2[P] of ALL kept rows pooled (the two
   moments the fitting report's rawstats takes from the frame, Framework.py:11-17).  Without odl_comm_init: the local
   chains alone.  NCCL is dlopen()ed at the first use; the library has no link dependency on it. */
#define ODL_COMM_ID_BYTES 128
int odl_comm_unique_id(unsigned char* id128);
int odl_comm_init(odl_model* m, const unsigned char* id128, int world, int rank);
int odl_comm_destroy(odl_model* m);
int odl_rhat(odl_model* m, const double* summaries, int n_chain_local, int n_param, int mem, double* rhat_host,
             double* pooled_host_or_null, long long* n_chain_total_or_null, void* stream);

/* device time (ms) of the kernels launched by the last odl_sweep/odl_mcmc/odl_trajectory call on this
   model, measured with CUDA events on the launching stream; blocks until they have completed */
int odl_model_last_kernel_ms(odl_model* m, float* ms);
/* the same split for the last ODL_SOLVER_AUTO sweep: ms3[0] cost ordering (of the first piece), ms3[1] DOPRI5 bulk
   pass, ms3[2] stiff pass -- when it runs beside the bulk pass: what it still needed after the bulk pass had ended
   (single-pass calls: ms3[0]) */
int odl_model_last_pass_ms(odl_model* m, float* ms3);
/* development aid: copies the first `count` ints of the device-side counter block of the last odl_sweep
   ([0] work counter, [16] feed count, [32] feed ticket, [48] warps entered, [64] warps left, [80] watchdog) */
int odl_debug_counters(odl_model* m, int* out, int count);
/* development aid: with ODL_TIMELINE=1 in the environment and kernels built with -DODL_TIMELINE=1 (ODL_KERNEL_DEFINES),
   an AUTO sweep records per feed entry %globaltimer (ns) at deferral, start and end of its stiff solve */
int odl_debug_timeline(odl_model* m, long long* out, long long entries);
/* number of kernel launches issued by this library in this process */
long long odl_launch_count(void);

/* FP64 roofline denominator: sustained DFMA throughput of `device` in TFLOP/s (FMA = 2 flop) */
int odl_fp64_peak(int device, int repeats, double* tflops, float* ms_per_launch);

#ifdef __cplusplus
}
#endif
#endif /* ODELIB_B200_H */
