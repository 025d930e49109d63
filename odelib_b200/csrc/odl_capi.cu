// odl_capi.cu -- host side of libodelib_b200.so: the C ABI declared in include/odelib_b200.h.
//
// Runtime API (static cudart) for memory / streams / events; NVRTC compiles the traced model + the
// integrator kernels (odl_kernels.cuh, embedded below) to an sm_100a cubin; the few driver entry
// points needed to load and launch that cubin are fetched through cudaGetDriverEntryPoint, so the
// library has no link-time dependency on libcuda and loads (but cannot compute) on a GPU-less host.
#include <cuda.h>
#include <cuda_runtime.h>
#include <nvrtc.h>
#include <nccl.h>     // types and prototypes only: the library is dlopen()ed at the first collective (no link dependency)
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <unistd.h>
#include <cstring>
#include <future>
#include <string>
#include <thread>
#include <vector>

#include "../../include/odelib_b200.h"
#include "odl_abi.h"

static const char* kAbiHeaderSrc =
#include "odl_abi_embedded.inc"
    ;
static const char* kKernelSrc =
#include "odl_kernels_embedded.inc"
    ;

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define ODL_CUDA(call)                                                                          \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess)                                                                      \
      return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? ODL_ENODEVICE : ODL_ECUDA, \
                  std::string(#call) + ": " + cudaGetErrorString(e_));                          \
  } while (0)

// ---------------------------------------------------------------------------------------------
// driver entry points (module load + launch of NVRTC output)
// ---------------------------------------------------------------------------------------------
struct Driver {
  CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
  CUresult (*ModuleUnload)(CUmodule) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream,
                           void**, void**) = nullptr;
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
  CUresult (*FuncGetAttribute)(int*, CUfunction_attribute, CUfunction) = nullptr;
  CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  CUresult (*LaunchKernelEx)(const CUlaunchConfig*, CUfunction, void**, void**) = nullptr;
  bool ok = false;
};
static Driver g_drv;

template <class F>
static bool entry(const char* name, F& fn) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
    return false;
  fn = reinterpret_cast<F>(p);
  return true;
}
static int load_driver() {
  if (g_drv.ok) return 0;
  bool ok = entry("cuModuleLoadData", g_drv.ModuleLoadData) && entry("cuModuleUnload", g_drv.ModuleUnload) &&
            entry("cuModuleGetFunction", g_drv.ModuleGetFunction) && entry("cuLaunchKernel", g_drv.LaunchKernel) &&
            entry("cuFuncSetAttribute", g_drv.FuncSetAttribute) && entry("cuFuncGetAttribute", g_drv.FuncGetAttribute) &&
            entry("cuOccupancyMaxActiveBlocksPerMultiprocessor", g_drv.OccupancyMaxActiveBlocksPerMultiprocessor) &&
            entry("cuGetErrorString", g_drv.GetErrorString) && entry("cuLaunchKernelEx", g_drv.LaunchKernelEx);
  if (!ok) {
    cudaGetLastError();
    return fail(ODL_ENODEVICE, "CUDA driver entry points unavailable (no GPU / driver on this host)");
  }
  g_drv.ok = true;
  return 0;
}
static std::string cu_err(CUresult r) {
  const char* s = nullptr;
  if (g_drv.GetErrorString) g_drv.GetErrorString(r, &s);
  return s ? s : "unknown CUDA driver error";
}
#define ODL_CU(call)                                                                  \
  do {                                                                                \
    CUresult r_ = (call);                                                             \
    if (r_ != CUDA_SUCCESS) return fail(ODL_ECUDA, std::string(#call) + ": " + cu_err(r_)); \
  } while (0)

// ---------------------------------------------------------------------------------------------
// model handle
// ---------------------------------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = std::max<size_t>(bytes, 256);
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) return fail(ODL_ECUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    cap = want;
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct Tables {
  DevBuf buf;
  OdlData d{};
  bool set = false;
};

// One NVRTC program per kernel ("unit", see ODL_UNIT in odl_kernels.cuh): compiled on demand, in parallel host
// threads, cached per unit -- a new model is ready as soon as the kernels the first call needs are, instead of after
// every stepper variant has been compiled (round 1: 19 s for two_i, 35 s for the 35-state network).
enum { U_SWEEP = 1, U_TRAJ = 2, U_MCMC = 3, U_SWEEP_BDF = 4, U_MCMC_BDF = 5, U_SWEEP_ROS = 6, U_MCMC_ROS = 7, U_MCMC_AUTO = 8,
       U_SWEEP_RADAU = 9, U_MCMC_RADAU = 10, U_SWEEP_COOP = 11, U_MCMC_COOP = 12, U_ORDER = 13, U_COUNT = 14 };
static const char* kUnitName[U_COUNT] = {"", "sweep", "traj", "mcmc", "sweep_bdf", "mcmc_bdf", "sweep_ros23", "mcmc_ros23",
                                         "mcmc_auto", "sweep_radau5", "mcmc_radau5", "sweep_coop", "mcmc_coop", "order"};
struct Compiled {
  int rc = 0;
  std::vector<char> cubin;
  std::string log, err;
  double seconds = 0.0;
  bool cache_hit = false;
};
struct Unit {
  std::future<Compiled> job;     // compile in flight (valid() until collected)
  Compiled c;
  bool have = false;             // c holds a cubin
  CUmodule mod = nullptr;
};

struct odl_model {
  int device = 0;
  int n_state = 0, n_param = 0, n_out = 0;
  int block = 128, minblocks = 4, dense = 1, y0p = 0, coop = 0;
  int sm_count = 0;
  bool on_gpu = false;
  std::string src, cache_dir;
  std::vector<std::string> opt;  // NVRTC options common to all units
  std::string log;
  Unit units[U_COUNT];
  CUfunction k_sweep = nullptr, k_traj = nullptr, k_mcmc = nullptr;
  CUfunction k_sweep_ros = nullptr, k_mcmc_ros = nullptr, k_mcmc_auto = nullptr;
  CUfunction k_sweep_radau = nullptr, k_mcmc_radau = nullptr;
  CUfunction k_sweep_bdf = nullptr, k_mcmc_bdf = nullptr;
  CUfunction k_order_key = nullptr, k_order_scan = nullptr, k_order_scatter = nullptr, k_feed_done = nullptr, k_gate = nullptr;
  CUfunction k_sweep_coop = nullptr, k_mcmc_coop = nullptr;   // n > 8 only: several lanes per system
  Tables data, grid;
  DevBuf counter;
  DevBuf scratch[24];                        // [0, 20): staging slots of one call; 21-23: select / gather / sample helpers
  DevBuf mt_state;                           // odl_reference_streams_device: MT19937 key arrays, [624][n_chain] words
  DevBuf handover;                           // AUTO sweeps: per row, what the DOPRI5 pass hands to the stiff pass
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t evp[2] = {nullptr, nullptr};   // between the cohort passes of an AUTO sweep
  cudaEvent_t ev_aux = nullptr, ev_fork = nullptr;
  cudaEvent_t ev_chunk[3] = {nullptr, nullptr, nullptr};   // host-memory sweeps: theta arrives in pieces on the helper stream
  cudaStream_t aux = nullptr;                // helper stream: theta pieces of a host-memory sweep
  cudaStream_t aux2 = nullptr;               // helper stream: the stiff pass beside the DOPRI5 pass
  cudaStream_t piece_stream[2] = {nullptr, nullptr};   // host-memory sweeps: the later pieces are ordered and swept here
  cudaEvent_t ev_piece[2] = {nullptr, nullptr};        //   ... and these mark the end of their bulk launches
  ncclComm_t comm = nullptr;                 // odl_comm_init: the ranks that share an MCMC run (R-hat all-gather)
  int comm_world = 1, comm_rank = 0;
  int n_pass = 0;
  bool timed = false;
  bool coop_model() const { return n_state > 8; }
};

// The calling thread's current device is the caller's business (torch keeps its own notion of it): every entry point
// switches to the model's device for its own duration only.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
    if (prev != device) { err = cudaSetDevice(device); switched = (err == cudaSuccess) && prev >= 0; }
  }
  ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};
#define ODL_ON_DEVICE(m)                                                                                   \
  DeviceGuard guard_((m)->device);                                                                         \
  if (guard_.err != cudaSuccess) return fail(ODL_ENODEVICE, std::string("cudaSetDevice: ") + cudaGetErrorString(guard_.err))

// One call in flight per handle (scratch buffers, counter block, events and helper streams are per model): a call on
// another stream than the previous one -- or a host-memory call right after a device-memory call that has not finished
// -- is ordered behind it ON THE DEVICE (no host synchronisation): the caller's stream waits for the event that closed
// the previous call; the helper streams are forked from the caller's stream after that.
static int serialize_after_previous_call(odl_model* m, cudaStream_t s) {
  if (m->timed) ODL_CUDA(cudaStreamWaitEvent(s, m->ev1, 0));
  return 0;
}

// cooperative kernels (n > 8): default lanes per system, as odl_kernels.cuh's ODL_G
// ODL_SOLVER_AUTO sweeps: attempted steps of the DOPRI5 pass, and the attempt at which a system whose progress projects
// beyond that cap leaves for the stiff pass (odl_sweep has the measurements)
static const int ODL_AUTO_CAP_DEFAULT = 704, ODL_AUTO_EARLY_DEFAULT = 384;

// Lanes per system of the cooperative kernels: the fewest that keep a lane's slice at <= 9 components (10 slice-sized
// vectors of DOPRI5 in registers).  Every lane of a group evaluates the whole right-hand side, so fewer lanes mean less
// redundant arithmetic -- measured on B200: 12 states, 2 lanes against 4: sweep 30.8 against 23.9 M solves/s, 8192 chains
// 17.6 against 16.3 M chain-steps/s; 35 states, 4 lanes against 8: sweep 8.3 against 5.9 M solves/s, 8192 chains 7.0
// against 4.7 M chain-steps/s (at 255 registers with 88 / 198 bytes of spill -- 8 lanes: none)
static int coop_lanes_default(int n_state) {
  return n_state <= 18 ? 2 : (n_state <= 36 ? 4 : (n_state <= 72 ? 8 : (n_state <= 144 ? 16 : 32)));
}

static uint64_t fnv1a(const void* data, size_t n, uint64_t h = 1469598103934665603ull) {
  const unsigned char* p = static_cast<const unsigned char*>(data);
  for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
  return h;
}

static double now_s() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

// NVRTC -> sm_100a cubin of one unit.  Pure function of its arguments (runs on worker threads; NVRTC programs are
// independent objects): no access to the model, no thread-local error state.
static Compiled compile_unit(std::string src, std::vector<std::string> opt, int unit, std::string cache_dir) {
  Compiled out;
  const double t0 = now_s();
  opt.push_back("-DODL_UNIT=" + std::to_string(unit));
  int major = 0, minor = 0;
  nvrtcVersion(&major, &minor);
  std::string keysrc = src + kKernelSrc + kAbiHeaderSrc;
  for (auto& o : opt) keysrc += o;
  keysrc += "nvrtc" + std::to_string(major) + "." + std::to_string(minor);
  char name[96];
  snprintf(name, sizeof name, "odl_%016llx_%s.cubin", (unsigned long long)fnv1a(keysrc.data(), keysrc.size()), kUnitName[unit]);
  std::string path;
  if (!cache_dir.empty()) {
    path = cache_dir + "/" + name;
    if (FILE* f = fopen(path.c_str(), "rb")) {
      fseek(f, 0, SEEK_END);
      long sz = ftell(f);
      fseek(f, 0, SEEK_SET);
      out.cubin.resize(sz > 0 ? sz : 0);
      size_t got = sz > 0 ? fread(out.cubin.data(), 1, sz, f) : 0;
      fclose(f);
      if (sz > 0 && got == (size_t)sz) {
        out.log = std::string(kUnitName[unit]) + ": cubin cache hit: " + path;
        out.cache_hit = true;
        out.seconds = now_s() - t0;
        return out;
      }
      out.cubin.clear();
    }
  }
  nvrtcProgram prog;
  const char* hdr_src[2] = {kKernelSrc, kAbiHeaderSrc};
  const char* hdr_name[2] = {"odl_kernels.cuh", "odl_abi.h"};
  std::string full = src + "\n#include \"odl_kernels.cuh\"\n";
  if (nvrtcCreateProgram(&prog, full.c_str(), "odl_model.cu", 2, hdr_src, hdr_name) != NVRTC_SUCCESS) {
    out.rc = ODL_ECOMPILE; out.err = "nvrtcCreateProgram failed";
    return out;
  }
  std::vector<const char*> copt;
  for (auto& o : opt) copt.push_back(o.c_str());
  nvrtcResult r = nvrtcCompileProgram(prog, (int)copt.size(), copt.data());
  size_t logsz = 0;
  nvrtcGetProgramLogSize(prog, &logsz);
  std::string log(logsz, '\0');
  if (logsz) nvrtcGetProgramLog(prog, &log[0]);
  while (!log.empty() && (log.back() == '\0' || log.back() == '\n')) log.pop_back();
  if (r != NVRTC_SUCCESS) {
    nvrtcDestroyProgram(&prog);
    out.rc = ODL_ECOMPILE;
    out.err = std::string("NVRTC (") + kUnitName[unit] + "): " + nvrtcGetErrorString(r) + "\n" + log;
    return out;
  }
  size_t sz = 0;
  if (nvrtcGetCUBINSize(prog, &sz) != NVRTC_SUCCESS || sz == 0) {
    nvrtcDestroyProgram(&prog);
    out.rc = ODL_ECOMPILE; out.err = "NVRTC produced no cubin";
    return out;
  }
  out.cubin.resize(sz);
  nvrtcGetCUBIN(prog, out.cubin.data());
  nvrtcDestroyProgram(&prog);
  if (!path.empty()) {
    std::string tmp = path + ".tmp." + std::to_string((long long)getpid()) + "." + std::to_string(unit);   // one writer per file
    if (FILE* f = fopen(tmp.c_str(), "wb")) {
      fwrite(out.cubin.data(), 1, sz, f);
      fclose(f);
      rename(tmp.c_str(), path.c_str());
    }
  }
  out.seconds = now_s() - t0;
  char head[96];
  snprintf(head, sizeof head, "%s: compiled in %.2f s", kUnitName[unit], out.seconds);
  out.log = head + (log.empty() ? std::string() : "\n" + log);
  return out;
}

static bool unit_applies(const odl_model* m, int unit) {
  if (unit == U_SWEEP_COOP || unit == U_MCMC_COOP) return m->coop_model();
  return unit >= 1 && unit < U_COUNT;
}
// start compiling a unit on a worker thread (no-op when it is compiled, compiling or loaded already)
static void unit_start(odl_model* m, int unit) {
  Unit& u = m->units[unit];
  if (u.have || u.job.valid() || !unit_applies(m, unit)) return;
  u.job = std::async(std::launch::async, compile_unit, m->src, m->opt, unit, m->cache_dir);
}
// wait for / run the compile of a unit
static int unit_compiled(odl_model* m, int unit) {
  Unit& u = m->units[unit];
  if (u.have) return 0;
  if (!unit_applies(m, unit)) return fail(ODL_EINVAL, std::string("kernel unit '") + kUnitName[unit] + "' does not exist for this model");
  if (!u.job.valid()) unit_start(m, unit);
  u.c = u.job.get();
  if (u.c.rc) return fail(u.c.rc, u.c.err);
  u.have = true;
  m->log += (m->log.empty() ? "" : "\n") + u.c.log;
  return 0;
}
static std::string cu_err(CUresult r);
static int load_driver();
struct KernelSlot { const char* name; CUfunction odl_model::*fn; int unit; };
static const KernelSlot kKernels[] = {
    {"odl_sweep_kernel", &odl_model::k_sweep, U_SWEEP}, {"odl_traj_kernel", &odl_model::k_traj, U_TRAJ},
    {"odl_mcmc_kernel", &odl_model::k_mcmc, U_MCMC}, {"odl_sweep_ros23_kernel", &odl_model::k_sweep_ros, U_SWEEP_ROS},
    {"odl_mcmc_ros23_kernel", &odl_model::k_mcmc_ros, U_MCMC_ROS}, {"odl_mcmc_auto_kernel", &odl_model::k_mcmc_auto, U_MCMC_AUTO},
    {"odl_sweep_radau5_kernel", &odl_model::k_sweep_radau, U_SWEEP_RADAU},
    {"odl_mcmc_radau5_kernel", &odl_model::k_mcmc_radau, U_MCMC_RADAU}, {"odl_sweep_bdf_kernel", &odl_model::k_sweep_bdf, U_SWEEP_BDF},
    {"odl_mcmc_bdf_kernel", &odl_model::k_mcmc_bdf, U_MCMC_BDF}, {"odl_order_key_kernel", &odl_model::k_order_key, U_ORDER},
    {"odl_order_scan_kernel", &odl_model::k_order_scan, U_ORDER}, {"odl_order_scatter_kernel", &odl_model::k_order_scatter, U_ORDER},
    {"odl_feed_done_kernel", &odl_model::k_feed_done, U_ORDER}, {"odl_gate_kernel", &odl_model::k_gate, U_ORDER},
    {"odl_sweep_coop_kernel", &odl_model::k_sweep_coop, U_SWEEP_COOP}, {"odl_mcmc_coop_kernel", &odl_model::k_mcmc_coop, U_MCMC_COOP}};

// compiled + loaded on the model's device + its CUfunctions fetched
static int unit_ready(odl_model* m, int unit) {
  Unit& u = m->units[unit];
  if (u.mod) return 0;
  if (!m->on_gpu) return fail(ODL_ENODEVICE, "model was created compile_only / without a GPU");
  int rc = unit_compiled(m, unit);
  if (rc) return rc;
  CUresult r = g_drv.ModuleLoadData(&u.mod, u.c.cubin.data());
  if (r != CUDA_SUCCESS) { u.mod = nullptr; return fail(ODL_ECUDA, "cuModuleLoadData: " + cu_err(r)); }
  for (const KernelSlot& k : kKernels) {
    if (k.unit != unit) continue;
    CUfunction f = nullptr;
    r = g_drv.ModuleGetFunction(&f, u.mod, k.name);
    if (r != CUDA_SUCCESS) return fail(ODL_ECUDA, std::string("cuModuleGetFunction(") + k.name + "): " + cu_err(r));
    m->*(k.fn) = f;
  }
  return 0;
}
static int units_ready(odl_model* m, std::initializer_list<int> units) {
  for (int u : units) unit_start(m, u);          // all of them compile side by side ...
  for (int u : units) { int rc = unit_ready(m, u); if (rc) return rc; }
  return 0;
}

extern "C" int odl_abi_version(void) { return ODL_ABI_VERSION; }
extern "C" const char* odl_last_error(void) { return g_err.c_str(); }
extern "C" long long odl_launch_count(void) { return g_launches.load(); }

extern "C" int odl_model_create(const char* model_cuda_src, int n_state, int n_param, int n_out,
                                const odl_build_opts* opts, odl_model** out) {
  if (!model_cuda_src || !out || n_state < 1 || n_param < 0 || n_out < 1)
    return fail(ODL_EINVAL, "odl_model_create: bad arguments");
  if (n_param > ODL_MAX_WALK) return fail(ODL_EINVAL, "odl_model_create: more than 64 parameters are not supported");
  *out = nullptr;
  odl_model* m = new odl_model();
  m->n_state = n_state; m->n_param = n_param; m->n_out = n_out;
  int device = -1;
  int compile_only = 0;
  if (opts) {
    device = opts->device;
    if (opts->block_threads > 0) m->block = opts->block_threads;
    if (opts->min_blocks > 0) m->minblocks = opts->min_blocks;
    m->dense = opts->dense_output ? 1 : 0;
    m->y0p = opts->y0_from_param ? 1 : 0;
    m->coop = opts->coop_lanes;
    compile_only = opts->compile_only;
    if (opts->cache_dir) m->cache_dir = opts->cache_dir;
  }
  if (m->block % 32 || m->block > 1024) { delete m; return fail(ODL_EINVAL, "block_threads must be a multiple of 32, <= 1024"); }
  if (m->coop == 0) m->coop = coop_lanes_default(n_state);
  if (m->coop != 2 && m->coop != 4 && m->coop != 8 && m->coop != 16 && m->coop != 32) { delete m; return fail(ODL_EINVAL, "coop_lanes must be 2, 4, 8, 16 or 32"); }
  m->src = model_cuda_src;
  m->opt = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "-default-device", "--ptxas-options=-v",
            "-DODL_BLOCK=" + std::to_string(m->block), "-DODL_MINBLOCKS=" + std::to_string(m->minblocks),
            "-DODL_DENSE=" + std::to_string(m->dense), "-DODL_Y0P=" + std::to_string(m->y0p)};
  if (m->n_state > 8) m->opt.push_back("-DODL_G=" + std::to_string(m->coop));
  // tuning hook (development): extra -D options for the kernel source, e.g. ODL_KERNEL_DEFINES="-DODL_INNER=8"
  if (const char* extra = getenv("ODL_KERNEL_DEFINES")) {
    std::string e(extra);
    size_t pos = 0;
    while (pos < e.size()) {
      size_t sp = e.find(' ', pos);
      if (sp == std::string::npos) sp = e.size();
      if (sp > pos) m->opt.push_back(e.substr(pos, sp - pos));
      pos = sp + 1;
    }
  }
  // The kernels the default paths use start compiling now, side by side on worker threads: forward sweep
  // (ordering + DOPRI5 + the BDF stiff pass), trajectories, chains.  The other steppers (ROS23, Radau5, the per-proposal
  // DOPRI5->ROS23 kernel, the BDF chain kernel) compile when a call first asks for them.  compile_only = 1: every unit,
  // waited for (fills the cubin cache, reports any compile error); compile_only = 2: the default set only.
  const bool coop = m->coop_model();
  std::vector<int> first = {U_ORDER, U_SWEEP_BDF, U_TRAJ, coop ? U_SWEEP_COOP : U_SWEEP, coop ? U_MCMC_COOP : U_MCMC};
  if (compile_only == 1) {
    first.clear();
    for (int u = 1; u < U_COUNT; ++u) if (unit_applies(m, u)) first.push_back(u);
  }
  if (compile_only) {
    // bounded parallelism on the build box: as many programs at once as there are cores
    size_t par = std::max(1u, std::thread::hardware_concurrency());
    for (size_t i = 0; i < first.size(); i += par) {
      for (size_t j = i; j < std::min(first.size(), i + par); ++j) unit_start(m, first[j]);
      for (size_t j = i; j < std::min(first.size(), i + par); ++j) {
        int rc = unit_compiled(m, first[j]);
        if (rc) { odl_model_destroy(m); return rc; }
      }
    }
    *out = m;
    return 0;
  }
  for (int u : first) unit_start(m, u);
  // ---- GPU side ----
  auto bail = [&](int code) { odl_model_destroy(m); return code; };
  int prev_device = -1;
  if (cudaGetDevice(&prev_device) != cudaSuccess) { cudaGetLastError(); prev_device = -1; }
  if (device >= 0) { cudaError_t e = cudaSetDevice(device); if (e != cudaSuccess) return bail(fail(ODL_ENODEVICE, std::string("cudaSetDevice: ") + cudaGetErrorString(e))); }
  { cudaError_t e = cudaFree(0); if (e != cudaSuccess) return bail(fail(ODL_ENODEVICE, std::string("no usable CUDA device: ") + cudaGetErrorString(e))); }
  if (cudaGetDevice(&m->device) != cudaSuccess) return bail(fail(ODL_ENODEVICE, "cudaGetDevice failed"));
  struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev_device};
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, m->device) != cudaSuccess) return bail(fail(ODL_ECUDA, "cudaGetDeviceProperties failed"));
  if (prop.major != 10) return bail(fail(ODL_ENODEVICE, std::string("device '") + prop.name + "' is not sm_100 (Blackwell B200); this library has no other code path"));
  m->sm_count = prop.multiProcessorCount;
  int rc;
  if ((rc = load_driver())) return bail(rc);
  if (cudaEventCreate(&m->ev0) != cudaSuccess || cudaEventCreate(&m->ev1) != cudaSuccess ||
      cudaEventCreate(&m->evp[0]) != cudaSuccess || cudaEventCreate(&m->evp[1]) != cudaSuccess ||
      cudaEventCreateWithFlags(&m->ev_aux, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&m->ev_chunk[0], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&m->ev_chunk[1], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&m->ev_chunk[2], cudaEventDisableTiming) != cudaSuccess ||
      cudaStreamCreateWithFlags(&m->aux, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&m->aux2, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&m->piece_stream[0], cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&m->piece_stream[1], cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&m->ev_piece[0], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&m->ev_piece[1], cudaEventDisableTiming) != cudaSuccess)
    return bail(fail(ODL_ECUDA, "cudaEventCreate / cudaStreamCreate failed"));
  if ((rc = m->counter.ensure(8192))) return bail(rc);
  m->on_gpu = true;
  *out = m;
  return 0;
}

extern "C" int odl_model_destroy(odl_model* m) {
  if (!m) return 0;
  for (Unit& u : m->units) if (u.job.valid()) u.job.wait();          // worker threads hold copies of their inputs only
  int prev = -1;
  if (m->on_gpu) {
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
    cudaSetDevice(m->device);
    for (Unit& u : m->units) if (u.mod && g_drv.ModuleUnload) g_drv.ModuleUnload(u.mod);
  }
  odl_comm_destroy(m);
  m->data.buf.release(); m->grid.buf.release(); m->counter.release(); m->mt_state.release(); m->handover.release();
  for (auto& s : m->scratch) s.release();
  if (m->ev0) cudaEventDestroy(m->ev0);
  if (m->ev1) cudaEventDestroy(m->ev1);
  for (auto& e : m->evp) if (e) cudaEventDestroy(e);
  if (m->ev_aux) cudaEventDestroy(m->ev_aux);
  if (m->ev_fork) cudaEventDestroy(m->ev_fork);
  for (auto& e : m->ev_chunk) if (e) cudaEventDestroy(e);
  if (m->aux) cudaStreamDestroy(m->aux);
  if (m->aux2) cudaStreamDestroy(m->aux2);
  for (auto& st_ : m->piece_stream) if (st_) cudaStreamDestroy(st_);
  for (auto& e : m->ev_piece) if (e) cudaEventDestroy(e);
  if (m->on_gpu && prev >= 0 && prev != m->device) cudaSetDevice(prev);
  delete m;
  return 0;
}

extern "C" const char* odl_model_build_log(const odl_model* m) { return m ? m->log.c_str() : ""; }

// seconds NVRTC spent on a unit (0 = cache hit), -1 = not compiled (yet); the unit is NOT compiled by asking
extern "C" int odl_model_unit_seconds(const odl_model* m, const char* unit, double* seconds, int* cache_hit) {
  if (!m || !unit || !seconds) return fail(ODL_EINVAL, "odl_model_unit_seconds: null argument");
  for (int u = 1; u < U_COUNT; ++u)
    if (!strcmp(unit, kUnitName[u])) {
      *seconds = m->units[u].have ? m->units[u].c.seconds : -1.0;
      if (cache_hit) *cache_hit = m->units[u].have && m->units[u].c.cache_hit;
      return 0;
    }
  return fail(ODL_EINVAL, "odl_model_unit_seconds: unknown unit");
}

// the unit's kernel, compiled and loaded on demand
static int kernel_by_name(odl_model* m, const char* k, CUfunction* f) {
  struct { const char* name; CUfunction odl_model::*fn; int unit; } tab[] = {
      {"sweep", &odl_model::k_sweep, U_SWEEP}, {"mcmc", &odl_model::k_mcmc, U_MCMC}, {"traj", &odl_model::k_traj, U_TRAJ},
      {"sweep_ros23", &odl_model::k_sweep_ros, U_SWEEP_ROS}, {"mcmc_ros23", &odl_model::k_mcmc_ros, U_MCMC_ROS},
      {"mcmc_auto", &odl_model::k_mcmc_auto, U_MCMC_AUTO}, {"sweep_radau5", &odl_model::k_sweep_radau, U_SWEEP_RADAU},
      {"mcmc_radau5", &odl_model::k_mcmc_radau, U_MCMC_RADAU}, {"sweep_coop", &odl_model::k_sweep_coop, U_SWEEP_COOP},
      {"mcmc_coop", &odl_model::k_mcmc_coop, U_MCMC_COOP}, {"sweep_bdf", &odl_model::k_sweep_bdf, U_SWEEP_BDF},
      {"mcmc_bdf", &odl_model::k_mcmc_bdf, U_MCMC_BDF}, {"order_key", &odl_model::k_order_key, U_ORDER},
      {"order_scatter", &odl_model::k_order_scatter, U_ORDER}};
  if (!k) return fail(ODL_EINVAL, "null kernel name");
  for (auto& t : tab)
    if (!strcmp(k, t.name)) {
      int rc = unit_ready(m, t.unit);
      if (rc) return rc;
      *f = m->*(t.fn);
      return 0;
    }
  return fail(ODL_EINVAL, "unknown kernel");
}

static size_t smem_bytes(const OdlData& d, int block);
// largest CTA (<= preferred) whose tables + per-thread staging leave room for at least two CTAs per SM
static unsigned pick_block(const OdlData& d, int preferred) {
  int b = preferred;
  while (b > 32 && smem_bytes(d, b) > 100 * 1024) b /= 2;
  return (unsigned)b;
}
// cooperative kernels (n > 8): lanes per system as in odl_kernels.cuh (ODL_G), CTA size, shared memory
static const unsigned kCoopBlock = 128;
static size_t coop_smem_bytes(const odl_model* m, const OdlData& d) {
  const size_t groups = kCoopBlock / m->coop;
  size_t doubles = 2 * ODL_LOGTAB + (size_t)d.n_slot + 3 * (size_t)d.n_obs + ((size_t)d.n_obs + 1) / 2 +
                   groups * (size_t)ODL_COOP_ROW(m->n_state + m->n_param + d.stage_stride, m->coop);
  return doubles * sizeof(double);
}
static size_t smem_bytes(const OdlData& d, int block) {
  size_t doubles = 2 * ODL_LOGTAB + (size_t)d.n_slot + 3 * (size_t)d.n_obs + ((size_t)d.n_obs + 1) / 2 + (size_t)block * d.stage_stride;
  return doubles * sizeof(double);
}

extern "C" int odl_model_kernel_info(odl_model* m, const char* kernel, int* regs, int* local_bytes,
                                     int* max_blocks_per_sm) {
  if (!m || !m->on_gpu) return fail(ODL_ENODEVICE, "odl_model_kernel_info: model is not loaded on a GPU");
  ODL_ON_DEVICE(m);
  CUfunction f = nullptr;
  int rc = kernel_by_name(m, kernel, &f);
  if (rc) return rc;
  int r = 0, l = 0, b = 0;
  ODL_CU(g_drv.FuncGetAttribute(&r, CU_FUNC_ATTRIBUTE_NUM_REGS, f));
  ODL_CU(g_drv.FuncGetAttribute(&l, CU_FUNC_ATTRIBUTE_LOCAL_SIZE_BYTES, f));
  const bool warp_cta = f == m->k_sweep_bdf || f == m->k_mcmc_bdf || f == m->k_sweep_radau || f == m->k_mcmc_radau;
  const bool coop = f == m->k_sweep_coop || f == m->k_mcmc_coop;
  const int block = coop ? (int)kCoopBlock : (warp_cta ? 32 : m->block);
  size_t sm = !m->data.set ? 0 : (coop ? coop_smem_bytes(m, m->data.d) : smem_bytes(m->data.d, block));
  g_drv.FuncSetAttribute(f, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)std::max<size_t>(sm, 1024));
  ODL_CU(g_drv.OccupancyMaxActiveBlocksPerMultiprocessor(&b, f, block, sm));
  if (regs) *regs = r;
  if (local_bytes) *local_bytes = l;
  if (max_blocks_per_sm) *max_blocks_per_sm = b;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// data tables
// ---------------------------------------------------------------------------------------------
static int upload_tables(odl_model* m, Tables& T, int n_slot, const double* slot_time, int n_obs, const int* obs_slot,
                         const int* obs_col, const double* ln_obs, const double* log_sigma, double sstot,
                         const double* y0, const int* y0_from_param, double t0) {
  if (!m->on_gpu) return fail(ODL_ENODEVICE, "model was created compile_only / without a GPU");
  if (n_slot < 1 || !slot_time) return fail(ODL_EINVAL, "need at least one output time");
  for (int i = 1; i < n_slot; ++i)
    if (!(slot_time[i] > slot_time[i - 1])) return fail(ODL_EINVAL, "output times must be strictly ascending");
  if (!(slot_time[0] >= t0)) return fail(ODL_EINVAL, "first output time precedes t0");
  ODL_ON_DEVICE(m);
  const int N = m->n_state;
  // layout: slot_t[K] lnO[n_obs] w[n_obs] lin[n_obs] y0[N] | src[n_obs] y0p[N]
  size_t nd = (size_t)n_slot + 3 * (size_t)n_obs + N;
  size_t ni = (size_t)n_obs + N;
  std::vector<double> hd(nd);
  std::vector<int> hi(ni);
  double* slot = hd.data();
  double* lnO = slot + n_slot;
  double* den = lnO + n_obs;
  double* lin = den + n_obs;
  double* hy0 = lin + n_obs;
  int* src = hi.data();
  int* y0p = src + n_obs;
  std::copy(slot_time, slot_time + n_slot, slot);
  for (int o = 0; o < n_obs; ++o) {
    if (obs_slot[o] < 0 || obs_slot[o] >= n_slot || obs_col[o] < 0 || obs_col[o] >= m->n_out)
      return fail(ODL_EINVAL, "observation row refers to a slot/column out of range");
    lnO[o] = ln_obs[o];
    const double s2 = log_sigma[o] * log_sigma[o];     // S**2
    den[o] = 1.0 / (2.0 * s2);                         // 1 / (2*(S**2))   (stats.py:41); sigma = 0 -> inf -> masked term
    lin[o] = std::exp(ln_obs[o]);                      // Framework.py:700
    src[o] = obs_slot[o] * m->n_out + obs_col[o];
  }
  for (int i = 0; i < N; ++i) {
    hy0[i] = y0 ? y0[i] : 0.0;
    y0p[i] = y0_from_param ? y0_from_param[i] : -1;
    if (y0p[i] >= m->n_param) return fail(ODL_EINVAL, "y0_from_param index out of range");
    if (y0p[i] >= 0 && !m->y0p)
      return fail(ODL_EINVAL, "model was built with y0_from_param = 0 but a state takes its initial value from a parameter");
  }
  int rc = T.buf.ensure(nd * sizeof(double) + ni * sizeof(int));
  if (rc) return rc;
  char* base = static_cast<char*>(T.buf.p);
  ODL_CUDA(cudaMemcpy(base, hd.data(), nd * sizeof(double), cudaMemcpyHostToDevice));
  ODL_CUDA(cudaMemcpy(base + nd * sizeof(double), hi.data(), ni * sizeof(int), cudaMemcpyHostToDevice));
  double* dd = reinterpret_cast<double*>(base);
  int* di = reinterpret_cast<int*>(base + nd * sizeof(double));
  OdlData& d = T.d;
  d.slot_t = dd; d.obs_lnO = dd + n_slot; d.obs_w = d.obs_lnO + n_obs; d.obs_lin = d.obs_w + n_obs;
  d.y0 = d.obs_lin + n_obs;
  d.obs_src = di; d.y0_from_param = di + n_obs;
  d.n_slot = n_slot; d.n_obs = n_obs;
  d.stage_stride = (n_slot * m->n_out) | 1;
  d.pad_ = 0; d.t0 = t0; d.sstot = sstot; d.inv_sstot = 1.0 / sstot;
  T.set = true;
  return 0;
}

extern "C" int odl_model_set_data(odl_model* m, int n_slot, const double* slot_time, int n_obs, const int* obs_slot,
                                  const int* obs_col, const double* ln_obs, const double* log_sigma, double sstot,
                                  const double* y0, const int* y0_from_param, double t0) {
  if (!m) return fail(ODL_EINVAL, "null model");
  if (n_obs < 1 || !obs_slot || !obs_col || !ln_obs || !log_sigma) return fail(ODL_EINVAL, "need observation rows");
  return upload_tables(m, m->data, n_slot, slot_time, n_obs, obs_slot, obs_col, ln_obs, log_sigma, sstot, y0,
                       y0_from_param, t0);
}

extern "C" int odl_model_set_grid(odl_model* m, int n_t, const double* times, const double* y0,
                                  const int* y0_from_param) {
  if (!m) return fail(ODL_EINVAL, "null model");
  return upload_tables(m, m->grid, n_t, times, 0, nullptr, nullptr, nullptr, nullptr, 1.0, y0, y0_from_param,
                       times ? times[0] : 0.0);
}

// ---------------------------------------------------------------------------------------------
// launches
// ---------------------------------------------------------------------------------------------
static void fill_opts(OdlOpts& o, const odl_solver_opts* so) {
  o.rtol = so && so->rtol > 0 ? so->rtol : 1.49012e-8;
  o.atol = so && so->atol > 0 ? so->atol : 1.49012e-8;
  o.h0 = so ? so->h0 : 0.0;
  o.hmax = so ? so->hmax : 0.0;
  o.max_steps = so && so->max_steps > 0 ? so->max_steps : 500000;
  o.stiff_check = so ? so->stiff_check : 0;
  o.stiff_min_steps = so && so->stiff_min_steps > 0 ? so->stiff_min_steps : 2000;
  o.early_check_steps = so && so->early_check_steps > 0 ? so->early_check_steps : 0;   // AUTO sets its own default
  o.lanes = 0;
  o.explicit_cap = 0; o.pad_ = 0;
  o.watchdog_spins = 75000000;   // ~30 s without a single entry and without the producer finishing
  if (const char* w = getenv("ODL_WATCHDOG_SPINS")) o.watchdog_spins = std::max(1000, atoi(w));
}

static int launch(odl_model* m, CUfunction f, unsigned grid, unsigned block, size_t smem, cudaStream_t s, void** params) {
  if (!f) return fail(ODL_ECUDA, "internal: launch of a kernel whose unit is not loaded");
  if (smem > 48 * 1024) ODL_CU(g_drv.FuncSetAttribute(f, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem));
  CUresult r = g_drv.LaunchKernel(f, grid, 1, 1, block, 1, 1, (unsigned)smem, (CUstream)s, params, nullptr);
  if (r != CUDA_SUCCESS) {
    const char* name = "?";
    for (const KernelSlot& k : kKernels) if (m->*(k.fn) == f) name = k.name;
    char buf[256];
    snprintf(buf, sizeof buf, "cuLaunchKernel(%s, grid %u, block %u, smem %zu): ", name, grid, block, smem);
    return fail(ODL_ECUDA, buf + cu_err(r));
  }
  g_launches.fetch_add(1);
  return 0;
}

// launch with thread-block clusters of `cluster` CTAs: the CTAs of a cluster are placed on SMs of one GPC next to each
// other (2 = the two SMs of a TPC).  The kernels do not use cluster features; this is about WHERE the CTAs land.
static int launch_clustered(odl_model* m, CUfunction f, unsigned grid, unsigned block, size_t smem, cudaStream_t s, void** params,
                            unsigned cluster) {
  if (cluster <= 1) return launch(m, f, grid, block, smem, s, params);
  if (!f) return fail(ODL_ECUDA, "internal: launch of a kernel whose unit is not loaded");
  if (smem > 48 * 1024) ODL_CU(g_drv.FuncSetAttribute(f, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem));
  CUlaunchAttribute attr{};
  attr.id = CU_LAUNCH_ATTRIBUTE_CLUSTER_DIMENSION;
  attr.value.clusterDim.x = cluster; attr.value.clusterDim.y = 1; attr.value.clusterDim.z = 1;
  CUlaunchConfig cfg{};
  cfg.gridDimX = grid; cfg.gridDimY = 1; cfg.gridDimZ = 1;
  cfg.blockDimX = block; cfg.blockDimY = 1; cfg.blockDimZ = 1;
  cfg.sharedMemBytes = (unsigned)smem; cfg.hStream = (CUstream)s; cfg.attrs = &attr; cfg.numAttrs = 1;
  CUresult r = g_drv.LaunchKernelEx(&cfg, f, params, nullptr);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof buf, "cuLaunchKernelEx(grid %u, block %u, smem %zu, cluster %u): ", grid, block, smem, cluster);
    return fail(ODL_ECUDA, buf + cu_err(r));
  }
  g_launches.fetch_add(1);
  return 0;
}

// stage a host array on the device (in) or reserve room for a result (out)
struct Staging {
  odl_model* m; cudaStream_t s; int next = 0; int mem;
  struct Out { void* host; void* dev; size_t bytes; };
  std::vector<Out> outs;
  template <class T> int in(const T* host, size_t count, const T** dev) {
    if (!host) { *dev = nullptr; return 0; }
    if (mem == ODL_MEM_DEVICE) { *dev = host; return 0; }
    if (next >= 20) return fail(ODL_EINVAL, "internal: staging slots exhausted");
    DevBuf& b = m->scratch[next++];
    int rc = b.ensure(count * sizeof(T)); if (rc) return rc;
    ODL_CUDA(cudaMemcpyAsync(b.p, host, count * sizeof(T), cudaMemcpyHostToDevice, s));
    *dev = static_cast<const T*>(b.p);
    return 0;
  }
  template <class T> int inout(T* host, size_t count, T** dev, bool copy_in) {
    if (!host) { *dev = nullptr; return 0; }
    if (mem == ODL_MEM_DEVICE) { *dev = host; return 0; }
    if (next >= 20) return fail(ODL_EINVAL, "internal: staging slots exhausted");
    DevBuf& b = m->scratch[next++];
    int rc = b.ensure(count * sizeof(T)); if (rc) return rc;
    if (copy_in) ODL_CUDA(cudaMemcpyAsync(b.p, host, count * sizeof(T), cudaMemcpyHostToDevice, s));
    *dev = static_cast<T*>(b.p);
    outs.push_back({host, b.p, count * sizeof(T)});
    return 0;
  }
  int finish() {
    if (mem == ODL_MEM_DEVICE) return 0;
    for (auto& o : outs) ODL_CUDA(cudaMemcpyAsync(o.host, o.dev, o.bytes, cudaMemcpyDeviceToHost, s));
    ODL_CUDA(cudaStreamSynchronize(s));
    return 0;
  }
};

extern "C" int odl_sweep(odl_model* m, const odl_solver_opts* so, long long n, const double* theta, int mem,
                         double* chi, double* r2, int* status, int* nsteps, double* pred_or_null, void* stream) {
  if (!m || !m->on_gpu) return fail(ODL_ENODEVICE, "odl_sweep: model is not loaded on a GPU (no CPU fallback exists)");
  if (!m->data.set) return fail(ODL_EINVAL, "odl_sweep: call odl_model_set_data first");
  // r2, status, nsteps are optional (NULL = not wanted: neither written nor copied back); the reference's own batch
  // seam returns chi alone (_Fit_worker, Framework.py:41-48), a failed solve shows as chi = NaN either way
  if (n < 0 || (n > 0 && (!theta || !chi))) return fail(ODL_EINVAL, "odl_sweep: null buffer");
  if (n == 0) return 0;
  const int solver = so ? so->solver : ODL_SOLVER_DOPRI5;
  if (solver < 0 || solver > ODL_SOLVER_BDF) return fail(ODL_EINVAL, "odl_sweep: unknown solver");
  const int tail_solver = so && so->tail_solver > 0 ? so->tail_solver : ODL_SOLVER_BDF;
  if (tail_solver != ODL_SOLVER_BDF && tail_solver != ODL_SOLVER_RADAU5)
    return fail(ODL_EINVAL, "odl_sweep: tail_solver must be ODL_SOLVER_BDF or ODL_SOLVER_RADAU5");
  ODL_ON_DEVICE(m);
  int rc;
  {
    // the kernels this call launches (compiled side by side if they are not yet)
    const bool coop = m->coop_model();
    const int tail_unit = tail_solver == ODL_SOLVER_RADAU5 ? U_SWEEP_RADAU : U_SWEEP_BDF;
    const bool coop_plain = coop && !(so && (so->stiff_check || so->early_check_steps > 0));
    if (solver == ODL_SOLVER_AUTO) rc = coop ? units_ready(m, {U_ORDER, U_SWEEP_COOP, tail_unit}) : units_ready(m, {U_ORDER, U_SWEEP, tail_unit});
    else if (solver == ODL_SOLVER_DOPRI5) rc = units_ready(m, {coop_plain ? U_SWEEP_COOP : U_SWEEP});
    else rc = units_ready(m, {solver == ODL_SOLVER_ROS23 ? U_SWEEP_ROS : (solver == ODL_SOLVER_RADAU5 ? U_SWEEP_RADAU : U_SWEEP_BDF)});
    if (rc) return rc;
  }
  if (solver == ODL_SOLVER_AUTO && n > 2147483647LL)
    return fail(ODL_EINVAL, "odl_sweep: ODL_SOLVER_AUTO orders rows through 32-bit indices; split tables beyond 2^31-1 rows");
  cudaStream_t s = (cudaStream_t)stream;
  if ((rc = serialize_after_previous_call(m, s))) return rc;
  Staging st{m, s, 0, mem};
  OdlSweepArgs A{};
  // Host-memory ODL_SOLVER_AUTO sweep of a large table: theta travels in two pieces on the helper stream and the second
  // piece arrives while the first is being ordered and integrated (each piece is ordered and swept on its own; the
  // stiff pass runs once over what both leave).  Rows are independent, so the pieces change nothing in the results.
  const int auto_flags = so ? so->auto_flags : 0;
  const bool chunked = mem == ODL_MEM_HOST && solver == ODL_SOLVER_AUTO && n >= (1 << 18) && !m->coop_model() &&
                       !(auto_flags & (ODL_AUTO_UNORDERED | ODL_AUTO_ONE_PIECE));
  if (chunked) {
    DevBuf& bt = m->scratch[st.next++];
    if ((rc = bt.ensure((size_t)n * m->n_param * sizeof(double)))) return rc;
    A.theta = static_cast<const double*>(bt.p);
  } else if ((rc = st.in(theta, (size_t)n * m->n_param, &A.theta))) return rc;
  if ((rc = st.inout(chi, (size_t)n, &A.chi, false))) return rc;
  if ((rc = st.inout(r2, (size_t)n, &A.r2, false))) return rc;
  if ((rc = st.inout(status, (size_t)n, &A.status, false))) return rc;
  if ((rc = st.inout(nsteps, (size_t)n, &A.nsteps, false))) return rc;
  if ((rc = st.inout(pred_or_null, (size_t)n * m->data.d.n_obs, &A.pred, false))) return rc;
  A.n = n; A.index = nullptr; A.index_count = nullptr;
  // counter block (zeroed per call): [0] work counter pass 0, [64] count of list A, [128] work counter pass 1,
  // [192] count of list B, [256] work counter pass 2
  char* cb = static_cast<char*>(m->counter.p);
  auto ctr = [&](int off) { return reinterpret_cast<unsigned long long*>(cb + off); };
  auto cnt = [&](int off) { return reinterpret_cast<int*>(cb + off); };
  A.counter = ctr(0);
  OdlOpts O; fill_opts(O, so);
  ODL_CUDA(cudaMemsetAsync(m->counter.p, 0, 8192, s));
  OdlData D = m->data.d;
  auto go = [&](cudaStream_t sx, CUfunction f, const OdlOpts& Ox, const OdlSweepArgs& Ax, unsigned block, long long items) -> int {
    const size_t smem = smem_bytes(D, (int)block);
    if (smem > 48 * 1024) ODL_CU(g_drv.FuncSetAttribute(f, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem));
    int per_sm = 0;
    ODL_CU(g_drv.OccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, f, (int)block, smem));
    if (per_sm < 1) return fail(ODL_ECUDA, "sweep kernel does not fit on an SM (shared memory / registers)");
    long long want = (items + block - 1) / block;
    unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(want, (long long)per_sm * m->sm_count));
    OdlData Dl = D; OdlOpts Ol = Ox; OdlSweepArgs Al = Ax;
    void* params[] = {&Dl, &Ol, &Al};
    return launch(m, f, grid, block, smem, sx, params);
  };
  auto go_coop = [&](cudaStream_t sx, const OdlOpts& Ox, const OdlSweepArgs& Ax, long long items) -> int {
    const size_t smem = coop_smem_bytes(m, D);
    if (smem > 227 * 1024) return fail(ODL_ECUDA, "cooperative sweep kernel: tables + staging exceed shared memory");
    if (smem > 48 * 1024) ODL_CU(g_drv.FuncSetAttribute(m->k_sweep_coop, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem));
    int per_sm = 0;
    ODL_CU(g_drv.OccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, m->k_sweep_coop, (int)kCoopBlock, smem));
    if (per_sm < 1) return fail(ODL_ECUDA, "cooperative sweep kernel does not fit on an SM");
    const long long groups = kCoopBlock / m->coop;
    unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((items + groups - 1) / groups, (long long)per_sm * m->sm_count));
    OdlData Dl = D; OdlOpts Ol = Ox; OdlSweepArgs Al = Ax;
    void* params[] = {&Dl, &Ol, &Al};
    return launch(m, m->k_sweep_coop, grid, kCoopBlock, smem, sx, params);
  };
  ODL_CUDA(cudaEventRecord(m->ev0, s));
  m->n_pass = 1;
  if (solver != ODL_SOLVER_AUTO) {
    A.defer_list[0] = A.defer_list[1] = nullptr; A.defer_count[0] = A.defer_count[1] = nullptr;
    const bool warp_cta = solver == ODL_SOLVER_RADAU5 || solver == ODL_SOLVER_BDF || solver == ODL_SOLVER_AUTO;   // compiled for one warp per CTA
    CUfunction f1 = solver == ODL_SOLVER_ROS23 ? m->k_sweep_ros : (solver == ODL_SOLVER_RADAU5 ? m->k_sweep_radau :
                    (solver == ODL_SOLVER_BDF ? m->k_sweep_bdf : m->k_sweep));
    if (solver == ODL_SOLVER_DOPRI5 && m->coop_model() && !O.stiff_check && O.early_check_steps == 0) {
      if ((rc = go_coop(s, O, A, n))) return rc;               // n > 8: several lanes per system
    } else if ((rc = go(s, f1, O, A, warp_cta ? 32u : pick_block(D, m->block), n))) return rc;
  } else {
    // Cost-ordered bulk pass + the stiff pass (no host synchronisation; list lengths stay on the device):
    //   order    key = |J(t0,y0,theta)|_inf (t_end-t0) per system, quarter-octave bins, highest first -> index[]
    //            (Spearman 0.73 with the DOPRI5 step count on the demo priors; the 1 % longest systems all sit in
    //            the first tenth).  Long systems start first, so the launch does not end on a few stragglers, and
    //            the ones DOPRI5 cannot finish are found while most of the sweep is still ahead.
    //   bulk     DOPRI5, every system in that order, at most cap0 attempted steps (a check at 3/4 of them drops systems
    //            whose progress projects beyond the cap; Hairer's test drops the ones it calls stiff) -> feed list
    //   stiff    variable-order BDF (or Radau5) over the feed list.  It is bound by the LATENCY of its longest systems
    //            (~900 sequential steps of ~600 dependent instructions, 1.4 ms), not by throughput, so it runs BESIDE
    //            the bulk pass on SMs of its own: `tail_sms` CTAs of 8 warps, launched first on the helper stream,
    //            each asking for so much shared memory that no bulk CTA fits next to it (sharing SMs was measured in
    //            round 1: every BDF step slows down next to twelve DOPRI5 warps, 6.5 ms against 5.0 ms one after the
    //            other); the bulk grid covers the remaining SMs.  The consumer takes feed entries as they land; a
    //            one-thread kernel after the last bulk launch marks the feed complete.  A last launch of the stiff
    //            kernel over the feed list picks up whatever the consumer did not finish (normally nothing: entries
    //            it finished are marked) -- so the results never depend on the consumer having kept up, and a consumer
    //            that gave up (watchdog) costs time, not rows.  Small sweeps (and ODL_AUTO_SEQUENTIAL) run the stiff
    //            pass after the bulk pass on single-warp CTAs spread over every SM, as in round 1.
    // Step counts are heavy-tailed (two_i prior: median 56, mean 103, p99.9 3200, max > 1e5 for DOPRI5).
    DevBuf& bidx = m->scratch[st.next++];
    DevBuf& bfeed = m->scratch[st.next++];
    DevBuf& bbins = m->scratch[st.next++];
    if ((rc = bidx.ensure((size_t)n * sizeof(int)))) return rc;
    if ((rc = bfeed.ensure((size_t)n * sizeof(int)))) return rc;
    if ((rc = bbins.ensure((size_t)n))) return rc;
    int* index = static_cast<int*>(bidx.p);
    int* feed = static_cast<int*>(bfeed.p);
    const int flags = so ? so->auto_flags : 0;
    const bool ordered = !(flags & ODL_AUTO_UNORDERED);
    if (getenv("ODL_TIMELINE")) {                                // development: see OdlSweepArgs.timeline
      DevBuf& bt = m->scratch[20];
      if ((rc = bt.ensure((size_t)n * 3 * sizeof(long long)))) return rc;
      ODL_CUDA(cudaMemsetAsync(bt.p, 0, (size_t)n * 3 * sizeof(long long), s));
      A.timeline = static_cast<long long*>(bt.p);
    }
    const bool coop_bulk = m->coop_model();                      // n > 8: the bulk pass is the cooperative kernel (go_coop)
    const bool beside = !coop_bulk && !(flags & ODL_AUTO_SEQUENTIAL) && ((flags & ODL_AUTO_CONCURRENT) || n >= (1 << 18));
    const int cap0 = so && so->pass_cap0 > 0 ? so->pass_cap0 : ODL_AUTO_CAP_DEFAULT;
    CUfunction k_tail = tail_solver == ODL_SOLVER_RADAU5 ? m->k_sweep_radau : m->k_sweep_bdf;
    // counter block (zeroed above): [0] bulk work counter (+ [384], [448] for later pieces), [64] feed count,
    // [128] feed ticket, [192] feed-complete flag, [256] work counter of the pick-up launch, [320] watchdog,
    // [512] consumer CTAs resident,
    // [1024] hist[256], [2048] cursor[256] (per piece)
    ODL_CUDA(cudaMemsetAsync(feed, 0xFF, (size_t)n * sizeof(int), s));
    // ordering of one piece [lo, hi) of the table -> index[lo..hi) (global row numbers)
    auto order_piece = [&](long long lo, long long hi, int piece, cudaStream_t so_) -> int {
      OdlOrderArgs R{};
      const long long np = hi - lo;
      R.theta = A.theta + lo * m->n_param; R.n = np; R.bins = static_cast<unsigned char*>(bbins.p) + lo;
      R.hist = cnt(1024 + 2048 * piece); R.cursor = cnt(2048 + 2048 * piece); R.index = index + lo;
      R.row_base = (int)lo; R.pad_ = 0;
      OdlData Dl = D;
      void* p1[] = {&Dl, &R};
      void* p2[] = {&R};
      const unsigned g1 = (unsigned)std::max<long long>(1, std::min<long long>((np + 255) / 256, (long long)m->sm_count * 8));
      const unsigned g3 = (unsigned)std::max<long long>(1, std::min<long long>((np + 2047) / 2048, (long long)m->sm_count * 8));
      int r;
      if ((r = launch(m, m->k_order_key, g1, 256, 0, so_, p1))) return r;
      if ((r = launch(m, m->k_order_scan, 1, ODL_ORDER_BINS, 0, so_, p2))) return r;
      return launch(m, m->k_order_scatter, g3, 256, 0, so_, p2);
    };
    // piece boundaries: [0, n/2, n] rounded to 1024 rows (one piece when the table is already on the device).  Measured
    // on B200, 1M two_i rows through host buffers: two halves 190 M solves/s, three pieces (1/8, 3/8, 1/2) 180 M/s --
    // every extra bulk launch ends on its own stragglers, which costs more than the shorter wait for the first piece.
    const bool beside_first_quarter = !m->coop_model() && !(flags & ODL_AUTO_SEQUENTIAL);
    // Two pieces.  Three (10 / 30 / 60 %, ODL_PIECES=3) keep the SMs fed while the table is still arriving but were measured
    // SLOWER (5.0 against 4.3 ms): every piece is cost-ordered on its own, the rows the stiff pass has to take start when
    // their PIECE starts, and from there they need ~0.9 ms to reach that pass and up to 1.5 ms in it -- the last piece
    // has to start early, so it has to be the big one.
    const int n_piece = chunked ? ((beside_first_quarter && getenv("ODL_PIECES") && atoi(getenv("ODL_PIECES")) == 3) ? 3 : 2) : 1;
    long long cut[4] = {0, n, n, n};
    if (chunked) {
      // share of the rows in the first piece: its upload is the one nothing hides.  Stiff pass beside the bulk pass, the
      // pieces' launches overlapping (below): 1M rows through pinned buffers 4.62 ms at 1/4, 4.40 at 0.15, 4.33 at 0.1,
      // 4.40 at 0.05 (round 2 before the overlap, one launch after the other: 5.44 at 1/2, 4.88 at 1/3 .. 1/5, 4.85 at 0.15).
      // Stiff pass after the bulk pass: halves (5.10 / 5.04 / 5.12 at 1/2, 1/3, 1/4).
      // With the hand-over (the stiff pass's rows are cheap and no longer sit on the critical path) the first piece is best
      // sized so that its sweep covers the upload of the rest: 3.61 ms at 0.1, 3.47 at 0.15, 3.41 at 0.2 and 0.25, 3.46 at 0.3.
      double first = beside_first_quarter ? 0.20 : 0.5;
      if (const char* e = getenv("ODL_FIRST_PIECE")) first = std::min(0.9, std::max(0.05, atof(e)));   // development knob
      cut[1] = (((long long)(n * first) + 1023) / 1024) * 1024;
      // three pieces (10 %, 30 %, 60 %): the upload runs at ~4x the speed of the sweep, so once the first piece is there
      // every later one lands before the SMs run out of rows (with two pieces they idled ~0.15 ms waiting for the second)
      if (n_piece == 3) {
        double second = 0.30;
        if (const char* e = getenv("ODL_SECOND_PIECE")) second = std::min(0.9, std::max(0.05, atof(e)));   // development knob
        if (!getenv("ODL_FIRST_PIECE")) cut[1] = (((long long)(n * 0.10) + 1023) / 1024) * 1024;
        cut[2] = std::min(n, cut[1] + (((long long)(n * second) + 1023) / 1024) * 1024);
      }
    }
    if (chunked) {
      ODL_CUDA(cudaEventRecord(m->ev_fork, s));                  // behind the previous call on this handle
      ODL_CUDA(cudaStreamWaitEvent(m->aux, m->ev_fork, 0));
      for (int c = 0; c < n_piece; ++c) {
        ODL_CUDA(cudaMemcpyAsync(const_cast<double*>(A.theta) + cut[c] * m->n_param, theta + cut[c] * m->n_param,
                                 (size_t)(cut[c + 1] - cut[c]) * m->n_param * sizeof(double), cudaMemcpyHostToDevice, m->aux));
        ODL_CUDA(cudaEventRecord(m->ev_chunk[c], m->aux));
      }
    }
    const unsigned block0 = pick_block(D, m->block);
    const size_t smem0 = smem_bytes(D, (int)block0), smem_t = smem_bytes(D, 32);
    if (!coop_bulk && smem0 > 48 * 1024) ODL_CU(g_drv.FuncSetAttribute(m->k_sweep, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem0));
    if (smem_t > 48 * 1024) ODL_CU(g_drv.FuncSetAttribute(k_tail, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem_t));
    // both kernels ask for the same L1/shared split (kernels with different carve-outs cannot share an SM, and a
    // carve-out switch drains the SM first)
    if (!coop_bulk) ODL_CU(g_drv.FuncSetAttribute(m->k_sweep, CU_FUNC_ATTRIBUTE_PREFERRED_SHARED_MEMORY_CARVEOUT, 100));
    ODL_CU(g_drv.FuncSetAttribute(k_tail, CU_FUNC_ATTRIBUTE_PREFERRED_SHARED_MEMORY_CARVEOUT, 100));
    int per_sm0 = 1, per_sm_t = 0;
    if (!coop_bulk) ODL_CU(g_drv.OccupancyMaxActiveBlocksPerMultiprocessor(&per_sm0, m->k_sweep, (int)block0, smem0));
    ODL_CU(g_drv.OccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_t, k_tail, 32, smem_t));
    if (per_sm0 < 1 || per_sm_t < 1) return fail(ODL_ECUDA, "sweep kernel does not fit on an SM (shared memory / registers)");
    // What the DOPRI5 pass had reached when it gave a row up travels with the row (OdlSweepArgs.handover): the stiff pass
    // continues from there.  [n][2 + n_state + stage_stride] doubles, written only for the rows of the feed list.
    double* handover = nullptr;
    int handover_stride = 0;
    // (a rule on the table's size alone, so that results do not depend on what else lives in device memory: tables whose
    // records would take more than 8 GB -- 23M rows of the two_i model -- go without, their stiff rows start from t0)
    const size_t handover_bytes = (size_t)n * (2 + m->n_state + D.stage_stride) * sizeof(double);
    if (!coop_bulk && tail_solver != ODL_SOLVER_RADAU5 && !(flags & ODL_AUTO_NO_HANDOVER) && handover_bytes <= ((size_t)8 << 30)) {
      handover_stride = 2 + m->n_state + D.stage_stride;
      DevBuf& bh = m->handover;
      if ((rc = bh.ensure(handover_bytes))) return rc;
      handover = static_cast<double*>(bh.p);
    }
    // SMs set aside for the stiff pass when it runs beside the bulk pass
    int tail_sms = 0;
    unsigned tail_cluster = 1;
    size_t smem_wide = 0;
    unsigned wide_block = 256;
    if (const char* e = getenv("ODL_WIDE_BLOCK")) wide_block = (unsigned)std::max(32, atoi(e));   // development knob
    if (beside) {
      // measured on B200, 1M two_i prior draws (15k rows, ~7M BDF steps for the consumer), back-to-back calls behind an
      // L2 flush as bench.py times them, consumer CTAs in clusters of 2: 34 SMs 3.98 ms, 36 3.88, 40 3.73, 42 3.62,
      // 44 3.67 (stiff pass after the bulk pass: 4.1-4.3) -- between a quarter and a third of the SMs
      // Round 2, with the DOPRI5 cap at 704 (projection check still at 384: 9.4k rows and 5.0M BDF steps instead of 15k
      // and 7.5M) the consumer keeps up on fewer SMs: 704 / 42 SMs 3.82 ms (its rows arrive later and its longest end after
      // the bulk pass), 36 3.44, 34 3.38, 32 3.32, 30 3.37, 28 3.61 (backlog when the bulk pass ends); cap 640 / 34 3.40,
      // 736 / 32 3.33, 768 / 34 3.66 (profiles/r2q_beside_grid.log)
      // With the HAND-OVER (the stiff pass continues from where the DOPRI5 pass stopped: 1.13M instead of 4.95M BDF steps for
      // the 9.4k rows of the two_i sweep -- the hard part of these rows is their start, and DOPRI5 has done it) the consumer
      // needs a fraction of that: 32 SMs 3.19 ms, 28 3.09, 24 2.99, 20 2.90, 16 2.82, 12 2.74, 8 3.26 (backlog).  A tenth
      // of the SMs, with a margin over the edge.
      // (two-piece host-memory sweeps: two SMs more -- the rows of the second piece reach the consumer in a shorter time)
      const int share = handover ? 95 : 220;                     // per mille of the SMs
      tail_sms = so && so->tail_warps > 0 ? so->tail_warps : (m->sm_count * share + 500) / 1000 + ((chunked && !handover) ? 2 : 0);
      tail_sms = std::max(1, std::min(tail_sms, m->sm_count / 3));
      // Placement: on an idle GPU the block scheduler packs the consumer's CTAs onto neighbouring SMs, behind other
      // work it scatters them -- and a consumer SM whose TPC partner runs the bulk kernel steps 8-12 % slower
      // (measured with %smid: both kernels are large, 46 and 58 KB of code).  Clusters of CTAs keep the consumer on
      // whole TPCs (2) or larger pieces of a GPC.
      tail_cluster = 2;
      if (const char* e = getenv("ODL_TAIL_CLUSTER")) tail_cluster = (unsigned)std::max(1, atoi(e));   // development knob
      tail_sms = std::max((int)tail_cluster, tail_sms / (int)tail_cluster * (int)tail_cluster);
      // a CTA of the consumer must leave no room for a bulk CTA on its SM: ask for (SM shared memory) - (one bulk CTA)
      const size_t sm_total = 228 * 1024, reserve = 1024;
      smem_wide = std::max(smem_bytes(D, (int)wide_block), sm_total - reserve - (smem0 + reserve) + 1024);
      smem_wide = std::min<size_t>(smem_wide, 227 * 1024);
      if (smem_bytes(D, (int)wide_block) > 227 * 1024) return fail(ODL_ECUDA, "stiff pass: tables + staging exceed shared memory");
      ODL_CU(g_drv.FuncSetAttribute(k_tail, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem_wide));
    }
    OdlOpts O0 = O; O0.stiff_check = 1; O0.max_steps = std::min(cap0, O.max_steps);
    // projection check at 3/4 of the cap: a system whose progress projects beyond the cap leaves there.  (At cap0/2 the
    // check has recall ~100 % but precision ~50 %: it doubled the stiff pass's load with long NON-stiff systems -- 19 %
    // of the zero_i sweep, 1-2 % of one_i / two_i, measured with tools/variant_ab.py.)
    // (and at most ODL_AUTO_EARLY_DEFAULT: the rows that leave here are on the stiff pass's critical path -- ~860 BDF steps of
    // 1.7 us -- and must reach it early whatever the cap is)
    O0.early_check_steps = so && so->early_check_steps != 0 ? std::max(0, so->early_check_steps)
                                                            : std::min(O0.max_steps * 3 / 4, ODL_AUTO_EARLY_DEFAULT);
    OdlSweepArgs A0 = A;
    A0.index = ordered ? index : nullptr;
    A0.handover = handover; A0.handover_stride = handover_stride;
    A0.defer_list[0] = A0.defer_list[1] = feed; A0.defer_count[0] = A0.defer_count[1] = cnt(64);
    OdlOpts O2 = O; O2.stiff_check = 0;
    // the pass over the feed list AFTER the bulk pass: as few lanes per warp as spreading the entries over every
    // resident warp takes (a warp pays for the union of its lanes' branches: 1.80 -> 1.61 ms at 11 of 32 lanes)
    O2.lanes = so && so->tail_lanes > 0 ? std::min(32, so->tail_lanes) : -1;
    OdlSweepArgs A2 = A;
    A2.index = feed; A2.index_count = cnt(64); A2.counter = ctr(256);
    A2.handover = handover; A2.handover_stride = handover_stride;
    A2.defer_list[0] = A2.defer_list[1] = nullptr; A2.defer_count[0] = A2.defer_count[1] = nullptr;
    const int bulk_sms = m->sm_count - tail_sms;
    const unsigned grid0 = (unsigned)std::max<long long>(1, std::min<long long>((n + block0 - 1) / block0, (long long)per_sm0 * bulk_sms));
    const unsigned grid_t = (unsigned)std::max<long long>(1, std::min<long long>((n + 31) / 32, (long long)per_sm_t * m->sm_count));
    OdlData Dl = D;
    if (beside) {
      // consumer first, on the helper stream: it is resident on its SMs before the bulk grid is launched
      OdlOpts Oc = O2; Oc.lanes = 0;
      OdlSweepArgs Ac = A2;
      Ac.counter = nullptr; Ac.feed_ticket = ctr(128); Ac.feed_done = cnt(192); Ac.watchdog = cnt(320); Ac.resident = cnt(512);
      void* pc[] = {&Dl, &Oc, &Ac};
      ODL_CUDA(cudaEventRecord(m->ev_fork, s));                  // after the counter block and the feed list are reset
      ODL_CUDA(cudaStreamWaitEvent(m->aux2, m->ev_fork, 0));
      if ((rc = launch_clustered(m, k_tail, (unsigned)tail_sms, wide_block, smem_wide, m->aux2, pc, tail_cluster))) return rc;
      ODL_CUDA(cudaEventRecord(m->ev_aux, m->aux2));
    }
    if (chunked) ODL_CUDA(cudaStreamWaitEvent(s, m->ev_chunk[0], 0));
    if (ordered && (rc = order_piece(0, cut[1], 0, s))) return rc;
    if (beside) {
      // the bulk grid must not arrive before the consumer's CTAs have their SMs (see odl_gate_kernel)
      const int* res = cnt(512);
      int want = tail_sms, spins = 4000;                         // ~2 ms at most
      void* pg[] = {&res, &want, &spins};
      if ((rc = launch(m, m->k_gate, 1, 1, 0, s, pg))) return rc;
    }
    ODL_CUDA(cudaEventRecord(m->evp[1], s));
    void* pb[] = {&Dl, &O0, &A0};
    if (coop_bulk) { if ((rc = go_coop(s, O0, A0, n))) return rc; }              // n > 8: several lanes per system
    else if (!chunked) { if ((rc = launch(m, m->k_sweep, grid0, block0, smem0, s, pb))) return rc; }
    else {
      const OdlSweepArgs Aall = A0;
      A0.n = cut[1];                                         // pb points at A0: first piece, index[0..cut[1])
      const unsigned g0 = (unsigned)std::max<long long>(1, std::min<long long>((cut[1] + block0 - 1) / block0, (long long)grid0));
      // The later pieces are ordered and swept on the HELPER stream, behind their own upload: their grids wait for SM
      // slots, not for the end of the launch before them.  One after the other on one stream every bulk launch ended on
      // its own stragglers -- a launch cannot end before its longest row, cap0 attempts of ~2.4 us, whatever the size
      // of its piece -- and with the cap at 704 that idle tail made the two-piece call SLOWER than with 512 (5.1 against
      // 4.8 ms per 1M rows through pinned buffers); now the CTAs of the next piece take the slots as the stragglers'
      // neighbours retire.  (The ordering kernels of a later piece get their first slot the same way.)
      ODL_CUDA(cudaEventRecord(m->ev_fork, s));              // counters reset, consumer resident, first piece ordered
      if ((rc = launch(m, m->k_sweep, g0, block0, smem0, s, pb))) return rc;
      for (int c = 1; c < n_piece; ++c) {
        if (cut[c + 1] <= cut[c]) continue;
        cudaStream_t sp = m->piece_stream[c - 1];            // a stream of its own: behind ITS upload, beside the others
        ODL_CUDA(cudaStreamWaitEvent(sp, m->ev_fork, 0));
        ODL_CUDA(cudaStreamWaitEvent(sp, m->ev_chunk[c], 0));
        if ((rc = order_piece(cut[c], cut[c + 1], c, sp))) return rc;
        OdlSweepArgs Ac = Aall;
        Ac.n = cut[c + 1] - cut[c]; Ac.index = index + cut[c]; Ac.counter = ctr(384 + 64 * (c - 1));   // chunked => ordered
        void* pbc[] = {&Dl, &O0, &Ac};
        const unsigned gc = (unsigned)std::max<long long>(1, std::min<long long>((Ac.n + block0 - 1) / block0, (long long)grid0));
        if ((rc = launch(m, m->k_sweep, gc, block0, smem0, sp, pbc))) return rc;
        ODL_CUDA(cudaEventRecord(m->ev_piece[c - 1], sp));
        ODL_CUDA(cudaStreamWaitEvent(s, m->ev_piece[c - 1], 0));   // every bulk launch has ended: the feed is complete
      }
    }
    ODL_CUDA(cudaEventRecord(m->evp[0], s));
    if (beside) {
      int* flag = cnt(192);
      void* pf[] = {&flag};
      if ((rc = launch(m, m->k_feed_done, 1, 1, 0, s, pf))) return rc;
      // The bulk pass has ended and left most SMs idle: a second consumer over the SAME ticket counter -- single-warp CTAs
      // on every SM, as many lanes per warp as spreading what is left of the feed takes -- helps the first one finish.
      // (Without it the sweep ends when the 20-30 % of the SMs the first consumer owns have worked the feed off alone:
      // one_i 0.37 ms, zero_i 2.1 ms after the bulk pass, and the split of the SMs had to be tuned per model.)
      if (!(flags & ODL_AUTO_NO_HELPER)) {
        OdlOpts Oh = O2; Oh.lanes = -1;
        OdlSweepArgs Ah = A2;
        Ah.counter = nullptr; Ah.feed_ticket = ctr(128); Ah.feed_done = cnt(192); Ah.watchdog = cnt(320); Ah.resident = nullptr;
        void* ph[] = {&Dl, &Oh, &Ah};
        ODL_CU(g_drv.FuncSetAttribute(k_tail, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)std::max(smem_t, smem_wide)));
        if ((rc = launch(m, k_tail, grid_t, 32, smem_t, s, ph))) return rc;
      }
      ODL_CUDA(cudaStreamWaitEvent(s, m->ev_aux, 0));            // the consumer has drained the feed
    }
    // the stiff pass proper (sequential mode) / the pick-up of entries the consumer left (beside mode: normally none)
    void* pt[] = {&Dl, &O2, &A2};
    if (smem_t > 48 * 1024 || beside) ODL_CU(g_drv.FuncSetAttribute(k_tail, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)std::max(smem_t, smem_wide)));
    if ((rc = launch(m, k_tail, grid_t, 32, smem_t, s, pt))) return rc;
    m->n_pass = 3;
  }
  ODL_CUDA(cudaEventRecord(m->ev1, s));
  m->timed = true;
  if ((rc = st.finish())) return rc;
  return 0;
}

extern "C" int odl_trajectory(odl_model* m, const odl_solver_opts* so, long long n, const double* theta,
                              const double* y0_or_null, int mem, double* traj, int* status, int* nsteps, void* stream) {
  if (!m || !m->on_gpu) return fail(ODL_ENODEVICE, "odl_trajectory: model is not loaded on a GPU (no CPU fallback exists)");
  if (!m->grid.set) return fail(ODL_EINVAL, "odl_trajectory: call odl_model_set_grid first");
  if (n < 0 || (n > 0 && (!theta || !traj || !status || !nsteps))) return fail(ODL_EINVAL, "odl_trajectory: null buffer");
  if (n == 0) return 0;
  ODL_ON_DEVICE(m);
  int rc;
  if ((rc = units_ready(m, {U_TRAJ}))) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if ((rc = serialize_after_previous_call(m, s))) return rc;
  Staging st{m, s, 0, mem};
  OdlTrajArgs A{};
  OdlData D = m->grid.d;
  if ((rc = st.in(theta, (size_t)n * m->n_param, &A.theta))) return rc;
  if ((rc = st.in(y0_or_null, (size_t)n * m->n_state, &A.y0))) return rc;
  if ((rc = st.inout(traj, (size_t)n * D.n_slot * m->n_state, &A.traj, false))) return rc;
  if ((rc = st.inout(status, (size_t)n, &A.status, false))) return rc;
  if ((rc = st.inout(nsteps, (size_t)n, &A.nsteps, false))) return rc;
  A.n = n; A.counter = nullptr;
  OdlOpts O; fill_opts(O, so);
  const size_t smem = (size_t)D.n_slot * sizeof(double);
  const unsigned block = 32;
  unsigned grid = (unsigned)((n + block - 1) / block);
  ODL_CUDA(cudaEventRecord(m->ev0, s));
  void* params[] = {&D, &O, &A};
  m->n_pass = 1;
  if ((rc = launch(m, m->k_traj, grid, block, smem, s, params))) return rc;
  ODL_CUDA(cudaEventRecord(m->ev1, s));
  m->timed = true;
  return st.finish();
}

extern "C" int odl_mcmc(odl_model* m, const odl_solver_opts* so, const odl_mcmc_opts* mo, const odl_mcmc_io* io, int mem,
                        void* stream) {
  if (!m || !m->on_gpu) return fail(ODL_ENODEVICE, "odl_mcmc: model is not loaded on a GPU (no CPU fallback exists)");
  if (!m->data.set) return fail(ODL_EINVAL, "odl_mcmc: call odl_model_set_data first");
  if (!mo || !io || !io->theta || !io->chain_state) return fail(ODL_EINVAL, "odl_mcmc: null argument");
  if (mo->n_chain < 1 || mo->nits < 2) return fail(ODL_EINVAL, "odl_mcmc: need n_chain >= 1 and nits >= 2");
  if (mo->n_walk < 0 || mo->n_walk > m->n_param || (mo->n_walk > 0 && !mo->walk)) return fail(ODL_EINVAL, "odl_mcmc: bad walk list");
  const int solver = so ? so->solver : ODL_SOLVER_DOPRI5;
  if (solver < 0 || solver > ODL_SOLVER_BDF) return fail(ODL_EINVAL, "odl_mcmc: unknown solver");
  const int P = m->n_param, C = mo->n_chain, n_iter = mo->nits - 1;
  int it_begin = mo->it_begin, it_end = mo->it_end;
  if (it_begin == 0 && it_end == 0) { it_begin = 1; it_end = mo->nits; }
  if (it_begin < 1 || it_end > mo->nits || it_begin > it_end) return fail(ODL_EINVAL, "odl_mcmc: bad iteration range");
  const int n_keep = std::max(0, n_iter - mo->burnin);
  const int stride = mo->row_stride > 0 ? mo->row_stride : P + 5;
  if (stride < P + 5) return fail(ODL_EINVAL, "odl_mcmc: row_stride < n_param+5");
  if (mo->rng_mode == ODL_RNG_HOST_STREAMS && (!io->z || !io->u)) return fail(ODL_EINVAL, "odl_mcmc: host streams need z and u");
  if (mo->rng_mode == ODL_RNG_FORCED && (!io->forced || !io->u)) return fail(ODL_EINVAL, "odl_mcmc: forced mode needs proposals and u");
  ODL_ON_DEVICE(m);
  int rc;
  const bool use_coop = solver == ODL_SOLVER_DOPRI5 && m->coop_model() && mo->speculate <= 0;
  if ((rc = units_ready(m, {use_coop ? U_MCMC_COOP : (solver == ODL_SOLVER_DOPRI5 ? U_MCMC : (solver == ODL_SOLVER_ROS23 ? U_MCMC_ROS :
                            (solver == ODL_SOLVER_RADAU5 ? U_MCMC_RADAU : (solver == ODL_SOLVER_BDF ? U_MCMC_BDF : U_MCMC_AUTO))))})))
    return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if ((rc = serialize_after_previous_call(m, s))) return rc;
  Staging st{m, s, 0, mem};
  OdlMcmcArgs A{};
  if ((rc = st.inout(io->theta, (size_t)C * P, &A.theta_cur, true))) return rc;
  if ((rc = st.inout(io->chain_state, (size_t)C * ODL_CHAIN_STATE, &A.chain_state, true))) return rc;
  if ((rc = st.inout(io->best_theta, (size_t)C * P, &A.best_theta, it_begin > 1))) return rc;
  if ((rc = st.in(io->chain_ids, (size_t)C, &A.chain_ids))) return rc;
  if ((rc = st.in(io->prior_table, (size_t)P * 4, &A.prior))) return rc;
  if ((rc = st.inout(io->samples, (size_t)C * n_keep * stride, &A.samples, it_begin > 1))) return rc;
  if ((rc = st.inout(io->summaries, (size_t)C * (1 + 2 * P), &A.summaries, true))) return rc;
  if ((rc = st.in(io->z, (size_t)C * n_iter * mo->n_walk, &A.z))) return rc;
  if ((rc = st.in(io->u, (size_t)C * n_iter, &A.u))) return rc;
  if ((rc = st.in(io->forced, (size_t)C * n_iter * P, &A.forced))) return rc;
  if ((rc = st.inout(io->trace_chinew, (size_t)C * n_iter, &A.trace_chinew, it_begin > 1))) return rc;
  if ((rc = st.inout(io->trace_accept, (size_t)C * n_iter, &A.trace_accept, it_begin > 1))) return rc;
  if ((rc = st.inout(io->fail_count, (size_t)C, &A.fail_count, true))) return rc;
  if ((rc = st.inout(io->step_count, (size_t)C, &A.step_count, true))) return rc;
  A.n_chain = C; A.chain_offset = mo->chain_offset; A.it_begin = it_begin; A.it_end = it_end;
  A.burnin = mo->burnin; A.n_keep = n_keep; A.row_stride = stride; A.rng_mode = mo->rng_mode;
  if (mo->sample_layout != ODL_SAMPLES_CHAIN_MAJOR && mo->sample_layout != ODL_SAMPLES_ITERATION_MAJOR)
    return fail(ODL_EINVAL, "odl_mcmc: unknown sample_layout");
  if (mo->sample_layout == ODL_SAMPLES_ITERATION_MAJOR) { A.smp_chain_pitch = stride; A.smp_row_pitch = (long long)C * stride; }
  else { A.smp_chain_pitch = (long long)n_keep * stride; A.smp_row_pitch = stride; }
  A.n_walk = mo->n_walk; A.pnum = mo->pnum;
  for (int j = 0; j < ODL_MAX_WALK; ++j) A.walk[j] = -1;
  for (int j = 0; j < mo->n_walk; ++j) {
    if (mo->walk[j] < 0 || mo->walk[j] >= P) return fail(ODL_EINVAL, "odl_mcmc: walk index out of range");
    A.walk[j] = mo->walk[j];
  }
  A.step_sd = mo->step_sd > 0 ? mo->step_sd : 0.05;
  A.seed = mo->seed; A.n_iter_total = n_iter;
  A.stop_failed = mo->stop_failed_chains ? 1 : 0; A.pad2_ = 0;
  OdlOpts O; fill_opts(O, so);
  OdlData D = m->data.d;
  const bool warp_cta = solver == ODL_SOLVER_RADAU5 || solver == ODL_SOLVER_BDF || solver == ODL_SOLVER_AUTO;   // compiled for one warp per CTA
  // Prefetching MH: K lanes per chain evaluate K iterations at once along the all-rejected path (odl_mcmc_body).
  // The chain itself does not depend on K; K only fills a GPU that few chains would leave latency-bound.
  // Automatic: the largest power of two that keeps chains*K within half of what the kernel can hold, at most 16.
  int K = mo->speculate;
  if (K <= 0) {
    const long long resident = (long long)m->sm_count * (warp_cta ? 8 * 32 : 512);
    K = 1;
    while (K < 16 && (long long)C * K * 2 * 2 <= resident) K *= 2;
  }
  if (K > 32 || (K & (K - 1))) return fail(ODL_EINVAL, "odl_mcmc: speculate must be a power of two <= 32");
  A.spec = K;
  const long long threads = (long long)C * K;
  // few threads: spread them over the SMs with one warp per CTA; many: full CTAs
  unsigned block = pick_block(D, m->block);
  while (block > 32 && threads < (long long)m->sm_count * block * 2) block /= 2;
  if (warp_cta) block = 32;
  const size_t smem = smem_bytes(D, (int)block);
  unsigned grid = (unsigned)((threads + block - 1) / block);
  if (use_coop) {
    // n > 8: one chain per K groups of lanes (odl_mcmc_coop_kernel); the chain is the same chain as with every other
    // mapping.  speculate = -K asks for K groups per chain (K * coop_lanes <= 32), 0 = automatic as above
    const size_t smem_c = coop_smem_bytes(m, D);
    if (smem_c > 227 * 1024) return fail(ODL_ECUDA, "cooperative MCMC kernel: tables + staging exceed shared memory");
    if (smem_c > 48 * 1024) ODL_CU(g_drv.FuncSetAttribute(m->k_mcmc_coop, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem_c));
    int per_sm_c = 0;
    ODL_CU(g_drv.OccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_c, m->k_mcmc_coop, (int)kCoopBlock, smem_c));
    const int kmax = 32 / m->coop;
    int Kc = -mo->speculate;
    if (Kc <= 0) {
      const long long resident_c = (long long)m->sm_count * std::max(1, per_sm_c) * kCoopBlock;
      Kc = 1;
      while (Kc < kmax && (long long)C * m->coop * Kc * 2 * 2 <= resident_c) Kc *= 2;
    }
    if (Kc > kmax || (Kc & (Kc - 1))) return fail(ODL_EINVAL, "odl_mcmc: -speculate must be a power of two with speculate * coop_lanes <= 32");
    const long long lanes = (long long)C * m->coop * Kc;
    A.spec = Kc;
    ODL_CUDA(cudaEventRecord(m->ev0, s));
    void* params_c[] = {&D, &O, &A};
    m->n_pass = 1;
    if ((rc = launch(m, m->k_mcmc_coop, (unsigned)((lanes + kCoopBlock - 1) / kCoopBlock), kCoopBlock, smem_c, s, params_c))) return rc;
    ODL_CUDA(cudaEventRecord(m->ev1, s));
    m->timed = true;
    return st.finish();
  }
  CUfunction f = (solver == ODL_SOLVER_DOPRI5) ? m->k_mcmc : (solver == ODL_SOLVER_ROS23 ? m->k_mcmc_ros :
                 (solver == ODL_SOLVER_RADAU5 ? m->k_mcmc_radau : (solver == ODL_SOLVER_BDF ? m->k_mcmc_bdf : m->k_mcmc_auto)));
  if (solver == ODL_SOLVER_AUTO) {
    // per solve: DOPRI5, and the same solve again on BDF when it gives up.  pass_cap0 > 0: "gives up" = that many attempted
    // steps (then Hairer's test runs only if the caller asked for it, so that a chain whose solves all stay within the
    // budget is the plain DOPRI5 chain under max_steps = pass_cap0); else Hairer's test routes.
    O.explicit_cap = so && so->pass_cap0 > 0 ? so->pass_cap0 : 0;
    O.stiff_check = O.explicit_cap > 0 ? (so->stiff_check ? 1 : 0) : 1;
  }
  ODL_CUDA(cudaEventRecord(m->ev0, s));
  void* params[] = {&D, &O, &A};
  m->n_pass = 1;
  if ((rc = launch(m, f, grid, block, smem, s, params))) return rc;
  ODL_CUDA(cudaEventRecord(m->ev1, s));
  m->timed = true;
  return st.finish();
}

/* development aid: the feed timeline of the last AUTO sweep run with ODL_TIMELINE=1 (3 timestamps per feed entry) */
extern "C" int odl_debug_timeline(odl_model* m, long long* out, long long entries) {
  if (!m || !out || !m->on_gpu || !m->scratch[20].p) return fail(ODL_EINVAL, "odl_debug_timeline: nothing recorded");
  ODL_ON_DEVICE(m);
  if ((size_t)entries * 3 * sizeof(long long) > m->scratch[20].cap) return fail(ODL_EINVAL, "odl_debug_timeline: count out of range");
  ODL_CUDA(cudaMemcpy(out, m->scratch[20].p, (size_t)entries * 3 * sizeof(long long), cudaMemcpyDeviceToHost));
  return 0;
}

/* development aid: the first `count` ints of the counter block of the last sweep (see odl_sweep) */
extern "C" int odl_debug_counters(odl_model* m, int* out, int count) {
  if (!m || !out || !m->on_gpu) return fail(ODL_EINVAL, "odl_debug_counters: bad argument");
  if (count < 0 || count > 1024) return fail(ODL_EINVAL, "odl_debug_counters: count out of range");
  ODL_ON_DEVICE(m);
  ODL_CUDA(cudaMemcpy(out, m->counter.p, (size_t)count * sizeof(int), cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int odl_model_last_pass_ms(odl_model* m, float* ms3) {
  if (!m || !ms3) return fail(ODL_EINVAL, "null argument");
  if (!m->on_gpu || !m->timed) return fail(ODL_EINVAL, "no kernel has been launched on this model yet");
  ODL_CUDA(cudaEventSynchronize(m->ev1));
  ms3[0] = ms3[1] = ms3[2] = 0.f;
  if (m->n_pass == 3) {
    ODL_CUDA(cudaEventElapsedTime(&ms3[0], m->ev0, m->evp[1]));
    ODL_CUDA(cudaEventElapsedTime(&ms3[1], m->evp[1], m->evp[0]));
    ODL_CUDA(cudaEventElapsedTime(&ms3[2], m->evp[0], m->ev1));
  } else {
    ODL_CUDA(cudaEventElapsedTime(&ms3[0], m->ev0, m->ev1));
  }
  return 0;
}

extern "C" int odl_model_last_kernel_ms(odl_model* m, float* ms) {
  if (!m || !ms) return fail(ODL_EINVAL, "null argument");
  if (!m->on_gpu || !m->timed) return fail(ODL_EINVAL, "no kernel has been launched on this model yet");
  ODL_CUDA(cudaEventSynchronize(m->ev1));
  ODL_CUDA(cudaEventElapsedTime(ms, m->ev0, m->ev1));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Chain-start selection (Framework.py:993-1016): the rows of a survey whose chi lies below the threshold, in
// survey order (what `fitsurvey[fitsurvey['chi'] < cutchi]` keeps; NaN never qualifies), and the gather of the
// rows the caller's random picks name.  Ordered compaction: per-tile counts, one scan, per-tile ranks.
// ---------------------------------------------------------------------------------------------
#define ODL_SEL_TILE 1024
__global__ void __launch_bounds__(256) odl_select_count_kernel(const double* chi, long long n, double cut, int* tile_count) {
  __shared__ int total;
  if (threadIdx.x == 0) total = 0;
  __syncthreads();
  const long long base = (long long)blockIdx.x * ODL_SEL_TILE;
  int c = 0;
  for (int k = threadIdx.x; k < ODL_SEL_TILE; k += blockDim.x) {
    const long long i = base + k;
    if (i < n && chi[i] < cut) ++c;
  }
  for (int m = 16; m > 0; m >>= 1) c += __shfl_xor_sync(0xffffffffu, c, m);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&total, c);
  __syncthreads();
  if (threadIdx.x == 0) tile_count[blockIdx.x] = total;
}
__global__ void __launch_bounds__(1024) odl_select_scan_kernel(int* tile_count, int n_tile, long long* total_out) {
  // exclusive prefix of the tile counts, in place; one CTA walks the tiles in chunks of blockDim.x
  __shared__ int s[1024];
  __shared__ long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int t0 = 0; t0 < n_tile; t0 += blockDim.x) {
    const int t = t0 + threadIdx.x;
    const int v = t < n_tile ? tile_count[t] : 0;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < (int)blockDim.x; off <<= 1) {
      const int a = threadIdx.x >= (unsigned)off ? s[threadIdx.x - off] : 0;
      __syncthreads();
      s[threadIdx.x] += a;
      __syncthreads();
    }
    if (t < n_tile) tile_count[t] = (int)(carry + s[threadIdx.x] - v);
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry += s[threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = carry;
}
__global__ void __launch_bounds__(256) odl_select_write_kernel(const double* chi, long long n, double cut, const int* tile_offset,
                                                               int* index_out) {
  // one warp per 256 rows of the tile, rows in order: ballot + popc give the rank inside a 32-row slice
  __shared__ int warp_count[8][4];
  const long long base = (long long)blockIdx.x * ODL_SEL_TILE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned masks[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const long long i = base + (long long)(warp * 4 + r) * 32 + lane;
    masks[r] = __ballot_sync(0xffffffffu, i < n && chi[i] < cut);
    if (lane == 0) warp_count[warp][r] = __popc(masks[r]);
  }
  __syncthreads();
  int before = tile_offset[blockIdx.x];
  for (int w = 0; w < warp; ++w) for (int r = 0; r < 4; ++r) before += warp_count[w][r];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const long long i = base + (long long)(warp * 4 + r) * 32 + lane;
    if ((masks[r] >> lane) & 1u) index_out[before + __popc(masks[r] & ((1u << lane) - 1u))] = (int)i;
    before += __popc(masks[r]);
  }
}
__global__ void __launch_bounds__(256) odl_gather_rows_kernel(const double* src, int row_len, const int* index, const long long* picks,
                                                              long long n_pick, double* dst) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_pick * row_len) return;
  const long long r = e / row_len;
  const int c = (int)(e - r * row_len);
  const long long row = index ? (long long)index[picks[r]] : picks[r];
  dst[e] = src[row * row_len + c];
}

extern "C" int odl_select_below(odl_model* m, const double* chi_dev, long long n, double cut, int* index_dev, long long* count_host,
                                void* stream) {
  if (!m || !m->on_gpu) return fail(ODL_ENODEVICE, "odl_select_below: model is not loaded on a GPU (no CPU fallback exists)");
  if (n < 0 || n > 2147483647LL || !count_host || (n > 0 && (!chi_dev || !index_dev))) return fail(ODL_EINVAL, "odl_select_below: bad argument");
  *count_host = 0;
  if (n == 0) return 0;
  ODL_ON_DEVICE(m);
  cudaStream_t s = (cudaStream_t)stream;
  const int n_tile = (int)((n + ODL_SEL_TILE - 1) / ODL_SEL_TILE);
  DevBuf& b = m->scratch[23];
  int rc = b.ensure((size_t)n_tile * sizeof(int) + 16);
  if (rc) return rc;
  int* tiles = static_cast<int*>(b.p);
  long long* total = reinterpret_cast<long long*>(static_cast<char*>(b.p) + (((size_t)n_tile * sizeof(int) + 7) & ~(size_t)7));
  odl_select_count_kernel<<<n_tile, 256, 0, s>>>(chi_dev, n, cut, tiles);
  odl_select_scan_kernel<<<1, 1024, 0, s>>>(tiles, n_tile, total);
  odl_select_write_kernel<<<n_tile, 256, 0, s>>>(chi_dev, n, cut, tiles, index_dev);
  g_launches.fetch_add(3);
  ODL_CUDA(cudaGetLastError());
  ODL_CUDA(cudaMemcpyAsync(count_host, total, sizeof(long long), cudaMemcpyDeviceToHost, s));
  ODL_CUDA(cudaStreamSynchronize(s));
  return 0;
}

extern "C" int odl_gather_rows(odl_model* m, const double* src_dev, int row_len, const int* index_dev_or_null,
                               const long long* picks_host, long long n_pick, double* dst_dev, void* stream) {
  if (!m || !m->on_gpu) return fail(ODL_ENODEVICE, "odl_gather_rows: model is not loaded on a GPU (no CPU fallback exists)");
  if (n_pick < 0 || row_len < 1 || (n_pick > 0 && (!src_dev || !picks_host || !dst_dev))) return fail(ODL_EINVAL, "odl_gather_rows: bad argument");
  if (n_pick == 0) return 0;
  ODL_ON_DEVICE(m);
  cudaStream_t s = (cudaStream_t)stream;
  DevBuf& b = m->scratch[22];
  int rc = b.ensure((size_t)n_pick * sizeof(long long));
  if (rc) return rc;
  ODL_CUDA(cudaMemcpyAsync(b.p, picks_host, (size_t)n_pick * sizeof(long long), cudaMemcpyHostToDevice, s));
  const long long elems = n_pick * row_len;
  odl_gather_rows_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, s>>>(src_dev, row_len, index_dev_or_null,
                                                                       static_cast<const long long*>(b.p), n_pick, dst_dev);
  g_launches.fetch_add(1);
  ODL_CUDA(cudaGetLastError());
  ODL_CUDA(cudaStreamSynchronize(s));      // picks_host may be reused by the caller
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Latin-hypercube sample of the priors on the device (Samplers.py:6-51 `sample_lhs`, Framework.py:589-615): row i of
// column j takes stratum perm_j(i) of n, a uniform point inside it, and the prior's ppf.  perm_j is a keyed bijection of
// [0, n) (4-round Feistel network on the next even power of two, cycle-walking back into range), so every column
// visits every stratum exactly once without a sort; the point inside the stratum is Philox4x32-10 keyed by
// (seed, column, row).  Priors: constant, lognorm(s, loc, scale), norm(loc, scale), uniform(loc, scale).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int odl_mix32(unsigned int x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ unsigned long long odl_feistel_perm(unsigned long long i, unsigned long long n, int half_bits,
                                                              unsigned int key) {
  const unsigned int mask = (half_bits >= 32) ? 0xffffffffu : ((1u << half_bits) - 1u);
  unsigned long long x = i;
  do {                                                            // cycle-walk: the domain is < 4 n
    unsigned int L = (unsigned int)(x >> half_bits) & mask, R = (unsigned int)x & mask;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const unsigned int F = odl_mix32(R ^ (key + 0x9E3779B9u * (unsigned int)(r + 1))) & mask;
      const unsigned int t = L ^ F;
      L = R; R = t;
    }
    x = ((unsigned long long)L << half_bits) | R;
  } while (x >= n);
  return x;
}
__device__ __forceinline__ void odl_philox_host(unsigned int c0, unsigned int c1, unsigned int c2, unsigned int c3,
                                                unsigned int k0, unsigned int k1, unsigned int* out) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned int hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const unsigned int hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const unsigned int n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
struct OdlLhsArgs {
  long long n;
  int n_param, half_bits;
  unsigned long long seed;
  const int* kind;               // [n_param] 0 constant a, 1 lognorm(s=a, loc=b, scale=c), 2 norm(loc=b, scale=c), 3 uniform(loc=b, scale=c)
  const double* a; const double* b; const double* c;
  double* theta;                 // [n][n_param]
};
__global__ void __launch_bounds__(256) odl_lhs_kernel(const OdlLhsArgs A) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < A.n; i += (long long)gridDim.x * blockDim.x) {
    for (int j = 0; j < A.n_param; ++j) {
      const int kind = A.kind[j];
      double v = A.a[j];
      if (kind != 0) {
        const unsigned int key = odl_mix32((unsigned int)A.seed ^ odl_mix32((unsigned int)(A.seed >> 32) + 0x85ebca6bu * (unsigned int)(j + 1)));
        const unsigned long long stratum = odl_feistel_perm((unsigned long long)i, (unsigned long long)A.n, A.half_bits, key);
        unsigned int r[4];
        odl_philox_host((unsigned int)i, (unsigned int)(i >> 32), (unsigned int)j, 0x4c4853u, (unsigned int)A.seed,
                        (unsigned int)(A.seed >> 32), r);
        const unsigned long long bits = (((unsigned long long)r[0]) << 21) ^ ((unsigned long long)r[1] >> 11);
        const double jitter = (double)(bits & ((1ull << 53) - 1)) * 1.1102230246251565e-16;          // [0, 1)
        double u = ((double)stratum + jitter) / (double)A.n;
        u = fmin(fmax(u, 1.1102230246251565e-16), 1.0 - 1.1102230246251565e-16);
        if (kind == 1) v = A.b[j] + A.c[j] * exp(A.a[j] * normcdfinv(u));
        else if (kind == 2) v = A.b[j] + A.c[j] * normcdfinv(u);
        else v = A.b[j] + A.c[j] * u;
      }
      A.theta[i * A.n_param + j] = v;
    }
  }
}

extern "C" int odl_sample_lhs(odl_model* m, long long n, int n_param, const int* kind, const double* a, const double* b,
                              const double* c, unsigned long long seed, double* theta_dev, void* stream) {
  if (!m || !m->on_gpu) return fail(ODL_ENODEVICE, "odl_sample_lhs: model is not loaded on a GPU (no CPU fallback exists)");
  if (n < 0 || n_param < 1 || n_param > 4096 || !kind || !a || !b || !c || (n > 0 && !theta_dev))
    return fail(ODL_EINVAL, "odl_sample_lhs: bad argument");
  for (int j = 0; j < n_param; ++j) if (kind[j] < 0 || kind[j] > 3) return fail(ODL_EINVAL, "odl_sample_lhs: unknown prior kind");
  if (n == 0) return 0;
  ODL_ON_DEVICE(m);
  cudaStream_t s = (cudaStream_t)stream;
  DevBuf& bk = m->scratch[21];
  int rc = bk.ensure((size_t)n_param * (sizeof(int) + 3 * sizeof(double)) + 64);
  if (rc) return rc;
  char* base = static_cast<char*>(bk.p);
  double* da = reinterpret_cast<double*>(base);
  double* db = da + n_param; double* dc = db + n_param;
  int* dk = reinterpret_cast<int*>(dc + n_param);
  ODL_CUDA(cudaMemcpyAsync(da, a, n_param * sizeof(double), cudaMemcpyHostToDevice, s));
  ODL_CUDA(cudaMemcpyAsync(db, b, n_param * sizeof(double), cudaMemcpyHostToDevice, s));
  ODL_CUDA(cudaMemcpyAsync(dc, c, n_param * sizeof(double), cudaMemcpyHostToDevice, s));
  ODL_CUDA(cudaMemcpyAsync(dk, kind, n_param * sizeof(int), cudaMemcpyHostToDevice, s));
  OdlLhsArgs A;
  A.n = n; A.n_param = n_param; A.seed = seed; A.kind = dk; A.a = da; A.b = db; A.c = dc; A.theta = theta_dev;
  int bits = 1;
  while ((1ull << bits) < (unsigned long long)n) ++bits;
  A.half_bits = (bits + 1) / 2;                                   // even number of bits: domain 2^(2*half_bits) < 4 n
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((n + 255) / 256, (long long)m->sm_count * 16));
  odl_lhs_kernel<<<grid, 256, 0, s>>>(A);
  g_launches.fetch_add(1);
  ODL_CUDA(cudaGetLastError());
  ODL_CUDA(cudaStreamSynchronize(s));                              // the host arrays may be reused by the caller
  return 0;
}

// ---------------------------------------------------------------------------------------------
// The reference chain's own random numbers, regenerated on the host (pure CPU code, no GPU needed): numpy's legacy
// RandomState(seed) -- MT19937 seeded with init_genrand, doubles from two 32-bit words, gaussians by the polar method
// with its one-value cache -- consumed exactly as Samplers.MetropolisHastings consumes it per iteration
// (Samplers.py:70, :108, :118-121, :127; Framework.py:103, :119): n_walk normals N(0, step_sd) (the proposal
// increments), n_prior_draws standard normals (the prior `rvs` inside the unused pdf() calls, one each for lognorm /
// norm priors), one uniform.  Feeding these to odl_mcmc (ODL_RNG_HOST_STREAMS) reproduces the reference chain.
// ---------------------------------------------------------------------------------------------
namespace {
struct LegacyMT {
  uint32_t key[624];
  int pos;
  bool has_gauss;
  double gauss;
  explicit LegacyMT(uint32_t seed) : pos(624), has_gauss(false), gauss(0.0) {
    for (int i = 0; i < 624; ++i) { key[i] = seed; seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)i + 1u; }
  }
  void twist() {
    const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, A = 0x9908b0dfu;
    int i;
    uint32_t y;
    for (i = 0; i < 624 - 397; ++i) { y = (key[i] & UPPER) | (key[i + 1] & LOWER); key[i] = key[i + 397] ^ (y >> 1) ^ ((y & 1u) ? A : 0u); }
    for (; i < 623; ++i) { y = (key[i] & UPPER) | (key[i + 1] & LOWER); key[i] = key[i + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? A : 0u); }
    y = (key[623] & UPPER) | (key[0] & LOWER);
    key[623] = key[396] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
    pos = 0;
  }
  uint32_t next32() {
    if (pos == 624) twist();
    uint32_t y = key[pos++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
  }
  double next_double() {
    const int32_t a = (int32_t)(next32() >> 5), b = (int32_t)(next32() >> 6);
    return (a * 67108864.0 + b) / 9007199254740992.0;
  }
  double next_gauss() {
    if (has_gauss) { has_gauss = false; const double t = gauss; gauss = 0.0; return t; }
    double f, x1, x2, r2;
    do {
      x1 = 2.0 * next_double() - 1.0;
      x2 = 2.0 * next_double() - 1.0;
      r2 = x1 * x1 + x2 * x2;
    } while (r2 >= 1.0 || r2 == 0.0);
    f = sqrt(-2.0 * log(r2) / r2);
    gauss = f * x1; has_gauss = true;
    return f * x2;
  }
};
}  // namespace

extern "C" int odl_reference_streams(const unsigned int* seeds, int n_chain, int n_iter, int n_walk, int n_prior_draws,
                                     double step_sd, double* z, double* u) {
  if (n_chain < 0 || n_iter < 0 || n_walk < 0 || n_prior_draws < 0 || (n_chain > 0 && (!seeds || !u || (n_walk > 0 && !z))))
    return fail(ODL_EINVAL, "odl_reference_streams: bad argument");
  auto run = [=](int c0, int c1) {
    for (int c = c0; c < c1; ++c) {
      LegacyMT rs(seeds[c]);
      double* zc = z + (size_t)c * n_iter * n_walk;
      double* uc = u + (size_t)c * n_iter;
      for (int i = 0; i < n_iter; ++i) {
        for (int j = 0; j < n_walk; ++j) zc[(size_t)i * n_walk + j] = 0.0 + step_sd * rs.next_gauss();
        for (int j = 0; j < n_prior_draws; ++j) rs.next_gauss();
        uc[i] = rs.next_double();
      }
    }
  };
  // chains are independent streams: split them over the host cores when there is enough work to pay for threads
  int n_thr = (int)std::thread::hardware_concurrency();
  if (n_thr > 16) n_thr = 16;
  if ((long long)n_chain * n_iter < 200000 || n_thr < 2) { run(0, n_chain); return 0; }
  if (n_thr > n_chain) n_thr = n_chain;
  std::vector<std::thread> pool;
  for (int t = 0; t < n_thr; ++t)
    pool.emplace_back(run, (int)((long long)n_chain * t / n_thr), (int)((long long)n_chain * (t + 1) / n_thr));
  for (auto& th : pool) th.join();
  return 0;
}

// The same streams generated ON THE DEVICE, for runs whose streams would not fit the host side (4096 chains x 10,000
// iterations x 11 gaussians are 1.2e9 MT19937 words: ~10 s on 16 host cores, and the chains themselves take 0.2 s): one
// thread per chain, its 624-word key array in global memory laid out [624][n_chain] (the lanes of a warp read and
// twist neighbouring words together), position and the polar method's cached value in registers.  Word for word the
// generator above.  Uniforms are bit-identical to numpy's; a gaussian goes through log(), where CUDA's and glibc's
// (both within 1 ulp of the truth) may round differently: z agrees with numpy to 1 ulp, most values exactly.
__device__ __forceinline__ unsigned int odl_mt_next(unsigned int* key, long long C, int& pos) {
  if (pos == 624) {
    const unsigned int UPPER = 0x80000000u, LOWER = 0x7fffffffu, A = 0x9908b0dfu;
    unsigned int cur = key[0], y;                     // cur / nxt: words i, i + 1 as they were BEFORE this twist
    for (int i = 0; i < 623; ++i) {
      const unsigned int nxt = key[(long long)(i + 1) * C];
      y = (cur & UPPER) | (nxt & LOWER);
      const int j = (i < 624 - 397) ? i + 397 : i + (397 - 624);
      key[(long long)i * C] = key[(long long)j * C] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
      cur = nxt;
    }
    y = (cur & UPPER) | (key[0] & LOWER);             // word 0 as the loop has just rewritten it (genrand's order)
    key[623LL * C] = key[396LL * C] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
    pos = 0;
  }
  unsigned int y = key[(long long)pos * C];
  ++pos;
  y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
  return y;
}
__device__ __forceinline__ double odl_mt_double(unsigned int* key, long long C, int& pos) {
  const int a = (int)(odl_mt_next(key, C, pos) >> 5), b = (int)(odl_mt_next(key, C, pos) >> 6);
  return __ddiv_rn(__dadd_rn(__dmul_rn((double)a, 67108864.0), (double)b), 9007199254740992.0);
}
__global__ void __launch_bounds__(128) odl_refstream_kernel(const unsigned int* seeds, unsigned int* keys, long long C, int n_iter,
                                                            int n_walk, int n_prior, double step_sd, double* z, double* u) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  unsigned int* key = keys + c;
  unsigned int seed = seeds[c];
  for (int i = 0; i < 624; ++i) { key[(long long)i * C] = seed; seed = 1812433253u * (seed ^ (seed >> 30)) + (unsigned int)i + 1u; }
  int pos = 624;
  bool has_gauss = false;
  double cached = 0.0;
  auto gauss = [&]() -> double {
    if (has_gauss) { has_gauss = false; const double t = cached; cached = 0.0; return t; }
    double x1, x2, r2;
    do {
      x1 = __dadd_rn(__dmul_rn(2.0, odl_mt_double(key, C, pos)), -1.0);
      x2 = __dadd_rn(__dmul_rn(2.0, odl_mt_double(key, C, pos)), -1.0);
      r2 = __dadd_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2));     // no contraction: numpy's build does not fuse these
    } while (r2 >= 1.0 || r2 == 0.0);
    const double f = __dsqrt_rn(__ddiv_rn(__dmul_rn(-2.0, log(r2)), r2));
    cached = __dmul_rn(f, x1); has_gauss = true;
    return __dmul_rn(f, x2);
  };
  double* zc = z + c * (long long)n_iter * n_walk;
  double* uc = u + c * (long long)n_iter;
  for (int i = 0; i < n_iter; ++i) {
    for (int j = 0; j < n_walk; ++j) zc[(long long)i * n_walk + j] = __dadd_rn(0.0, __dmul_rn(step_sd, gauss()));
    for (int j = 0; j < n_prior; ++j) gauss();
    uc[i] = odl_mt_double(key, C, pos);
  }
}

extern "C" int odl_reference_streams_device(odl_model* m, const unsigned int* seeds_host, int n_chain, int n_iter, int n_walk,
                                            int n_prior_draws, double step_sd, double* z_dev, double* u_dev, void* stream) {
  if (!m || !m->on_gpu) return fail(ODL_ENODEVICE, "odl_reference_streams_device: model is not loaded on a GPU (odl_reference_streams is the host generator)");
  if (n_chain < 0 || n_iter < 0 || n_walk < 0 || n_prior_draws < 0 || (n_chain > 0 && (!seeds_host || !u_dev || (n_walk > 0 && !z_dev))))
    return fail(ODL_EINVAL, "odl_reference_streams_device: bad argument");
  if (n_chain == 0 || n_iter == 0) return 0;
  ODL_ON_DEVICE(m);
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
  if ((rc = serialize_after_previous_call(m, s))) return rc;
  if ((rc = m->mt_state.ensure(((size_t)624 + 1) * n_chain * sizeof(unsigned int)))) return rc;
  unsigned int* keys = static_cast<unsigned int*>(m->mt_state.p);
  unsigned int* seeds = keys + (size_t)624 * n_chain;
  ODL_CUDA(cudaMemcpyAsync(seeds, seeds_host, (size_t)n_chain * sizeof(unsigned int), cudaMemcpyHostToDevice, s));
  ODL_CUDA(cudaEventRecord(m->ev0, s));
  odl_refstream_kernel<<<(unsigned)((n_chain + 127) / 128), 128, 0, s>>>(seeds, keys, n_chain, n_iter, n_walk, n_prior_draws, step_sd, z_dev, u_dev);
  g_launches.fetch_add(1);
  ODL_CUDA(cudaGetLastError());
  ODL_CUDA(cudaEventRecord(m->ev1, s));
  m->timed = true; m->n_pass = 1;
  ODL_CUDA(cudaStreamSynchronize(s));          // seeds_host may be reused by the caller
  return 0;
}

// ---------------------------------------------------------------------------------------------
// The one collective of the path: Gelman-Rubin R-hat over the chains of every rank.  The reference
// has no counterpart (its gather is pd.concat of the workers' frames, Framework.py:1035-1038, and it has no R-hat);
// here the per-chain Welford summaries (count, mean[P], M2[P] of ln theta, written by the MCMC kernel) are all-gathered
// with ncclAllGather over NVLink / NVSwitch and reduced on the device: per parameter W = mean_j s_j^2,
// B = n var_j(mean_j) (ddof 1), R-hat = sqrt(((n-1)/n W + B/n) / W), plus the pooled count / mean / M2 of all kept rows
// (what the fitting report's rawstats needs).  NCCL is loaded with dlopen at the first use: no link-time dependency,
// the library still loads (and computes on one GPU) where NCCL is absent.
// ---------------------------------------------------------------------------------------------
struct Nccl {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
static Nccl g_nccl;
static int load_nccl() {
  if (g_nccl.ok) return 0;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);       // a copy torch has loaded already is reused (same SONAME)
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail(ODL_ENODEVICE, std::string("NCCL is not available: ") + dlerror());
  auto sym = [&](const char* n) { return dlsym(h, n); };
  g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(sym("ncclGetUniqueId"));
  g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(sym("ncclCommInitRank"));
  g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(sym("ncclCommDestroy"));
  g_nccl.AllGather = reinterpret_cast<decltype(g_nccl.AllGather)>(sym("ncclAllGather"));
  g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(sym("ncclAllReduce"));
  g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(sym("ncclGetErrorString"));
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllGather || !g_nccl.AllReduce || !g_nccl.GetErrorString)
    return fail(ODL_ENODEVICE, "NCCL library lacks an expected entry point");
  g_nccl.ok = true;
  return 0;
}
#define ODL_NCCL(call)                                                                                      \
  do {                                                                                                      \
    ncclResult_t r_ = (call);                                                                               \
    if (r_ != ncclSuccess) return fail(ODL_ECUDA, std::string(#call) + ": " + g_nccl.GetErrorString(r_));   \
  } while (0)

extern "C" int odl_comm_unique_id(unsigned char* id128) {
  if (!id128) return fail(ODL_EINVAL, "odl_comm_unique_id: null argument");
  int rc = load_nccl();
  if (rc) return rc;
  ncclUniqueId id;
  ODL_NCCL(g_nccl.GetUniqueId(&id));
  static_assert(sizeof(id) == ODL_COMM_ID_BYTES, "ncclUniqueId size");
  memcpy(id128, &id, sizeof id);
  return 0;
}

extern "C" int odl_comm_init(odl_model* m, const unsigned char* id128, int world, int rank) {
  if (!m || !m->on_gpu) return fail(ODL_ENODEVICE, "odl_comm_init: model is not loaded on a GPU");
  if (!id128 || world < 1 || rank < 0 || rank >= world) return fail(ODL_EINVAL, "odl_comm_init: bad argument");
  odl_comm_destroy(m);
  m->comm_world = world; m->comm_rank = rank;
  if (world == 1) return 0;                                       // nothing to gather from
  int rc = load_nccl();
  if (rc) return rc;
  ODL_ON_DEVICE(m);
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  ODL_NCCL(g_nccl.CommInitRank(&m->comm, world, id, rank));
  return 0;
}

extern "C" int odl_comm_destroy(odl_model* m) {
  if (m && m->comm && g_nccl.ok) { g_nccl.CommDestroy(m->comm); }
  if (m) { m->comm = nullptr; m->comm_world = 1; m->comm_rank = 0; }
  return 0;
}

// one CTA per parameter; chains with count == 0 (padding of ragged shards) are skipped.  out[q] = R-hat,
// pooled[0] = N, pooled[1+q] = mean, pooled[1+P+q] = M2 over all kept rows of all chains.
__device__ __forceinline__ double odl_block_sum(double v, double* sh) {
  const int t = threadIdx.x;
  sh[t] = v;
  __syncthreads();
  for (int off = blockDim.x >> 1; off > 0; off >>= 1) {            // fixed tree: the same result on every rank, every run
    if (t < off) sh[t] += sh[t + off];
    __syncthreads();
  }
  const double r = sh[0];
  __syncthreads();
  return r;
}
__global__ void __launch_bounds__(256) odl_rhat_kernel(const double* summ, long long n_chain, int P, double* rhat, double* pooled) {
  __shared__ double sh[256];
  const int q = blockIdx.x;
  const int W = 1 + 2 * P;
  double m = 0.0, n_first = 0.0, N = 0.0, sm = 0.0, sw = 0.0, nsm = 0.0;
  for (long long c = threadIdx.x; c < n_chain; c += blockDim.x) {
    const double n = summ[c * W];
    if (n > 0.0) {
      m += 1.0; N += n;
      if (n_first == 0.0) n_first = n;
      sm += summ[c * W + 1 + q];
      nsm += n * summ[c * W + 1 + q];
      sw += summ[c * W + 1 + P + q] / (n - 1.0);                  // s_j^2
    }
  }
  const double M = odl_block_sum(m, sh), Nall = odl_block_sum(N, sh);
  const double mean_of_means = odl_block_sum(sm, sh) / M;
  const double pooled_mean = odl_block_sum(nsm, sh) / Nall;
  const double Wv = odl_block_sum(sw, sh) / M;
  // n: the reference keeps the same number of rows in every chain (nits - 1 - burnin); take the largest count seen
  sh[threadIdx.x] = n_first;
  __syncthreads();
  for (int off = blockDim.x >> 1; off > 0; off >>= 1) { if (threadIdx.x < off) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + off]); __syncthreads(); }
  const double n = sh[0];
  __syncthreads();
  double sb = 0.0, m2 = 0.0;
  for (long long c = threadIdx.x; c < n_chain; c += blockDim.x) {
    const double nc = summ[c * W];
    if (nc > 0.0) {
      const double mu = summ[c * W + 1 + q];
      sb += (mu - mean_of_means) * (mu - mean_of_means);
      m2 += summ[c * W + 1 + P + q] + nc * (mu - pooled_mean) * (mu - pooled_mean);
    }
  }
  const double B = n * odl_block_sum(sb, sh) / (M - 1.0);
  const double M2 = odl_block_sum(m2, sh);
  if (threadIdx.x == 0) {
    rhat[q] = sqrt(((n - 1.0) / n * Wv + B / n) / Wv);
    pooled[1 + q] = pooled_mean; pooled[1 + P + q] = M2;
    if (q == 0) pooled[0] = Nall;
  }
}

extern "C" int odl_rhat(odl_model* m, const double* summaries, int n_chain_local, int n_param, int mem, double* rhat_host,
                        double* pooled_host_or_null, long long* n_chain_total_or_null, void* stream) {
  if (!m || !m->on_gpu) return fail(ODL_ENODEVICE, "odl_rhat: model is not loaded on a GPU (no CPU fallback exists)");
  if (n_chain_local < 0 || n_param < 1 || !rhat_host || (n_chain_local > 0 && !summaries)) return fail(ODL_EINVAL, "odl_rhat: bad argument");
  ODL_ON_DEVICE(m);
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
  if ((rc = serialize_after_previous_call(m, s))) return rc;
  ODL_CUDA(cudaEventRecord(m->ev0, s));
  const size_t W = 1 + 2 * (size_t)n_param;
  // ODL_RHAT_LOCAL: this rank's chains only, whatever communicator the handle has joined -- a single-GPU caller (the
  // facade without distributed=True) on a handle that other code has joined to a communicator must not start a collective
  // the other ranks never enter (bench.py under torchrun hung exactly there)
  const bool local_only = (mem & ODL_RHAT_LOCAL) != 0;
  mem &= ~ODL_RHAT_LOCAL;
  const int world = (m->comm && !local_only) ? m->comm_world : 1;
  // scratch: [0] padded local block, [1] gathered table, [2] results (rhat[P], pooled[1+2P], n_pad)
  DevBuf &bloc = m->scratch[0], &ball = m->scratch[1], &bres = m->scratch[2];
  if ((rc = bres.ensure((n_param + W + 2) * sizeof(double)))) return rc;
  double* d_rhat = static_cast<double*>(bres.p);
  double* d_pooled = d_rhat + n_param;
  int* d_npad = reinterpret_cast<int*>(d_pooled + W);
  int n_pad = n_chain_local;
  if (world > 1) {
    // shards may differ in length: everybody sends the largest shard's row count, short shards pad with count-0 rows
    ODL_CUDA(cudaMemcpyAsync(d_npad, &n_chain_local, sizeof(int), cudaMemcpyHostToDevice, s));
    ODL_NCCL(g_nccl.AllReduce(d_npad, d_npad, 1, ncclInt32, ncclMax, m->comm, s));
    ODL_CUDA(cudaMemcpyAsync(&n_pad, d_npad, sizeof(int), cudaMemcpyDeviceToHost, s));
    ODL_CUDA(cudaStreamSynchronize(s));
  }
  const double* table = summaries;
  long long n_total = n_chain_local;
  if (world > 1 || mem == ODL_MEM_HOST) {
    if ((rc = bloc.ensure(std::max<size_t>(1, (size_t)n_pad) * W * sizeof(double)))) return rc;
    ODL_CUDA(cudaMemsetAsync(bloc.p, 0, std::max<size_t>(1, (size_t)n_pad) * W * sizeof(double), s));
    if (n_chain_local > 0)
      ODL_CUDA(cudaMemcpyAsync(bloc.p, summaries, (size_t)n_chain_local * W * sizeof(double),
                               mem == ODL_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s));
    table = static_cast<const double*>(bloc.p);
  }
  if (world > 1) {
    if ((rc = ball.ensure((size_t)world * n_pad * W * sizeof(double)))) return rc;
    ODL_NCCL(g_nccl.AllGather(bloc.p, ball.p, (size_t)n_pad * W, ncclFloat64, m->comm, s));
    table = static_cast<const double*>(ball.p);
    n_total = (long long)world * n_pad;
  }
  if (n_total < 1) return fail(ODL_EINVAL, "odl_rhat: no chains");
  odl_rhat_kernel<<<n_param, 256, 0, s>>>(table, n_total, n_param, d_rhat, d_pooled);
  g_launches.fetch_add(1);
  ODL_CUDA(cudaGetLastError());
  ODL_CUDA(cudaMemcpyAsync(rhat_host, d_rhat, n_param * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (pooled_host_or_null) ODL_CUDA(cudaMemcpyAsync(pooled_host_or_null, d_pooled, W * sizeof(double), cudaMemcpyDeviceToHost, s));
  ODL_CUDA(cudaEventRecord(m->ev1, s));
  m->timed = true; m->n_pass = 1;
  ODL_CUDA(cudaStreamSynchronize(s));
  if (n_chain_total_or_null) *n_chain_total_or_null = n_total;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// FP64 roofline denominator: 8 independent DFMA chains per thread, no memory traffic
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) odl_dfma_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // never true; keeps the chains alive
}

extern "C" int odl_fp64_peak(int device, int repeats, double* tflops, float* ms_per_launch) {
  if (!tflops) return fail(ODL_EINVAL, "null argument");
  int prev = -1;
  if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
  struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{device >= 0 && device != prev ? prev : -1};
  if (device >= 0) ODL_CUDA(cudaSetDevice(device));
  ODL_CUDA(cudaFree(0));
  int dev = 0;
  ODL_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  ODL_CUDA(cudaGetDeviceProperties(&prop, dev));
  const int block = 256, per_sm = 8, iters = 4096;
  const int grid = prop.multiProcessorCount * per_sm;
  double* d = nullptr;
  ODL_CUDA(cudaMalloc(&d, (size_t)grid * block * sizeof(double)));
  cudaEvent_t e0, e1;
  ODL_CUDA(cudaEventCreate(&e0));
  ODL_CUDA(cudaEventCreate(&e1));
  if (repeats < 1) repeats = 5;
  for (int w = 0; w < 2; ++w) odl_dfma_peak_kernel<<<grid, block>>>(d, iters, 0.999999, 1e-9);
  float best = 1e30f;
  for (int r = 0; r < repeats; ++r) {
    ODL_CUDA(cudaEventRecord(e0));
    odl_dfma_peak_kernel<<<grid, block>>>(d, iters, 0.999999, 1e-9);
    g_launches.fetch_add(1);
    ODL_CUDA(cudaEventRecord(e1));
    ODL_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    ODL_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    best = std::min(best, ms);
  }
  ODL_CUDA(cudaGetLastError());
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  const double flops = 2.0 * 64.0 * (double)iters * (double)grid * block;   // 8 chains x 8 unroll x FMA
  *tflops = flops / (best * 1e-3) / 1e12;
  if (ms_per_launch) *ms_per_launch = best;
  return 0;
}
