// C ABI (include/ddp_b200.h) over the kernels in kernels.cuh: arena carving, the
// per-iteration launch sequence, host<->device array access.  No CPU compute path exists
// here: every phase is a CUDA launch, and errors surface as negative status codes.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/ddp_b200.h"
#include "kernels.cuh"
#include "backward_sym.cuh"
#include "quadruped_fused.cuh"
#include "quadruped_rollout.cuh"
#include "arm_rollout.cuh"
#include "quadruped_quat_fused.cuh"

using namespace ddp;

static thread_local std::string g_err;
#define CK(call)                                                                      \
  do {                                                                                \
    cudaError_t e_ = (call);                                                          \
    if (e_ != cudaSuccess) {                                                          \
      g_err = std::string(#call) + ": " + cudaGetErrorString(e_);                     \
      return DDP_ERR_CUDA;                                                            \
    }                                                                                 \
  } while (0)

struct ddp_solver {
  int model, np;
  int device;            // ordinal of the GPU the arena lives on; every entry point runs under it
  Dev d;
  cudaStream_t stream;
  // per-solver (hence per-device) launch configuration, set on first use
  bool cfg_bwd, cfg_bwd_sym;
  int fused_ctas;
  std::vector<double> eps_host;
  double beta;
  int* h_counters;  // pinned [2]
  cudaEvent_t ev[4];
  float ms[4];
  long long launches;
  bool timings_valid;
  bool scalar_backward;  // debug: force the scalar shared-memory kernel for n >= 16
  bool small_backward;   // n <= 4, m = 1: thread-per-trajectory register kernel (DDP_SMALL_BACKWARD=cta: the CTA kernel)
  bool quad_rollout8;    // 8-lane quadruped rollout (DDP_QUAD_ROLLOUT=generic selects rollout_kernel)
  bool quad_fused;       // fused structured quadruped linearization (DDP_QUAD_LINEARIZE=fused|ad)
  bool arm_rollout8;     // 8-lane arm + ball rollout (DDP_ARM_ROLLOUT=generic selects rollout_kernel)
  int quad_sub;          // substeps of the quadruped model (the fused linearization needs 2)
  int sms;               // SMs of the solver's device (0: not queried yet)
  // array table
  double* darr[16];
  size_t dsize[16];
};

namespace {

// Every entry point that takes a solver runs with the solver's device current and restores the
// caller's device on return: the arena, the stream and the cudaFuncSetAttribute opt-ins are all
// per device, and the caller may have another GPU current.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = (cudaSetDevice(dev) == cudaSuccess);
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};
#define GUARD(s) DeviceGuard guard_((s)->device)

struct Carver {
  char* base;
  size_t off;
  template <class T>
  T* take(size_t count) {
    off = (off + 255) & ~(size_t)255;
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    // 16 bytes of slack: the bulk copies of odd-sized fx / fu tiles are rounded out to 16-byte
    // boundaries and may read 8 bytes past the last tile of an array (backward_sym.cuh)
    off += count * sizeof(T) + 16;
    return p;
  }
};

constexpr int kMaxEps = 2048;

void carve(Dev& d, double** params, int np, Carver& c) {
  const size_t B = d.B, N = d.N, T = d.T, n = d.n, m = d.m, A = d.A;
  *params = c.take<double>(np);
  d.Q = c.take<double>(n * n);
  d.R = c.take<double>(m * m);
  d.Qf = c.take<double>(n * n);
  d.x_nom = c.take<double>(B * n);
  d.x0 = c.take<double>(B * n);
  d.eps_table = c.take<double>(kMaxEps);
  d.x_bar = c.take<double>(B * N * n);
  d.u_bar = c.take<double>(B * T * m);
  d.K = c.take<double>(B * T * m * n);
  d.kappa = c.take<double>(B * T * m);
  d.dV = c.take<double>(B * T);
  d.fx = c.take<double>(B * T * n * n);
  d.fu = c.take<double>(B * T * n * m);
  d.xc = c.take<double>(B * A * N * n);
  d.uc = c.take<double>(B * A * T * m);
  d.Lc = c.take<double>(B * A);
  d.Ec = c.take<double>(B * A);
  d.L = c.take<double>(B);
  d.L_new = c.take<double>(B);
  d.eps = c.take<double>(B);
  d.improvement = c.take<double>(B);
  d.ls_iters = c.take<int>(B);
  d.status = c.take<int>(B);
  d.active = c.take<int>(B);
  d.resolved = c.take<int>(B);
  d.acc = c.take<int>(B);
  d.iters = c.take<int>(B);
  d.rearm = c.take<int>(B);
  d.rearm_mark = c.take<int>(B);
  d.resolves = c.take<int>(B);
  d.L_conv = c.take<double>(B);
  d.mpc_target_adv_buf = c.take<double>(n);
  d.u_lim_buf = c.take<double>(2 * m);
  d.active_save = c.take<int>(B);
  d.status_save = c.take<int>(B);
  d.counters = c.take<int>(4);
  d.sm_slots = c.take<int>(1024);
  d.unres = c.take<int>(B);
  d.kplist = c.take<int>(B * T);
  d.kpcount = c.take<int>(B);
  d.seg_s = c.take<int>(B * T);
  d.seg_e = c.take<int>(B * T);
  d.flag = c.take<unsigned char>(B * N);
  d.done = c.take<unsigned char>(B * N);
  d.segs[0] = c.take<int>(B * 2 * T);
  d.segs[1] = c.take<int>(B * 2 * T);
  d.nseg[0] = c.take<int>(B);
  d.nseg[1] = c.take<int>(B);
  d.evallist = c.take<int>(B * T);
  d.evalcount = c.take<int>(B);
}

int model_dims(int model_id, int* n, int* m, int* np) {
  DDP_MODEL_SWITCH(model_id, { *n = Model::n; *m = Model::m; *np = Model::np; });
  return 0;
}

inline int cdiv(size_t a, size_t b) { return (int)((a + b - 1) / b); }

int device_sms(ddp_solver* s) {
  if (!s->sms) cudaDeviceGetAttribute(&s->sms, cudaDevAttrMultiProcessorCount, s->device);
  return s->sms > 0 ? s->sms : 1;
}

// ---- templated launchers ---------------------------------------------------------------
template <class Model>
int launch_rollout(ddp_solver* s, int ls_base, int per_traj, int n_items) {
  constexpr int G = Cfg<Model>::G_ROLL;
  const size_t threads = (size_t)n_items * G;
  rollout_kernel<Model, G><<<cdiv(threads, 128), 128, 0, s->stream>>>(s->d, ls_base, per_traj, n_items);
  s->launches++;
  return 0;
}
template <class Model>
int launch_linearize(ddp_solver* s, const int* list, const int* count) {
  constexpr int G = Cfg<Model>::G_LIN, K = Cfg<Model>::K_LIN, P = Cfg<Model>::P_LIN;
  const size_t threads = (size_t)s->d.B * s->d.T * G;
  linearize_kernel<Model, G, K, P><<<cdiv(threads, 128), 128, 0, s->stream>>>(s->d, list, count);
  s->launches++;
  return 0;
}
template <class Model>
int launch_backward_sym(ddp_solver* s) {
  typedef BsCfg<Model::n, Model::m> C;
  const size_t smem = sizeof(BsSmem<Model::n, Model::m>);
  if (!s->cfg_bwd_sym) {
    cudaError_t e = cudaFuncSetAttribute(backward_sym_kernel<Model>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      g_err = std::string("cudaFuncSetAttribute(backward_sym): ") + cudaGetErrorString(e);
      return DDP_ERR_CUDA;
    }
    s->cfg_bwd_sym = true;
  }
  backward_sym_kernel<Model><<<s->d.B, C::NT, smem, s->stream>>>(s->d);
  s->launches++;
  return 0;
}
template <class Model>
int launch_backward(ddp_solver* s) {
  if constexpr (Model::n >= 16) {
    if (!s->scalar_backward) return launch_backward_sym<Model>(s);
  }
  if constexpr (Model::n <= 4 && Model::m == 1) {
    // thread per trajectory, everything in registers: wins for a handful of trajectories (the drop-in
    // class: B = 1; pendulum 129 -> 63 us, cart-pole 243 -> 180 us per sweep); at B = 50 its per-thread
    // tile loads are 32 uncoalesced streams per warp and the CTA kernel is faster (455 vs 503 us)
    if (s->small_backward && s->d.B <= 8) {
      backward_small_kernel<Model><<<cdiv(s->d.B, 32), 32, 0, s->stream>>>(s->d);
      s->launches++;
      return 0;
    }
  }
  constexpr int NT = Cfg<Model>::BWD_THREADS;
  const size_t smem = sizeof(BwdSmem<Model::n, Model::m>);
  if (!s->cfg_bwd) {
    cudaError_t e = cudaFuncSetAttribute(backward_kernel<Model, NT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      g_err = std::string("cudaFuncSetAttribute(backward): ") + cudaGetErrorString(e);
      return DDP_ERR_CUDA;
    }
    s->cfg_bwd = true;
  }
  backward_kernel<Model, NT><<<s->d.B, NT, smem, s->stream>>>(s->d);
  s->launches++;
  return 0;
}

template <bool QUAT>
void launch_rollout_quad8(ddp_solver* s, int ls_base, int per_traj, int n_items) {
  const bool shared = ls_base == 0 && per_traj == kRqCands && n_items == s->d.B * kRqCands;
  const int ctas = shared ? s->d.B : cdiv(n_items, kRqCands);
  const bool few = ctas <= 4 * device_sms(s);   // the high-register build: four CTAs (8 warps x 255 registers) per SM
  const int nt = kRqLanes * kRqCands;
  if (shared) {
    if (few) rollout_quad8_kernel<true, QUAT, 3><<<ctas, nt, 0, s->stream>>>(s->d, ls_base, per_traj, n_items);
    else rollout_quad8_kernel<true, QUAT, 7><<<ctas, nt, 0, s->stream>>>(s->d, ls_base, per_traj, n_items);
  } else {
    if (few) rollout_quad8_kernel<false, QUAT, 3><<<ctas, nt, 0, s->stream>>>(s->d, ls_base, per_traj, n_items);
    else rollout_quad8_kernel<false, QUAT, 7><<<ctas, nt, 0, s->stream>>>(s->d, ls_base, per_traj, n_items);
  }
  s->launches++;
}
int do_rollout(ddp_solver* s, int ls_base, int per_traj, int n_items) {
  if (s->quad_rollout8 && s->model == MODEL_QUADRUPED) {
    launch_rollout_quad8<false>(s, ls_base, per_traj, n_items);
    return 0;
  }
  if (s->quad_rollout8 && s->model == MODEL_QUADRUPED_QUAT) {   // the reference's n = 37 layout
    launch_rollout_quad8<true>(s, ls_base, per_traj, n_items);
    return 0;
  }
  if (s->arm_rollout8 && s->model == MODEL_ARM_BALL && s->d.diag_cost) {
    const int ctas = cdiv(n_items, kRaCands);
    if (ctas <= 2 * device_sms(s))
      rollout_arm8_kernel<2><<<ctas, kRaLanes * kRaCands, 0, s->stream>>>(s->d, ls_base, per_traj, n_items);
    else
      rollout_arm8_kernel<4><<<ctas, kRaLanes * kRaCands, 0, s->stream>>>(s->d, ls_base, per_traj, n_items);
    s->launches++;
    return 0;
  }
  DDP_MODEL_SWITCH(s->model, return launch_rollout<Model>(s, ls_base, per_traj, n_items));
  return 0;
}
int launch_quad_fused(ddp_solver* s, const int* list, const int* count) {
  const size_t smem = sizeof(QfWarpSmem) * kQfWarps;
  if (!s->fused_ctas) {
    cudaError_t e = cudaFuncSetAttribute(quad_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      g_err = std::string("cudaFuncSetAttribute(quad_fused): ") + cudaGetErrorString(e);
      return DDP_ERR_CUDA;
    }
    int sms = 0, per_sm = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, quad_fused_kernel, kQfWarps * 32, smem);
    s->fused_ctas = sms * (per_sm > 0 ? per_sm : 1);
  }
  const int n_items = s->d.B * s->d.T;
  const int grid = std::min(s->fused_ctas, cdiv(n_items, kQfWarps));
  quad_fused_kernel<<<grid, kQfWarps * 32, smem, s->stream>>>(s->d, list, count, n_items);
  s->launches++;
  return 0;
}
int launch_quad_quat_fused(ddp_solver* s, const int* list, const int* count) {
  const size_t smem = sizeof(QqWarpSmem) * kQfWarps;
  if (!s->fused_ctas) {
    cudaError_t e = cudaFuncSetAttribute(quad_quat_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      g_err = std::string("cudaFuncSetAttribute(quad_quat_fused): ") + cudaGetErrorString(e);
      return DDP_ERR_CUDA;
    }
    int sms = 0, per_sm = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, quad_quat_fused_kernel, kQfWarps * 32, smem);
    s->fused_ctas = sms * (per_sm > 0 ? per_sm : 1);
  }
  const int n_items = s->d.B * s->d.T;
  const int grid = std::min(s->fused_ctas, cdiv(n_items, kQfWarps));
  quad_quat_fused_kernel<<<grid, kQfWarps * 32, smem, s->stream>>>(s->d, list, count, n_items);
  s->launches++;
  return 0;
}
int do_linearize(ddp_solver* s, const int* list, const int* count) {
  if (s->model == MODEL_QUADRUPED && s->quad_sub == 2 && s->quad_fused) return launch_quad_fused(s, list, count);
  if (s->model == MODEL_QUADRUPED_QUAT && s->quad_sub == 2 && s->quad_fused) return launch_quad_quat_fused(s, list, count);
  DDP_MODEL_SWITCH(s->model, return launch_linearize<Model>(s, list, count));
  return 0;
}
int do_backward(ddp_solver* s) {
  DDP_MODEL_SWITCH(s->model, return launch_backward<Model>(s));
  return 0;
}

#define LAUNCH1(kernel, ...)                                                  \
  do {                                                                        \
    kernel<<<cdiv(s->d.B, 128), 128, 0, s->stream>>>(__VA_ARGS__);            \
    s->launches++;                                                            \
  } while (0)

// _linesearch (ilqr.py:274-337) + commit (:375-376).  Round 0 evaluates the first A candidates of
// every trajectory.  With sync_rounds, the trajectories that accepted none are compacted and
// later rounds spread the whole candidate buffer (B*A slots) over them, so the usual case is
// one or two rounds; each later round costs one small D2H counter read.
int phase_linesearch(ddp_solver* s, bool sync_rounds) {
  Dev& d = s->d;
  const int slots = d.B * d.A;
  int ls_base = 0, per_traj = d.A, n_traj = d.B;
  while (true) {
    int rc = do_rollout(s, ls_base, per_traj, n_traj * per_traj);
    if (rc) return rc;
    pick_kernel<<<cdiv(n_traj, 128), 128, 0, s->stream>>>(d, ls_base, per_traj, n_traj);
    s->launches++;
    {
      dim3 grid(d.B, cdiv((size_t)d.N * d.n + (size_t)d.T * d.m, 256 * 4));
      commit_kernel<<<grid, 256, 0, s->stream>>>(d);
      s->launches++;
    }
    LAUNCH1(commit_done_kernel, d);
    ls_base += per_traj;
    if (!sync_rounds || ls_base >= d.n_eps) break;
    CK(cudaMemsetAsync(d.counters, 0, sizeof(int), s->stream));
    LAUNCH1(unresolved_kernel, d);
    CK(cudaMemcpyAsync(s->h_counters, d.counters, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    n_traj = s->h_counters[0];
    if (n_traj == 0) break;
    per_traj = slots / n_traj;
    if (per_traj > d.n_eps - ls_base) per_traj = d.n_eps - ls_base;
  }
  return 0;
}

// _get_derivatives (ilqr.py:380-415)
int phase_derivatives(ddp_solver* s) {
  Dev& d = s->d;
  int rc = 0;
  if (d.kp_method == DDP_KP_SET_INTERVAL) {
    kp_set_interval_kernel<<<d.B, 128, 0, s->stream>>>(d);
    s->launches++;
  } else if (d.kp_method == DDP_KP_ADAPTIVE_JERK) {
    if (d.N > 3) {
      jerk_flag_kernel<<<cdiv((size_t)d.B * (d.N - 3), 128), 128, 0, s->stream>>>(d);
      s->launches++;
    }
    LAUNCH1(jerk_scan_kernel, d);
  } else if (d.kp_method == DDP_KP_ITERATIVE_ERROR) {
    LAUNCH1(ie_init_kernel, d);
    int levels = 2;
    for (int len = d.N; len > 1; len >>= 1) levels++;
    int cur = 0;
    for (int l = 0; l < levels; ++l) {
      LAUNCH1(ie_collect_kernel, d, cur);
      rc = do_linearize(s, d.evallist, d.evalcount);
      if (rc) return rc;
      ie_check_kernel<<<d.B, 128, 0, s->stream>>>(d, cur);
      s->launches++;
      cur ^= 1;
    }
    LAUNCH1(ie_finish_kernel, d);
  } else {
    g_err = "unknown interpolation method";
    return DDP_ERR_ARG;
  }
  if (d.kp_method != DDP_KP_ITERATIVE_ERROR) {
    rc = do_linearize(s, d.kplist, d.kpcount);
    if (rc) return rc;
  }
  if (!(d.kp_method == DDP_KP_SET_INTERVAL && d.minN == 1)) {
    segments_kernel<<<d.B, 128, 0, s->stream>>>(d);
    s->launches++;
    interp_kernel<<<(unsigned)((size_t)d.B * d.T), 128, 0, s->stream>>>(d);
    s->launches++;
  }
  return 0;
}

int iterate_linesearch_impl(ddp_solver* s, bool sync, bool force_all) {
  Dev& d = s->d;
  LAUNCH1(begin_iter_kernel, d, force_all ? 1 : 0);
  CK(cudaEventRecord(s->ev[0], s->stream));
  int rc = phase_linesearch(s, sync);
  if (rc) return rc;
  CK(cudaEventRecord(s->ev[1], s->stream));
  return 0;
}
int iterate_finish_impl(ddp_solver* s) {
  Dev& d = s->d;
  int rc = phase_derivatives(s);
  if (rc) return rc;
  CK(cudaEventRecord(s->ev[2], s->stream));
  rc = do_backward(s);
  if (rc) return rc;
  CK(cudaEventRecord(s->ev[3], s->stream));
  LAUNCH1(finish_iter_kernel, d);
  if (d.mpc_replan > 0) {
    mpc_rearm_kernel<<<d.B, 256, 0, s->stream>>>(d);
    s->launches++;
  }
  s->timings_valid = true;
  CK(cudaGetLastError());
  return 0;
}
int iterate_impl(ddp_solver* s, bool sync, bool force_all) {
  int rc = iterate_linesearch_impl(s, sync, force_all);
  if (rc) return rc;
  return iterate_finish_impl(s);
}

struct ArrInfo {
  void* ptr;
  size_t elems;
};
ArrInfo arr(ddp_solver* s, int which) {
  Dev& d = s->d;
  const size_t B = d.B, N = d.N, T = d.T, n = d.n, m = d.m, A = d.A;
  switch (which) {
    case DDP_X_BAR: return {d.x_bar, B * N * n};
    case DDP_U_BAR: return {d.u_bar, B * T * m};
    case DDP_K: return {d.K, B * T * m * n};
    case DDP_KAPPA: return {d.kappa, B * T * m};
    case DDP_DV: return {d.dV, B * T};
    case DDP_FX: return {d.fx, B * T * n * n};
    case DDP_FU: return {d.fu, B * T * n * m};
    case DDP_COST: return {d.L, B};
    case DDP_EPS: return {d.eps, B};
    case DDP_IMPROVEMENT: return {d.improvement, B};
    case DDP_X0: return {(void*)d.x0, B * n};
    case DDP_X_NOM: return {(void*)d.x_nom, B * n};
    case DDP_CONVERGED_COST: return {d.L_conv, B};
    case DDP_CAND_COST: return {d.Lc, B * A};
    case DDP_CAND_EXPECTED: return {d.Ec, B * A};
    case DDP_CAND_X: return {d.xc, B * A * N * n};
    case DDP_CAND_U: return {d.uc, B * A * T * m};
    default: return {nullptr, 0};
  }
}
ArrInfo iarr(ddp_solver* s, int which) {
  Dev& d = s->d;
  const size_t B = d.B, T = d.T;
  switch (which) {
    case DDP_I_STATUS: return {d.status, B};
    case DDP_I_LS_ITERS: return {d.ls_iters, B};
    case DDP_I_ITERS: return {d.iters, B};
    case DDP_I_NUM_KEYPOINTS: return {d.kpcount, B};
    case DDP_I_KEYPOINTS: return {d.kplist, B * T};
    case DDP_I_ACTIVE: return {d.active, B};
    case DDP_I_RESOLVES: return {d.resolves, B};
    default: return {nullptr, 0};
  }
}

}  // namespace

extern "C" {

const char* ddp_last_error(void) { return g_err.c_str(); }

int ddp_model_dims(int model_id, int* n, int* m, int* nparams) {
  if (model_dims(model_id, n, m, nparams) != 0) {
    g_err = "unknown model id";
    return DDP_ERR_ARG;
  }
  return 0;
}

size_t ddp_workspace_bytes(int model_id, int N, int B, int A) {
  Dev d;
  memset(&d, 0, sizeof(d));
  int np;
  if (model_dims(model_id, &d.n, &d.m, &np) != 0 || N < 3 || B < 1 || A < 1) return 0;
  d.N = N;
  d.T = N - 1;
  d.B = B;
  d.A = A;
  Carver c{nullptr, 0};
  double* p;
  carve(d, &p, np, c);
  return c.off + 256;
}

int ddp_create(ddp_solver_t** out, int model_id, const double* params_host, int nparams, int N,
               int B, int A, void* workspace_dev, size_t workspace_bytes, void* stream) {
  int n, m, np;
  if (!out || model_dims(model_id, &n, &m, &np) != 0) {
    g_err = "unknown model id";
    return DDP_ERR_ARG;
  }
  *out = nullptr;
  if (nparams != np || N < 3 || B < 1 || A < 1 || !workspace_dev || !params_host) {
    g_err = "bad argument (nparams/N/B/A/workspace)";
    return DDP_ERR_ARG;
  }
  if (workspace_bytes < ddp_workspace_bytes(model_id, N, B, A)) {
    g_err = "workspace too small";
    return DDP_ERR_WORKSPACE;
  }
  // the arena decides the device: every later call switches to it (and back)
  cudaPointerAttributes attr;
  CK(cudaPointerGetAttributes(&attr, workspace_dev));
  if (attr.type != cudaMemoryTypeDevice) {
    g_err = "workspace_dev is not device memory";
    return DDP_ERR_ARG;
  }
  ddp_solver* s = new ddp_solver();
  memset(&s->d, 0, sizeof(Dev));
  s->device = attr.device;
  s->model = model_id;
  s->np = np;
  s->stream = (cudaStream_t)stream;
  s->launches = 0;
  s->timings_valid = false;
  s->cfg_bwd = s->cfg_bwd_sym = false;
  s->fused_ctas = 0;
  s->sms = 0;
  s->h_counters = nullptr;
  for (int i = 0; i < 4; ++i) s->ev[i] = nullptr;
  s->scalar_backward = getenv("DDP_SCALAR_BACKWARD") != nullptr;
  {
    const char* sb = getenv("DDP_SMALL_BACKWARD");
    s->small_backward = !(sb && std::string(sb) == "cta");
  }
  Dev& d = s->d;
  d.n = n; d.m = m; d.N = N; d.T = N - 1; d.B = B; d.A = A;
  Carver c{(char*)workspace_dev, 0};
  double* params;
  carve(d, &params, np, c);
  s->quad_sub = 0;
  if (model_id == MODEL_QUADRUPED || model_id == MODEL_QUADRUPED_QUAT) {
    const int sub = (int)params_host[1];
    if (sub == 1 || sub == 2) s->quad_sub = sub;
  }
  {
    const char* inv = getenv("DDP_BWD_INVERSE");   // "gauss-jordan": no Newton-Schulz (debug / parity tests)
    s->d.bwd_flags = (inv && std::string(inv) == "gauss-jordan") ? 1 : 0;
    const char* rmode = getenv("DDP_QUAD_ROLLOUT");
    s->quad_rollout8 = !(rmode && std::string(rmode) == "generic");
    const char* amode = getenv("DDP_ARM_ROLLOUT");
    s->arm_rollout8 = !(amode && std::string(amode) == "generic");
    // default: the fused structured kernel; DDP_QUAD_LINEARIZE=ad selects the generic AD kernel
    const char* mode = getenv("DDP_QUAD_LINEARIZE");
    s->quad_fused = !(mode && std::string(mode) == "ad");
  }
  d.params = params;
  for (int i = 0; i < 32; ++i) d.pm[i] = (i < np) ? params_host[i] : 0.0;
  GUARD(s);
  // from here on a failure must release what was acquired: run the rest in a lambda and
  // destroy the half-built solver if it reports an error
  auto init = [&]() -> int {
    CK(cudaMemsetAsync(workspace_dev, 0, c.off, s->stream));
    CK(cudaMemcpyAsync(params, params_host, np * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    CK(cudaHostAlloc(&s->h_counters, 4 * sizeof(int), cudaHostAllocDefault));
    for (int i = 0; i < 4; ++i) CK(cudaEventCreate(&s->ev[i]));
    // Q = R = Qf = I (ilqr.py:65-67)
    std::vector<double> I(n * n, 0.0), Im(m * m, 0.0);
    for (int i = 0; i < n; ++i) I[i * n + i] = 1.0;
    for (int i = 0; i < m; ++i) Im[i * m + i] = 1.0;
    CK(cudaMemcpyAsync((void*)d.Q, I.data(), n * n * 8, cudaMemcpyHostToDevice, s->stream));
    CK(cudaMemcpyAsync((void*)d.Qf, I.data(), n * n * 8, cudaMemcpyHostToDevice, s->stream));
    CK(cudaMemcpyAsync((void*)d.R, Im.data(), m * m * 8, cudaMemcpyHostToDevice, s->stream));
    d.diag_cost = 1;
    CK(cudaStreamSynchronize(s->stream));
    int rc = ddp_set_options(s, 1e-2, 0.95, 0.0);
    if (rc) return rc;
    rc = ddp_set_keypoints(s, DDP_KP_SET_INTERVAL, 1, 0, 0.0, 0.0);
    if (rc) return rc;
    return ddp_begin_solve(s);
  };
  const int rc = init();
  if (rc) {
    const std::string keep = g_err;
    ddp_destroy(s);
    g_err = keep;
    return rc;
  }
  *out = s;
  return 0;
}

int ddp_destroy(ddp_solver_t* s) {
  if (!s) return 0;
  {
    GUARD(s);
    cudaStreamSynchronize(s->stream);
    for (int i = 0; i < 4; ++i)
      if (s->ev[i]) cudaEventDestroy(s->ev[i]);
    if (s->h_counters) cudaFreeHost(s->h_counters);
  }
  delete s;
  return 0;
}

int ddp_set_options(ddp_solver_t* s, double delta, double beta, double gamma) {
  GUARD(s);
  if (!(beta > 0.0 && beta < 1.0)) {
    g_err = "beta must be in (0,1)";
    return DDP_ERR_ARG;
  }
  s->d.delta = delta;
  s->d.gamma = gamma;
  s->beta = beta;
  s->eps_host.clear();
  double eps = 1.0;
  while (eps >= 1e-8 && (int)s->eps_host.size() < kMaxEps) {  // ilqr.py:300-302,335
    s->eps_host.push_back(eps);
    eps *= beta;
  }
  s->d.n_eps = (int)s->eps_host.size();
  CK(cudaMemcpyAsync((void*)s->d.eps_table, s->eps_host.data(), s->eps_host.size() * 8,
                     cudaMemcpyHostToDevice, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  return 0;
}

int ddp_set_regularization(ddp_solver_t* s, double quu_reg) {
  if (!(quu_reg >= 0.0)) {
    g_err = "quu_reg must be >= 0";
    return DDP_ERR_ARG;
  }
  s->d.quu_reg = quu_reg;
  return 0;
}

int ddp_set_control_limits(ddp_solver_t* s, const double* u_min, const double* u_max) {
  GUARD(s);
  const int m = s->d.m;
  if (!u_min || !u_max) {   // off: the reference's behaviour (SetControlLimits is a no-op there)
    s->d.u_min = s->d.u_max = nullptr;
    return 0;
  }
  for (int i = 0; i < m; ++i)
    if (!(u_min[i] <= u_max[i])) {
      g_err = "u_min must be <= u_max";
      return DDP_ERR_ARG;
    }
  CK(cudaMemcpyAsync(s->d.u_lim_buf, u_min, m * 8, cudaMemcpyHostToDevice, s->stream));
  CK(cudaMemcpyAsync(s->d.u_lim_buf + m, u_max, m * 8, cudaMemcpyHostToDevice, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  s->d.u_min = s->d.u_lim_buf;
  s->d.u_max = s->d.u_lim_buf + m;
  return 0;
}

int ddp_set_keypoints(ddp_solver_t* s, int method, int minN, int maxN, double jerk_threshold,
                      double iterative_error_threshold) {
  if (method < 0 || method > 2) {
    g_err = "unknown interpolation method";
    return DDP_ERR_ARG;
  }
  if (minN < 1) {
    g_err = "minN must be >= 1";
    return DDP_ERR_ARG;
  }
  s->d.kp_method = method;
  s->d.minN = minN;
  s->d.maxN = maxN;
  s->d.jerk_thr = jerk_threshold;
  s->d.err_thr = iterative_error_threshold;
  return 0;
}

int ddp_set_cost(ddp_solver_t* s, const double* Q, const double* R, const double* Qf) {
  GUARD(s);
  const int n = s->d.n, m = s->d.m;
  bool diag = true;
  for (int i = 0; i < n && diag; ++i)
    for (int j = 0; j < n; ++j)
      if (i != j && (Q[i * n + j] != 0.0 || Qf[i * n + j] != 0.0)) {
        diag = false;
        break;
      }
  for (int i = 0; i < m && diag; ++i)
    for (int j = 0; j < m; ++j)
      if (i != j && R[i * m + j] != 0.0) {
        diag = false;
        break;
      }
  s->d.diag_cost = diag ? 1 : 0;
  CK(cudaMemcpyAsync((void*)s->d.Q, Q, n * n * 8, cudaMemcpyHostToDevice, s->stream));
  CK(cudaMemcpyAsync((void*)s->d.R, R, m * m * 8, cudaMemcpyHostToDevice, s->stream));
  CK(cudaMemcpyAsync((void*)s->d.Qf, Qf, n * n * 8, cudaMemcpyHostToDevice, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  return 0;
}

int ddp_set_target(ddp_solver_t* s, const double* x_nom, int per_trajectory) {
  GUARD(s);
  const int n = s->d.n, B = s->d.B;
  if (per_trajectory) {
    CK(cudaMemcpyAsync((void*)s->d.x_nom, x_nom, (size_t)B * n * 8, cudaMemcpyHostToDevice, s->stream));
  } else {
    std::vector<double> rep((size_t)B * n);
    for (int b = 0; b < B; ++b) memcpy(&rep[(size_t)b * n], x_nom, n * 8);
    CK(cudaMemcpyAsync((void*)s->d.x_nom, rep.data(), (size_t)B * n * 8, cudaMemcpyHostToDevice, s->stream));
    CK(cudaStreamSynchronize(s->stream));   // rep goes out of scope
    return 0;
  }
  CK(cudaStreamSynchronize(s->stream));
  return 0;
}

int ddp_set_initial_state(ddp_solver_t* s, const double* x0) {
  GUARD(s);
  CK(cudaMemcpyAsync((void*)s->d.x0, x0, (size_t)s->d.B * s->d.n * 8, cudaMemcpyHostToDevice, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  return 0;
}

int ddp_set_initial_guess(ddp_solver_t* s, const double* u_guess) {
  GUARD(s);
  CK(cudaMemcpyAsync(s->d.u_bar, u_guess, (size_t)s->d.B * s->d.T * s->d.m * 8, cudaMemcpyHostToDevice,
                     s->stream));
  CK(cudaStreamSynchronize(s->stream));
  return 0;
}

int ddp_reset(ddp_solver_t* s) {
  GUARD(s);
  const int which[] = {DDP_X_BAR, DDP_U_BAR, DDP_K, DDP_KAPPA, DDP_DV, DDP_FX, DDP_FU};
  for (int w : which) {
    ArrInfo a = arr(s, w);
    CK(cudaMemsetAsync(a.ptr, 0, a.elems * 8, s->stream));
  }
  return ddp_begin_solve(s);
}

int ddp_mpc_shift(ddp_solver_t* s, int replan_steps) {
  GUARD(s);
  if (replan_steps < 1 || replan_steps >= s->d.N) {
    g_err = "replan_steps must be in [1, N)";
    return DDP_ERR_ARG;
  }
  mpc_shift_kernel<<<s->d.B, 64, 0, s->stream>>>(s->d, replan_steps);
  s->launches++;
  CK(cudaGetLastError());
  return 0;
}

int ddp_set_mpc_rearm(ddp_solver_t* s, int replan_steps, const double* target_advance) {
  GUARD(s);
  if (replan_steps < 0 || replan_steps >= s->d.N) {
    g_err = "replan_steps must be in [0, N)";
    return DDP_ERR_ARG;
  }
  s->d.mpc_replan = replan_steps;
  s->d.mpc_target_adv = nullptr;
  if (replan_steps > 0 && target_advance) {
    CK(cudaMemcpyAsync(s->d.mpc_target_adv_buf, target_advance, s->d.n * 8, cudaMemcpyHostToDevice, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    s->d.mpc_target_adv = s->d.mpc_target_adv_buf;
  }
  return 0;
}

int ddp_apply_staged_inputs(ddp_solver_t* s, const double* x0_dev, const double* u_dev) {
  GUARD(s);
  if (!x0_dev || !u_dev) {
    g_err = "null staging pointer";
    return DDP_ERR_ARG;
  }
  apply_staged_kernel<<<s->d.B, 128, 0, s->stream>>>(s->d, x0_dev, u_dev);
  s->launches++;
  CK(cudaGetLastError());
  return 0;
}

int ddp_begin_solve(ddp_solver_t* s) {
  GUARD(s);
  const int B = s->d.B;
  std::vector<double> inf(B, INFINITY);
  std::vector<int> ones(B, 1);
  CK(cudaMemcpyAsync(s->d.L, inf.data(), B * 8, cudaMemcpyHostToDevice, s->stream));
  CK(cudaMemcpyAsync(s->d.improvement, inf.data(), B * 8, cudaMemcpyHostToDevice, s->stream));
  CK(cudaMemcpyAsync(s->d.active, ones.data(), B * 4, cudaMemcpyHostToDevice, s->stream));
  CK(cudaMemsetAsync(s->d.status, 0, B * 4, s->stream));
  CK(cudaMemsetAsync(s->d.iters, 0, B * 4, s->stream));
  CK(cudaMemsetAsync(s->d.ls_iters, 0, B * 4, s->stream));
  CK(cudaMemsetAsync(s->d.rearm, 0, B * 4, s->stream));
  CK(cudaMemsetAsync(s->d.rearm_mark, 0, B * 4, s->stream));
  CK(cudaMemsetAsync(s->d.resolves, 0, B * 4, s->stream));
  // per-SM CTA-slot bitmasks of the backward sweep: all free between launches; re-zero them in case
  // an earlier launch was aborted with slots taken (they only steer warp roles, never results)
  CK(cudaMemsetAsync(s->d.sm_slots, 0, 1024 * sizeof(int), s->stream));
  CK(cudaStreamSynchronize(s->stream));
  return 0;
}

int ddp_iterate(ddp_solver_t* s, int* n_active) {
  GUARD(s);
  int rc = iterate_impl(s, true, false);
  if (rc) return rc;
  CK(cudaMemcpyAsync(s->h_counters + 1, s->d.counters + 1, sizeof(int), cudaMemcpyDeviceToHost,
                     s->stream));
  CK(cudaStreamSynchronize(s->stream));
  CK(cudaGetLastError());
  if (n_active) *n_active = s->h_counters[1];
  return 0;
}

int ddp_iterate_linesearch(ddp_solver_t* s) {
  GUARD(s);
  int rc = iterate_linesearch_impl(s, true, false);
  if (rc) return rc;
  CK(cudaStreamSynchronize(s->stream));
  CK(cudaGetLastError());
  return 0;
}

int ddp_iterate_finish_async(ddp_solver_t* s) {
  GUARD(s);
  int rc = iterate_finish_impl(s);
  if (rc) return rc;
  CK(cudaMemcpyAsync(s->h_counters + 1, s->d.counters + 1, sizeof(int), cudaMemcpyDeviceToHost,
                     s->stream));
  return 0;
}

int ddp_iterate_wait(ddp_solver_t* s, int* n_active) {
  GUARD(s);
  CK(cudaStreamSynchronize(s->stream));
  CK(cudaGetLastError());
  if (n_active) *n_active = s->h_counters[1];
  return 0;
}

int ddp_solve(ddp_solver_t* s, int max_iters, int* iters_done) {
  GUARD(s);
  int rc = ddp_begin_solve(s);
  if (rc) return rc;
  int it = 0, n_active = 1;
  while (n_active > 0 && (max_iters <= 0 || it < max_iters)) {
    rc = ddp_iterate(s, &n_active);
    if (rc) return rc;
    ++it;
  }
  if (iters_done) *iters_done = it;
  return 0;
}

int ddp_run_phase(ddp_solver_t* s, int phase) {
  GUARD(s);
  Dev& d = s->d;
  if (phase < DDP_PHASE_LINESEARCH || phase > DDP_PHASE_BACKWARD) {
    g_err = "unknown phase";
    return DDP_ERR_ARG;
  }
  // every trajectory runs the phase, converged or not; the active flags and statuses of a solve
  // in progress are put back afterwards
  CK(cudaMemcpyAsync(d.active_save, d.active, d.B * sizeof(int), cudaMemcpyDeviceToDevice, s->stream));
  CK(cudaMemcpyAsync(d.status_save, d.status, d.B * sizeof(int), cudaMemcpyDeviceToDevice, s->stream));
  LAUNCH1(begin_iter_kernel, d, 1);
  int rc = 0;
  switch (phase) {
    case DDP_PHASE_LINESEARCH: rc = phase_linesearch(s, true); break;
    case DDP_PHASE_DERIVATIVES: rc = phase_derivatives(s); break;
    default: rc = do_backward(s); break;
  }
  if (rc) return rc;
  CK(cudaMemcpyAsync(d.active, d.active_save, d.B * sizeof(int), cudaMemcpyDeviceToDevice, s->stream));
  CK(cudaMemcpyAsync(d.status, d.status_save, d.B * sizeof(int), cudaMemcpyDeviceToDevice, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  CK(cudaGetLastError());
  return 0;
}

int ddp_get(ddp_solver_t* s, int which, double* dst_host) {
  GUARD(s);
  ArrInfo a = arr(s, which);
  if (!a.ptr) {
    g_err = "unknown array";
    return DDP_ERR_ARG;
  }
  CK(cudaMemcpyAsync(dst_host, a.ptr, a.elems * 8, cudaMemcpyDeviceToHost, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  return 0;
}

int ddp_put(ddp_solver_t* s, int which, const double* src_host) {
  GUARD(s);
  ArrInfo a = arr(s, which);
  if (!a.ptr) {
    g_err = "unknown array";
    return DDP_ERR_ARG;
  }
  CK(cudaMemcpyAsync(a.ptr, src_host, a.elems * 8, cudaMemcpyHostToDevice, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  return 0;
}

int ddp_get_int(ddp_solver_t* s, int which, int* dst_host) {
  GUARD(s);
  ArrInfo a = iarr(s, which);
  if (!a.ptr) {
    g_err = "unknown int array";
    return DDP_ERR_ARG;
  }
  CK(cudaMemcpyAsync(dst_host, a.ptr, a.elems * 4, cudaMemcpyDeviceToHost, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  return 0;
}

void* ddp_device_ptr(ddp_solver_t* s, int which) { return arr(s, which).ptr; }
size_t ddp_array_elems(ddp_solver_t* s, int which) { return arr(s, which).elems; }

int ddp_last_timings(ddp_solver_t* s, float ms[4]) {
  GUARD(s);
  if (!s->timings_valid) {
    g_err = "no iteration has run";
    return DDP_ERR_ARG;
  }
  CK(cudaEventSynchronize(s->ev[3]));
  CK(cudaEventElapsedTime(&ms[0], s->ev[0], s->ev[1]));
  CK(cudaEventElapsedTime(&ms[1], s->ev[1], s->ev[2]));
  CK(cudaEventElapsedTime(&ms[2], s->ev[2], s->ev[3]));
  CK(cudaEventElapsedTime(&ms[3], s->ev[0], s->ev[3]));
  return 0;
}

long long ddp_launch_count(ddp_solver_t* s) { return s->launches; }

int ddp_peak_fp64(void* stream, int use_mma, double* tflops) {
  cudaStream_t st = (cudaStream_t)stream;
  cudaDeviceProp prop;
  int dev;
  CK(cudaGetDevice(&dev));
  CK(cudaGetDeviceProperties(&prop, dev));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
  double* out;
  CK(cudaMalloc(&out, (size_t)blocks * threads * 8));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaEventRecord(e0, st));
    if (use_mma) peak_dmma_kernel<<<blocks, threads, 0, st>>>(out, iters);
    else peak_dfma_kernel<<<blocks, threads, 0, st>>>(out, iters);
    CK(cudaEventRecord(e1, st));
    CK(cudaEventSynchronize(e1));
  }
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  double flops;
  if (use_mma) flops = (double)blocks * (threads / 32) * iters * 4.0 * (8 * 8 * 4 * 2);
  else flops = (double)blocks * threads * iters * 8.0 * 2.0;
  *tflops = flops / (ms * 1e-3) / 1e12;
  CK(cudaFree(out));
  CK(cudaEventDestroy(e0));
  CK(cudaEventDestroy(e1));
  return 0;
}

#ifdef DDP_ROLL_PROFILE
int ddp_debug_roll_profile(long long* out16) {
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpyFromSymbol(out16, ddp::g_roll_prof, sizeof(long long) * 16));
  return 0;
}
#endif
#ifdef DDP_BWD_PROFILE
// debug builds only: per-phase cycle totals recorded by backward_sym_kernel
int ddp_debug_bwd_profile(long long* out128) {
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpyFromSymbol(out128, ddp::g_bwd_prof, sizeof(long long) * 128));
  return 0;
}
#endif

}  // extern "C"
