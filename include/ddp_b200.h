/*
 * ddp_b200.h -- C ABI of the B200-native batched iLQR hot path.
 *
 * The reference (vincekurtz/drake_ddp) has no FFI: its boundary is the Python class
 * IterativeLinearQuadraticRegulator (/root/reference/ilqr.py:12).  This header is what
 * that class's hot loops bind to when they are replaced by the CUDA library; each entry
 * point cites the reference interface it stands in for.  Plain pointers and sizes only;
 * no torch / C++ types cross this boundary.  All matrices are row-major fp64.
 *
 * Memory: the caller owns device memory.  It asks ddp_workspace_bytes() for the arena
 * size, allocates it (e.g. one torch uint8 tensor), and passes the device pointer and a
 * cudaStream_t to ddp_create().  The library never calls cudaMalloc for trajectory data.
 *
 * Device layouts (B trajectories, N knot points, T = N-1, A = line-search candidates
 * evaluated per round):
 *   x_bar [B][N][n]      u_bar [B][T][m]      kappa [B][T][m]     dV [B][T]
 *   K     [B][T][m][n]   fx    [B][T][n][n]   fu    [B][T][n][m]
 * The reference keeps time last: x_bar (n,N), K (m,n,N-1), fx (n,n,N-1)
 * (/root/reference/ilqr.py:70-83); the Python class converts at its surface.
 *
 * Every function returns 0 on success, a negative ddp_status on error (message via
 * ddp_last_error()).  No C++ exception crosses the boundary.
 */
#ifndef DDP_B200_H_
#define DDP_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ddp_solver ddp_solver_t;

enum ddp_status {
  DDP_OK = 0,
  DDP_ERR_ARG = -1,       /* bad argument / unknown model                           */
  DDP_ERR_CUDA = -2,      /* CUDA runtime error                                      */
  DDP_ERR_WORKSPACE = -3, /* arena too small                                         */
};

/* per-trajectory status codes (ddp_get_int(DDP_I_STATUS)) */
enum ddp_traj_status {
  DDP_TRAJ_RUNNING = 0,
  DDP_TRAJ_CONVERGED = 1,          /* improvement <= delta (ilqr.py:692)               */
  DDP_TRAJ_LINESEARCH_FAILED = 2,  /* RuntimeError at ilqr.py:337                      */
};

/* keypoint methods (utils_derivs_interpolation.py:5, dispatch at ilqr.py:396-404) */
enum ddp_keypoint_method {
  DDP_KP_SET_INTERVAL = 0,
  DDP_KP_ADAPTIVE_JERK = 1,
  DDP_KP_ITERATIVE_ERROR = 2,
};

/* double arrays addressable with ddp_get / ddp_put / ddp_device_ptr */
enum ddp_array {
  DDP_X_BAR = 0,   /* [B][N][n]     ilqr.py:70  */
  DDP_U_BAR = 1,   /* [B][T][m]     ilqr.py:71  */
  DDP_K = 2,       /* [B][T][m][n]  ilqr.py:79  */
  DDP_KAPPA = 3,   /* [B][T][m]     ilqr.py:78  */
  DDP_DV = 4,      /* [B][T]        ilqr.py:83  */
  DDP_FX = 5,      /* [B][T][n][n]  ilqr.py:74  */
  DDP_FU = 6,      /* [B][T][n][m]  ilqr.py:75  */
  DDP_COST = 7,    /* [B]  total cost L of the current x_bar,u_bar (ilqr.py:707)      */
  DDP_EPS = 8,     /* [B]  accepted line-search step of the last iteration            */
  DDP_IMPROVEMENT = 9, /* [B] L_prev - L_new of the last iteration (ilqr.py:706)      */
  DDP_X0 = 10,     /* [B][n]                                                          */
  DDP_X_NOM = 11,  /* [B][n]                                                          */
  DDP_CAND_COST = 12,     /* [B][A] costs of the last line-search round               */
  DDP_CAND_EXPECTED = 13, /* [B][A] expected improvements of the last round           */
  DDP_CAND_X = 14,        /* [B][A][N][n] candidate rollouts of the last round         */
  DDP_CAND_U = 15,        /* [B][A][T][m]                                              */
  DDP_CONVERGED_COST = 16, /* [B] return value L of the last Solve() that converged (ilqr.py:710) */
};

/* int arrays addressable with ddp_get_int */
enum ddp_int_array {
  DDP_I_STATUS = 0,     /* [B] ddp_traj_status                                        */
  DDP_I_LS_ITERS = 1,   /* [B] rollouts the reference would have run (ilqr.py:303)    */
  DDP_I_ITERS = 2,      /* [B] iLQR iterations done since ddp_begin_solve              */
  DDP_I_NUM_KEYPOINTS = 3, /* [B] len(keyPoints) of the last iteration (ilqr.py:406)  */
  DDP_I_KEYPOINTS = 4,  /* [B][T] ascending keypoint indices, first NUM_KEYPOINTS valid */
  DDP_I_ACTIVE = 5,     /* [B] 1 while the trajectory is still iterating               */
  DDP_I_RESOLVES = 6,   /* [B] MPC resolves completed on the device (ddp_set_mpc_rearm)  */
};

/* phases, for teacher-forced checks of one reference function at a time */
enum ddp_phase {
  DDP_PHASE_LINESEARCH = 0,  /* _linesearch + commit          ilqr.py:274-337,375-376 */
  DDP_PHASE_DERIVATIVES = 1, /* _get_derivatives              ilqr.py:380-621         */
  DDP_PHASE_BACKWARD = 2,    /* _backward_pass                ilqr.py:623-667         */
};

const char* ddp_last_error(void);

/* n, m of a compiled-in analytic model; replaces context.get_discrete_state_vector().size()
 * and input_port.size() (ilqr.py:57-58). */
int ddp_model_dims(int model_id, int* n, int* m, int* nparams);

/* Arena size for (model, N, B, A). */
size_t ddp_workspace_bytes(int model_id, int N, int B, int A);

/* Constructor, ilqr.py:21-100.  params_host: model parameters (params[0] = dt).
 * workspace_dev: device arena of at least ddp_workspace_bytes(); stream: cudaStream_t.
 * All trajectory state starts zeroed, costs Q=R=Qf=I, like a fresh reference object. */
int ddp_create(ddp_solver_t** out, int model_id, const double* params_host, int nparams, int N,
               int B, int A, void* workspace_dev, size_t workspace_bytes, void* stream);
int ddp_destroy(ddp_solver_t* s);

/* delta, beta, gamma of the constructor (ilqr.py:52-54).  Builds the eps table
 * 1, beta, beta^2 ... >= 1e-8 by repeated multiplication (ilqr.py:300-302,335). */
int ddp_set_options(ddp_solver_t* s, double delta, double beta, double gamma);
/* derivs_keypoint_method (ilqr.py:97-100). */
int ddp_set_keypoints(ddp_solver_t* s, int method, int minN, int maxN, double jerk_threshold,
                      double iterative_error_threshold);
/* Extension (no counterpart in the reference, which inverts Quu as is,
 * ilqr.py:654-655): Quu <- Quu + quu_reg * I before the inverse.  Default 0 = reference. */
int ddp_set_regularization(ddp_solver_t* s, double quu_reg);
/* SetControlLimits (ilqr.py:158-159) is `pass` in the reference.  Extension: with
 * u_min, u_max ([m], host pointers) the line-search rollout clamps every control to the box,
 * u_t = clip(u_bar_t - eps kappa_t - K_t (x_t - x_bar_t)); NULL, NULL (the default) switches it off
 * and restores the reference's behaviour bit for bit. */
int ddp_set_control_limits(ddp_solver_t* s, const double* u_min, const double* u_max);
/* SetRunningCost / SetTerminalCost (ilqr.py:120-146); host pointers, shared by the batch. */
int ddp_set_cost(ddp_solver_t* s, const double* Q, const double* R, const double* Qf);
/* SetTargetState (ilqr.py:111-118); x_nom is [n] (per_trajectory=0) or [B][n]. */
int ddp_set_target(ddp_solver_t* s, const double* x_nom, int per_trajectory);
/* SetInitialState (ilqr.py:102-109); x0 is [B][n] on the host. */
int ddp_set_initial_state(ddp_solver_t* s, const double* x0);
/* SetInitialGuess (ilqr.py:148-156); u_guess is [B][T][m] on the host. */
int ddp_set_initial_guess(ddp_solver_t* s, const double* u_guess);
/* Back to a freshly constructed object's trajectory state (zeros), keeping costs/options. */
int ddp_reset(ddp_solver_t* s);

/* Receding-horizon warm start without a host round trip (the np.block shift of
 * acrobot.py:145-153 / mini_cheetah.py:190-198): u_bar <- [u_bar[:, r:], last column x r],
 * x0 <- x_bar[:, r] for every trajectory.  K, kappa, x_bar stay (stale-state semantics). */
int ddp_mpc_shift(ddp_solver_t* s, int replan_steps);

/* The whole receding-horizon loop of mini_cheetah.py:186-206 / acrobot.py:142-160 on the device:
 * with replan_steps > 0 a trajectory whose Solve() converges (ilqr.py:692) is not frozen; the same
 * kernel sequence applies ddp_mpc_shift to it alone, adds target_advance [n] (host pointer, may be
 * NULL) to its x_nom (the moving target of mini_cheetah.py:151-156), sets L = improvement = inf
 * (ilqr.py:681-682) and the next ddp_iterate is the first iteration of its next resolve -- no
 * host round trip, every trajectory of the batch keeps iterating.  DDP_I_RESOLVES counts,
 * DDP_CONVERGED_COST holds the value the last converged Solve() returned.  replan_steps = 0
 * (default) restores Solve() semantics: converged trajectories stop. */
int ddp_set_mpc_rearm(ddp_solver_t* s, int replan_steps, const double* target_advance);

/* SetInitialState + SetInitialGuess (ilqr.py:102-109,148-156) of a loop that keeps x0 / u_guess in
 * host buffers and has already staged them on the device (x0_dev [B][n], u_dev [B][T][m], device
 * pointers): enqueues the copy on the solver's stream.  Trajectories the device re-armed itself since
 * the last call (ddp_set_mpc_rearm) keep their newer x0 / tape: the staged rows were read back
 * before that shift. */
int ddp_apply_staged_inputs(ddp_solver_t* s, const double* x0_dev, const double* u_dev);

/* Solve(), ilqr.py:669-710, split so the host can print the per-iteration table:
 * ddp_begin_solve sets L = inf, improvement = inf for every trajectory (ilqr.py:681-682);
 * ddp_iterate runs one forward pass + backward pass (ilqr.py:695-697) for every
 * trajectory whose improvement is still > delta and returns how many remain active;
 * ddp_solve loops ddp_iterate until none is active or max_iters (<=0: unlimited). */
int ddp_begin_solve(ddp_solver_t* s);
int ddp_iterate(ddp_solver_t* s, int* n_active);
int ddp_solve(ddp_solver_t* s, int max_iters, int* iters_done);

/* ddp_iterate in three calls, for callers that exchange host buffers every iteration (MPC
 * loops, acrobot.py:142-160 / mini_cheetah.py:186-206):
 *   ddp_iterate_linesearch    _linesearch + commit (ilqr.py:274-337, 375-376); returns when u_bar,
 *                             x_bar of the iteration are final (the line search is synchronous:
 *                             its rounds are sized from the number of unresolved trajectories);
 *   ddp_iterate_finish_async  enqueues _get_derivatives + _backward_pass (ilqr.py:380-415,
 *                             623-667) and the bookkeeping of :706-708 on the solver's stream and
 *                             returns; they only read u_bar / x_bar, so the caller may copy the
 *                             new controls to the host on another stream meanwhile;
 *   ddp_iterate_wait          waits for them and returns how many trajectories remain active.
 * ddp_iterate(s, n) == linesearch; finish_async; wait(n). */
int ddp_iterate_linesearch(ddp_solver_t* s);
int ddp_iterate_finish_async(ddp_solver_t* s);
int ddp_iterate_wait(ddp_solver_t* s, int* n_active);

/* One phase only (teacher-forced tests). */
int ddp_run_phase(ddp_solver_t* s, int phase);

/* Array access.  ddp_get/ddp_put copy between host memory and the device arena on the
 * solver's stream and synchronise; ddp_device_ptr exposes the arena for zero-copy views
 * (e.g. the per-trajectory cost vector handed to an NCCL all-gather). */
int ddp_get(ddp_solver_t* s, int which, double* dst_host);
int ddp_put(ddp_solver_t* s, int which, const double* src_host);
int ddp_get_int(ddp_solver_t* s, int which, int* dst_host);
void* ddp_device_ptr(ddp_solver_t* s, int which);
size_t ddp_array_elems(ddp_solver_t* s, int which);

/* Device time of the last iteration's phases in ms: [0] line search (time_fp, ilqr.py:367),
 * [1] derivatives (time_getDerivs, :372), [2] backward pass (time_backwardsPass, :699),
 * [3] whole iteration. */
int ddp_last_timings(ddp_solver_t* s, float ms[4]);
/* Number of kernels this library has launched on the solver's stream so far. */
long long ddp_launch_count(ddp_solver_t* s);

/* Raw microbenchmarks used by bench.py to state the fp64 roofline next to the HBM one:
 * returns achieved TFLOP/s of a register-resident DFMA loop / DMMA (mma.sync m8n8k4 f64). */
int ddp_peak_fp64(void* stream, int use_mma, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* DDP_B200_H_ */
