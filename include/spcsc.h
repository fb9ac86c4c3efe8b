/* spcsc.h -- C ABI of libspcsc.so, the B200-native convolutional sparse coding engine.
 *
 * The reference (bwohlberg/sporco, pure Python) has no FFI for this path; its extension
 * contract is "subclass and override" (sporco/admm/admm.py:331-377 calls xstep, relax_AX,
 * ystep, ustep, compute_residuals, update_rho; sporco/pgm/pgm.py:328-370 calls
 * on_iteration_start, backtrack.update | xstep+ystep, compute_residuals) plus one real
 * plug-in seam, the `sporco_cuda` import in sporco/cuda/__init__.py:6-18.  Each entry point
 * below names the reference routine(s) it stands in for.  INTEGRATION.md shows the ctypes
 * binding a sporco maintainer would add.
 *
 * Conventions
 *  - every function returns 0 (SPCSC_OK) or a negative spcsc_status; no C++ exception crosses
 *    the ABI; spcsc_last_error() gives the message of the most recent failure;
 *  - all array arguments are HOST pointers in the reference's own layouts, C order:
 *      D  (hd, wd, Cd, M)     S  (N0, N1, C, K)     X/Y/U  (N0, N1, Cx, K, M), Cx = C-Cd+1
 *      spectra Xf/Df/Sf as returned by sporco.fft.rfftn(..., axes=(0,1)):  (N0, N1/2+1, ...)
 *    element type float or double according to spcsc_problem.dtype (complex = 2 reals);
 *  - the caller owns host buffers (borrowed for the duration of the call); the library owns
 *    all device memory of a handle; get_* calls copy out and synchronise;
 *  - a handle is bound to one device and one stream and is not thread safe; distinct handles
 *    are independent.
 */
#ifndef SPCSC_H_
#define SPCSC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct spcsc_handle spcsc_handle;

typedef enum spcsc_status {
    SPCSC_OK = 0,
    SPCSC_ERR_INVALID = -1,      /* bad argument / unsupported configuration */
    SPCSC_ERR_CUDA = -2,         /* CUDA runtime error (sticky for the handle) */
    SPCSC_ERR_NOMEM = -3,
    SPCSC_ERR_UNSUPPORTED = -4,  /* valid for the reference but not implemented here */
    SPCSC_ERR_STATE = -5,        /* call sequence error (e.g. iterate before set_dict) */
    SPCSC_ERR_NCCL = -6
} spcsc_status;

enum { SPCSC_F32 = 0, SPCSC_F64 = 1 };

/* Problem dimensions: what sporco/cnvrep.py:33-198 (CSC_ConvRepIndexing) infers. */
typedef struct spcsc_problem {
    int32_t N0, N1;        /* spatial size (both powers of two, N0 >= 2, N1 >= 4) */
    int32_t C, Cd, K, M;   /* signal channels, dictionary channels, images, filters */
    int32_t hd, wd;        /* filter support */
    int32_t dtype;         /* SPCSC_F32 | SPCSC_F64 */
    int32_t device;        /* CUDA device ordinal */
} spcsc_problem;

/* Options of the ADMM solver: sporco/admm/admm.py:148-161 + admm/cbpdn.py:127-134. */
typedef struct spcsc_admm_opts {
    double lmbda;          /* l1 weight                       admm/cbpdn.py:581 */
    double mu;             /* l2,1 weight (joint only)        admm/cbpdn.py:780 */
    double rlx;            /* RelaxParam                      admm/admm.py:877-885 */
    double abs_tol, rel_tol;          /* AbsStopTol, RelStopTol      admm/admm.py:481-484 */
    double ar_scaling;     /* AutoRho.Scaling (tau)           admm/admm.py:553 */
    double ar_rsdl_ratio;  /* AutoRho.RsdlRatio (mu)          admm/admm.py:554 */
    double ar_rsdl_target; /* AutoRho.RsdlTarget (xi), resolved by the caller  admm/cbpdn.py:588-593 */
    int32_t ar_enabled, ar_period, ar_autoscaling, ar_std_residuals;
    int32_t joint;         /* 1: ConvBPDNJoint prox (admm/cbpdn.py:785-794) */
    int32_t nonneg;        /* NonNegCoef                      admm/cbpdn.py:306-307 */
    int32_t no_bndry_cross;/* NoBndryCross                    admm/cbpdn.py:308-311 */
    int32_t fast_solve;    /* FastSolve: skip objective       admm/admm.py:356 */
    int32_t aux_var_obj;   /* AuxVarObj: objective on Y       admm/cbpdn.py:151-164 */
    int32_t linsolve_check;/* LinSolveCheck                   admm/cbpdn.py:283-293 */
    double l2_weight;      /* ConvElasticNet mu: (mu/2)||x||^2, x-step diagonal mu + rho; 0 = plain
                              ConvBPDN.  With it the regl21 column of the rows carries RegL2.
                              admm/cbpdn.py:948-986 */
    int32_t ams_maps;      /* AddMaskSim: the last ams_maps filters are the appended impulse(s); their
                              coefficient maps are neither clipped (NonNegCoef, NoBndryCross) nor counted
                              in RegL1, and the mask enters through the l1 weight (0 / huge).
                              admm/cbpdn.py:2377-2409 */
    int32_t reserved_;
} spcsc_admm_opts;

/* One row of IterationStats (admm/admm.py:182-189, admm/cbpdn.py:512-514, 737-740). */
typedef struct spcsc_itstat {
    double iter, objfun, dfid, regl1, regl21, primal_rsdl, dual_rsdl, eps_primal, eps_dual,
        rho, xslv_relres, reserved;
} spcsc_itstat;

typedef enum spcsc_array {
    SPCSC_ARR_Y = 0, SPCSC_ARR_U = 1, SPCSC_ARR_X = 2,      /* real, (N0,N1,Cx,K,M) */
    SPCSC_ARR_XF = 3,                                       /* complex, (N0,N1f,Cx,K,M) */
    SPCSC_ARR_DF = 4,                                       /* complex, (N0,N1f,Cd,1,M) */
    SPCSC_ARR_SF = 5,                                       /* complex, (N0,N1f,C,K,1) */
    SPCSC_ARR_PGM_X = 6,                                    /* PGM iterate X, real (N0,N1,Cx,K,M) */
    SPCSC_ARR_PGM_XF = 7, SPCSC_ARR_PGM_YF = 8              /* PGM Xf / Yf, complex (N0,N1f,Cx,K,M) */
} spcsc_array;

/* Options of the PGM/FISTA solver that reach the device (pgm/cbpdn.py:115-118, 288-298). */
typedef struct spcsc_pgm_opts {
    double lmbda;
    int32_t nonneg, no_bndry_cross;
} spcsc_pgm_opts;

/* What one proximal-gradient trial reports (all sums over the stored half spectrum):
   [0] F  = obfn_f(Xf)  = 1/2 sum |sum_m Df Xf - Sf|^2, unweighted     pgm/cbpdn.py:358-370
   [1] FY = obfn_f(Yf)                                                  backtrack.py:93
   [2] lin = sum Re(conj(Xf - Yf) gradY)                                pgm/pgm.py:886-894
   [3] dxy2 = sum |Xf - Yf|^2                                           backtrack.py:95
   [4] rsdl = rfl2norm2(Xf - Yf)  (Hermitian weights, 1/N)             pgm/cbpdn.py:314-318
   [5] dfid = rfl2norm2(sum_m Df Xf - Sf)/2                             pgm/cbpdn.py:334-344
   [6] regl1 = ||wl1 * X||_1                                            pgm/cbpdn.py:348-354 */
enum { SPCSC_PGM_F = 0, SPCSC_PGM_FY = 1, SPCSC_PGM_LIN = 2, SPCSC_PGM_DXY2 = 3,
       SPCSC_PGM_RSDL = 4, SPCSC_PGM_DFID = 5, SPCSC_PGM_REGL1 = 6, SPCSC_PGM_NOUT = 8 };

/* ---- library / device queries (stand-ins for sporco_cuda.util, docs/source/modules/sporco.cuda.rst:60-104) */
int spcsc_version(void);
int spcsc_device_count(void);
int spcsc_device_name(int device, char* buf, int buflen);
int spcsc_memory_info(int device, uint64_t* free_bytes, uint64_t* total_bytes);
const char* spcsc_last_error(const spcsc_handle* h);   /* h may be NULL: last create() failure */

/* ---- lifetime */
int spcsc_create(const spcsc_problem* prob, spcsc_handle** out);
int spcsc_destroy(spcsc_handle* h);
int spcsc_synchronize(spcsc_handle* h);

/* ---- problem data */
/* GenericConvBPDN.setdict: Df = rfftn(D, Nv), Gram for the solve.   admm/cbpdn.py:242-256 */
int spcsc_set_dict(spcsc_handle* h, const void* D);
/* Sf = rfftn(S).                                                    admm/cbpdn.py:227-231 */
int spcsc_set_signal(spcsc_handle* h, const void* S);
/* L1Weight after cnvrep.l1Wshape: `shape` is its 5-D internal shape (each entry 1 or the full
   extent of (N0,N1,Cx,K,M)).                                        admm/cbpdn.py:596-597 */
int spcsc_set_l1_weight(spcsc_handle* h, const void* w, const int64_t shape[5]);
/* L21Weight broadcast over (K, M): shape entries 1 or full.         admm/cbpdn.py:781 */
int spcsc_set_l21_weight(spcsc_handle* h, const void* w, const int64_t shape[2]);

/* ---- ADMM solver (ConvBPDN / ConvBPDNJoint) */
int spcsc_admm_configure(spcsc_handle* h, const spcsc_admm_opts* opts);
/* Y = U = 0 (or later set_array), k = 0, rho as given.              admm/admm.py:243-275 */
int spcsc_admm_reset(spcsc_handle* h, double rho);
int spcsc_admm_set_rho(spcsc_handle* h, double rho);
/* Set the iteration counter (restoring a pickled solver; admm/admm.py:331 resumes from self.k). */
int spcsc_admm_set_iter(spcsc_handle* h, int32_t k);
/* Run up to n_iter iterations of admm/admm.py:331-377 on the device.  Stops early (device
   side, no host round trip) once r < epri and s < edua.  `rows` (n_iter entries, may be NULL)
   receives one spcsc_itstat per executed iteration.  Clears a previous stop flag on entry,
   as re-entering ADMM.solve() does. */
int spcsc_admm_iterate(spcsc_handle* h, int32_t n_iter, spcsc_itstat* rows, int32_t* n_done,
                       int32_t* stopped);
int spcsc_admm_get_scalars(spcsc_handle* h, double* rho, int32_t* k);
/* Device time (CUDA events on the handle's stream) of the kernels launched by the most recent
   spcsc_admm_iterate call, in milliseconds, and the number of kernel launches it made. */
int spcsc_admm_last_timing(spcsc_handle* h, float* elapsed_ms, int64_t* launches);
/* Run n_iter iterations with a CUDA event between consecutive kernels and return the summed
   device time per kernel, ms: [0] k_row_fwd, [1] k_col, [2] k_row_inv_prox, [3] k_admm_scalars.
   Measurement aid for bench.py (same kernels, same stream as spcsc_admm_iterate). */
int spcsc_admm_profile(spcsc_handle* h, int32_t n_iter, float kernel_ms[4]);
/* Which kernel schedule the handle uses: info[0] register-plan row-forward kernel, [1] cluster
   column kernel, [2] register-plan prox kernel, [3] cross-iteration fusion (the prox kernel also
   emits the next x-step's row spectra; the row-forward launch is then a gated no-op unless rho
   changed), [4] column kernel of the last batch (0 general, 2 k_col2 clusters, 3 k_col3 persistent clusters
   with pushed sums), [5] images per wavefront group (0: whole batch per launch), [6] streams of the
   wavefront schedule, [7] reserved. */
int spcsc_admm_schedule_info(spcsc_handle* h, int32_t info[8]);

/* ---- PGM / FISTA solver (sporco.pgm.cbpdn.ConvBPDN).  The host keeps the scalar control flow
   of pgm/pgm.py:328-370 and pgm/backtrack.py:74-107 (step size L, momentum t, F <= Q test);
   each call below is one batch of kernels. */
int spcsc_pgm_configure(spcsc_handle* h, const spcsc_pgm_opts* opts);
/* X = X0 (NULL: zeros), Xf = Yf = rfftn(X).                           pgm/cbpdn.py:217-241 */
int spcsc_pgm_reset(spcsc_handle* h, const void* X0);
/* One proximal step from the current Yf with step 1/L (grad_f, PGMDFT.xstep): candidate X, Xf
   are kept on the device, the sums needed for the stopping / backtracking tests come back. */
int spcsc_pgm_trial(spcsc_handle* h, double L, double out[8]);
/* Accept the candidate and take the momentum step Yf = Xf + coef (Xf - Xfprv)  (PGMDFT.ystep). */
int spcsc_pgm_accept(spcsc_handle* h, double coef);
/* Scalars for the step-size policies and the monotone variant, at the current state (two passes over the
   spectra that only form per-frequency sums):
     out[0] = ||grad f(Yf)||^2, out[1] = <grad, Hess grad>               StepSizePolicyCauchy, pgm/stepsize.py:50-87
     out[2] = ||grad - grad_prev||^2, out[3] = <Xf - Xf_prev, grad - grad_prev>   StepSizePolicyBB, :90-145
     out[4] = DFid of the accepted Xf, out[5] = RegL1 of the X last produced      eval_objfn, pgm/cbpdn.py:320-356
   (plain sums over the stored half spectrum for [0..3], like np.sum over rfftn output).  store != 0 afterwards
   remembers the current gradient / iterate as "previous" (StepSizePolicyBB.store_prev_state). */
int spcsc_pgm_policy_stats(spcsc_handle* h, int32_t store, double out[8]);
/* Robust backtracking (pgm/backtrack.py:110-210): Yf = a Xf + b Z with the auxiliary sequence Z (set to Xf on
   first use); save_prev != 0 first keeps the old Yf for the residual of the iteration. */
int spcsc_pgm_combine_y(spcsc_handle* h, double a, double b, int32_t save_prev);
/* Ends an iteration other than by spcsc_pgm_accept.  SPCSC_PGM_FINISH_REJECT (monotone FISTA, pgm/pgm.py:802-831,
   the candidate Z raised the objective): Xf stays, Yf = Xf + c0 (Zf - Xf).  SPCSC_PGM_FINISH_ROBUST: Z += c0 (Xf_new
   - Yf), the candidate is accepted, Yf untouched.  out[0] = rfl2norm2(Xf - Yfprv), the residual of pgm/cbpdn.py:314-318. */
enum { SPCSC_PGM_FINISH_REJECT = 1, SPCSC_PGM_FINISH_ROBUST = 2 };
int spcsc_pgm_finish(spcsc_handle* h, int32_t mode, double c0, double out[2]);

/* Device-side all-reduce of the per-iteration accumulators over peer memory (NVLink / NVSwitch), replacing the
   NCCL call on that path: every rank exports a small block (CUDA IPC, 64-byte handle), the handles of all
   ranks are gathered by the caller and attached; the exchange then happens inside the scalar kernel.  Needs
   an attached communicator (spcsc_attach_comm) first and at most 8 ranks on one node.  The blocks belong to the
   communicator: once one solver has attached them, further solvers on the same communicator pass
   handles64 = NULL and need no exchange of handles.  nranks = 0 detaches.  If exporting or attaching fails
   (SPCSC_ERR_UNSUPPORTED, the handle stays usable) the NCCL path simply stays in use. */
int spcsc_p2p_export(spcsc_handle* h, void* handle64);
int spcsc_p2p_attach(spcsc_handle* h, int32_t rank, int32_t nranks, const void* handles64);

/* pgm.cbpdn.ConvBPDNMask (pgm/cbpdn.py:387-508): data fidelity (1/2)||W (sum_m d_m * x_m - s)||^2.  W: real,
   shape[4] = (N0|1, N1|1, C|1, K|1) broadcast against the signal; NULL switches the mask off.  Affects the
   spcsc_pgm_* calls only.  Single-channel dictionary. */
int spcsc_pgm_set_mask(spcsc_handle* h, const void* W, const int64_t shape[4]);

/* ConvBPDNGradReg (admm/cbpdn.py:993-1216): gradient regulariser (mu/2) sum_m w_m ||G x_m||^2 with weight
   `mu` = spcsc_admm_opts.mu.  ghg: (N0, N1f) real, sum_i |G_i|^2 as signal.gradient_filters returns it;
   wgrd: M reals (GradWeight).  The x-step solves with the diagonal mu w_m ghg + rho (linalg.solvedbd_sm);
   the regl21 column of the rows carries RegGrad.  ghg == NULL switches it off. Single-channel dictionary. */
int spcsc_set_gradreg(spcsc_handle* h, const void* ghg, const void* wgrd);

/* ---- dictionary update: sporco.pgm.ccmod.ConvCnstrMOD as the D step of
   sporco.dictlrn.cbpdndl.ConvBPDNDictLearn (dictlrn/dictlrn.py:327-363), sharing the handle -- and
   the device arrays -- of the X step.  Greyscale; colour signals with a single-channel dictionary (the
   channels then count as further images, pgm/ccmod.py:232-237); multi-channel dictionaries (Cd == C).
   out[] of spcsc_ccmod_step: [0] DFid = rfl2norm2(sum_m Zf Xf - Sf)/2 (pgm/ccmod.py:360-367),
   [1] Cnstr = ||Pcn(X) - X|| (:370-376), [2] Rsdl = rfl2norm2(Xf - Yfprv) (:341-345), [3] obfn_f(Yf). */
/* X = zero-padded D0 (already normalised by the caller, cbpdndl.py:448-454), Xf = Yf = rfftn(X). */
int spcsc_ccmod_reset(spcsc_handle* h, const void* D0, int32_t zero_mean);
/* setcoef: Zf = rfftn(Z) with Z the X step's current coefficient maps on this handle, device to
   device (dictlrn.py:379-383 without the host round trip): the ADMM Y or the PGM iterate X ... */
enum { SPCSC_COEF_ADMM_Y = 0, SPCSC_COEF_PGM_X = 1 };
int spcsc_ccmod_setcoef_device(spcsc_handle* h, int32_t source);
/* ... or from a host array Z (N0,N1,1,K,M).                                pgm/ccmod.py:264-281 */
int spcsc_ccmod_setcoef(spcsc_handle* h, const void* Z);
/* One PGM iteration with step 1/L and momentum coefficient coef = (t_prev - 1)/t  (pgm/pgm.py:779-831).
   flags select the statistics that cost a pass of their own (the reference computes both unless
   FastSolve is set, pgm/pgm.py:347-356): out[0] needs SPCSC_CCMOD_DFID, out[1] SPCSC_CCMOD_CNSTR. */
enum { SPCSC_CCMOD_DFID = 1, SPCSC_CCMOD_CNSTR = 2, SPCSC_CCMOD_LINSOLVE = 4, SPCSC_CCMOD_OBJ_X = 8 };
int spcsc_ccmod_step(spcsc_handle* h, double L, double coef, int32_t flags, double out[4]);
/* The same iteration in two parts for a backtracking search over L (sporco/pgm/backtrack.py:74-107 on
   pgm/ccmod.py:295-318, 379-393, pgm/pgm.py:850-894).  spcsc_ccmod_trial: proximal step at 1/L from the current
   momentum point (the gradient is computed once per iteration and kept across trials); out = { f(X) = obfn_f of the
   candidate, f(Y), <grad f(Y), X - Y>, ||X - Y||^2 }, plain sums over the stored half spectra and over all ranks.
   spcsc_ccmod_accept: ystep, residual, objective of the last candidate; out as spcsc_ccmod_step. */
int spcsc_ccmod_trial(spcsc_handle* h, double L, double out[4]);
int spcsc_ccmod_accept(spcsc_handle* h, double coef, int32_t flags, double out[4]);
/* getdict(crop=True): (hd, wd, Cd, M).                                     pgm/ccmod.py:283-291 */
int spcsc_ccmod_get_dict(spcsc_handle* h, void* D_out);
/* xstep.setdict(dstep.getdict()) on the device (dictlrn.py:386-389): Df <- Xf. */
int spcsc_ccmod_push_dict(spcsc_handle* h);
/* Multi-scale dictionaries (dsz a tuple of blocks, sporco/cnvrep.py:277-360, 609-668, 894-950): hw[2 m], hw[2 m + 1] is the
   support of filter m inside the handle's hd x wd (the largest support); the constraint projection Pcn of both dictionary
   updates then crops / zero-means / normalises every filter over its own support.  NULL: one support for all filters. */
int spcsc_ccmod_set_supports(spcsc_handle* h, const int32_t* hw);
/* ConvCnstrMOD.Xf / .Yf (pgm/ccmod.py:139-261): spectrum of the dictionary iterate (which = 0) or of the momentum point (1),
   complex, device order [Cd][N1f][M][N0]. */
int spcsc_ccmod_get_spectrum(spcsc_handle* h, int32_t which, void* out);

/* ---- consensus dictionary update: sporco.admm.ccmod.ConvCnstrMOD_Consensus (sporco/admm/ccmod.py:613-911 over
   ADMMConsensus, sporco/admm/admm.py:1419-1707) on the same handle and the same dictionary / coefficient state as
   the PGM update above (spcsc_ccmod_reset sets Y0, spcsc_ccmod_setcoef* the coefficient maps, spcsc_ccmod_get_dict /
   _push_dict read the consensus variable Y).  One block per (image, coefficient channel).
   spcsc_ccmod_cns_init: U_i = Y0 / rho when a Y0 was given, else 0 (ccmod.py:739-750); nb_global = number of
   blocks over all ranks when the images are sharded (0: this handle holds them all).
   spcsc_ccmod_cns_step: one iteration -- xstep (:787-813: solvedbi_sm per block against its coefficient spectra),
   relax_AX (admm.py:1608-1616, rlx = RelaxParam), ystep (admm.py:1585-1591 with prox_g = Pcn, ccmod.py:842-846),
   ustep -- with U read as U / udiv (the lazy form of U /= rsf after a change of rho, admm.py:549-575).
   out[]: [0] DFid on Y (ccmod.py:884-892 with fEvalX False), [1] Cnstr = ||Pcn(Y) - Y|| (:895-902) -- with
   SPCSC_CCMOD_OBJ_X (AuxVarObj False) DFid of every block with its own X_i and Cnstr of the block mean of X instead --,
   [2] ||X||^2, [3] ||X - Y||^2, [4] ||U||^2, [5] ||Y||^2, [6] ||Yprev - Y||^2 (the host forms the residuals of
   admm.py:1673-1707 from them), [7] XSlvRelRes with SPCSC_CCMOD_LINSOLVE (LinSolveCheck, ccmod.py:815-824: the solve then
   runs out of place), else -1.  Other flags as for spcsc_ccmod_step.  With images sharded over ranks the filter supports of
   the block mean and the norms are summed over the ranks (peer memory / NCCL).
   spcsc_ccmod_cns_get: the block variables X (which = 0, after the last step) or U (1) in device order
   [K*C][M][N0][N1], batch index (image, channel). */
int spcsc_ccmod_cns_init(spcsc_handle* h, double rho, int32_t y0_given, int64_t nb_global);
int spcsc_ccmod_cns_step(spcsc_handle* h, double rho, double udiv, double rlx, int32_t flags, double out[8]);
int spcsc_ccmod_cns_get(spcsc_handle* h, int32_t which, void* out);

/* ---- multi-GPU: images are sharded over ranks (one process per GPU); the only exchange of
   the path is the all-reduce of the residual / objective sums that drive the shared rho and the
   stopping test (admm/admm.py:462-486 are global over all K images).  NCCL is resolved at run
   time from `nccl_lib` (path or soname of the libnccl the process already uses).
   spcsc_comm_unique_id: rank 0 creates the 128-byte NCCL id, the caller broadcasts it.
   The caller broadcasts it and every rank creates its communicator from it. */
int spcsc_comm_unique_id(const char* nccl_lib, void* id128);
typedef struct spcsc_comm spcsc_comm;
/* Collective over all ranks (ncclCommInitRank); a communicator can serve any number of handles. */
int spcsc_comm_create(const char* nccl_lib, const void* id128, int32_t rank, int32_t nranks,
                      int32_t device, spcsc_comm** out);
int spcsc_comm_destroy(spcsc_comm* c);
/* `global_nx` = total number of coefficient elements over all ranks (sets the AbsStopTol
   scaling, admm/admm.py:481-484).  `c` must outlive the handle's iterations. */
int spcsc_attach_comm(spcsc_handle* h, spcsc_comm* c, double global_nx);

/* ---- host memory: page-locked buffers for results (direct DMA on spcsc_get_array); device
   and pinned allocations are pooled per process, spcsc_trim_pools() returns them to CUDA. */
int spcsc_host_alloc(uint64_t bytes, void** out);
int spcsc_host_free(void* p);
int spcsc_trim_pools(void);

/* ---- state access */
int spcsc_get_array(spcsc_handle* h, int32_t which, void* host_out);
int spcsc_set_array(spcsc_handle* h, int32_t which, const void* host_in);   /* Y or U */
/* GenericConvBPDN.reconstruct: irfftn(sum_m Df * rfftn(X)); X == NULL means current Y.
   out has shape (N0, N1, C, K).                                     admm/cbpdn.py:373-380 */
int spcsc_reconstruct(spcsc_handle* h, const void* X, void* out);

/* ---- level-1 entry points mirroring the reference's own unit-tested functions */
/* sporco.fft.rfftn / irfftn over the last two axes of a (batch, N0, N1) array
   (fft.py:257-314).  xf has shape (batch, N0, N1/2+1). */
int spcsc_rfft2(int32_t dtype, int32_t device, int32_t batch, int32_t N0, int32_t N1,
                const void* x, void* xf);
int spcsc_irfft2(int32_t dtype, int32_t device, int32_t batch, int32_t N0, int32_t N1,
                 const void* xf, void* x);

/* sporco.linalg.solvedbi_sm (linalg.py:232-297; Cd == 1) and solvemdbi_ism (linalg.py:370-444; Cd > 1):
   x solves (rho I + sum_c a_c a_c^H) x = b with a_c = conj(ah_c), independently for each of the nf positions and
   nk right-hand sides, the inner products running over the last axis (M).  Host arrays, complex of `dtype`:
   ah (nf, Cd, M), b and x (nf, nk, M) -- the reference's (N0, N1f, C, K, M) arrays with the leading axes flattened.
   The arithmetic is the one inside the fused column kernels (k_col / k_col2 / k_col3).   tests/test_linalg.py:147-207 */
int spcsc_solvedbi_sm(int32_t dtype, int32_t device, int64_t nf, int32_t nk, int32_t Cd, int32_t M, double rho,
                      const void* ah, const void* b, void* x);
/* sporco.prox.prox_l1 (prox/_lp.py:144-183): out = sign(v) max(|v| - alpha w, 0); v, out real (n); w: weights of the
   same shape or NULL (w = 1).                                                         tests/test_prox.py:77-96 */
int spcsc_prox_l1(int32_t dtype, int32_t device, int64_t n, double alpha, const void* w, const void* v, void* out);
/* sporco.prox.prox_sl1l2 (prox/_l21.py:51-88): prox_l1 with alpha, then the l2 shrinkage by beta of the vectors
   along the middle axis of v (n_outer, C, n_inner).                                   tests/test_prox.py:130-139 */
int spcsc_prox_sl1l2(int32_t dtype, int32_t device, int64_t n_outer, int32_t C, int64_t n_inner, double alpha,
                     double beta, const void* v, void* out);

/* sporco.signal.tikhonov_filter (signal.py:244-303), the highpass pre-processing step of the example
   scripts: every image of s (batch, N0, N1) is padded symmetrically by npd, lowpass filtered by solving
   (I + lmbda (Gr^T Gr + Gc^T Gc)) x = s in the DFT domain, cropped; sl = lowpass, sh = s - sl. */
int spcsc_tikhonov_filter(int32_t dtype, int32_t device, int32_t batch, int32_t N0, int32_t N1,
                          double lmbda, int32_t npd, const void* s, void* sl, void* sh);

#ifdef __cplusplus
}
#endif
#endif /* SPCSC_H_ */
