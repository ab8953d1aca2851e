/* cathy_b200.h -- C ABI of libcathy_b200.so, the B200-native CATHY processor.
 *
 * Boundary being replaced: pyCATHY reaches the Fortran processor only as a
 * child process `./cathy` in a project directory
 * (pyCATHY/cathy_tools.py:669 `subprocess.run(["./cathy"], ...)`,
 * pyCATHY/cathy_tools.py:76-98 `subprocess_run_multi`).  The processor's own
 * entry is PROGRAM CATHY_MAIN (SRC/cathy_main.f:2495-3945) whose time loop
 * calls FLOW3D (SRC/flow3d.f:8-35, ~200 positional arguments + COMMON blocks).
 * There is no FFI in the reference; this header is the interface a binding
 * would target.  Each entry point names the reference routine(s) it stands for.
 * SRC = pyCATHY/tests/weil_exemple/my_cathy_prj/src.
 *
 * Conventions: plain C, caller-owned host buffers, all reals are double
 * (REAL*8), all integers int32 (INTEGER*4).  Node / cell ids crossing this
 * boundary are 1-based exactly as in the CATHY files.  Return value 0 =
 * success, negative = error (text via cathy_last_error()).  A handle owns one
 * CUDA stream and all device memory of one simulation; handles share nothing,
 * so many may live in one process (ensemble members) or in many processes.
 */
#ifndef CATHY_B200_H
#define CATHY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CATHY_ABI_VERSION 8
#define CATHY_MAXIT 64 /* upper bound on ITUNS kept in a step report (CATHY.H MAXIT=30) */

/* Everything DATIN / INITAL read from the project files (SRC/datin.f:80-514,
 * SRC/atmone.f, SRC/bcone.f), already parsed into flat arrays.  Rasters are
 * row-major with the NORTH row first, exactly as they appear in the files. */
typedef struct CathyProblem {
    int32_t abi_version;
    /* --- DEM mesh (SRC/datin.f:195-204,266; SRC/triangoli.f; SRC/gen3d.f) */
    int32_t nrow, ncol, nstr, nzone, nveg;
    int32_t ivert;
    int32_t _pad0;
    double dx, dy, west, south, factor, base;
    const double *dem;       /* [nrow*ncol] cell elevations (prepro/dem)            */
    const int32_t *zone;     /* [nrow*ncol] material zone 1..nzone (prepro/zone)    */
    const double *root_map;  /* [nrow*ncol] vegetation type map (input/root_map)    */
    const double *zratio;    /* [nstr] layer thickness fractions                    */
    /* --- soil (SRC/datin.f:421-458,510-514); tables are [nstr][nzone]            */
    const double *permx, *permy, *permz, *elstor, *poros, *vgn, *vgrmc, *vgpsat;
    const double *pcana, *pcref, *pcwlt, *zroot, *pz, *omgc; /* [nveg] Feddes      */
    double pmin, scf;
    int32_t ivghu;
    /* --- parm (SRC/datin.f:80-122) */
    int32_t isimgr, kslope, lump, iopt, nlrelx, l2norm;
    int32_t ituns, ituns1, ituns2, isolv, itmxcg;
    double pondh_min, tolksl, tetaf, omega, toluns, tolswi, ernlmx, tolcg;
    double deltat, dtmin, dtmax, tmax, dtmaga, dtmagm, dtreds, dtredm;
    /* --- initial conditions (SRC/datin.f:380-403, SRC/icvhe.f, icvhwt.f, icvdwt.f) */
    int32_t indp, ipond;
    double wtposition;
    const double *ic_psi;  /* [N] used when indp = 0 or 1 (already expanded)        */
    const double *ic_pond; /* [NNOD] initial ponding heads (zeros when ipond = 0)   */
    /* --- atmospheric forcing table (SRC/atmone.f, SRC/atmnxt.f) */
    int32_t atm_none; /* 1: HSPATM = 9999 or empty file -> no atmospheric nodes     */
    int32_t hspatm;   /* 0: one value per surface node, else homogeneous            */
    int32_t ieto;     /* 0: linear interpolation in time, else piecewise constant   */
    int32_t natm;     /* number of (time, values) records                           */
    const double *atm_time; /* [natm]                                               */
    const double *atm_val;  /* [natm * (hspatm ? 1 : NNOD)] rate per unit area      */
    /* --- non-atmospheric, non-seepage Dirichlet / Neumann records (SRC/bcone.f, rdndbc.f) */
    int32_t ndir_rec, nneu_rec;
    const double *dir_time;  /* [ndir_rec]                                          */
    const int32_t *dir_ptr;  /* [ndir_rec+1] CSR style offsets into dir_node/val    */
    const int32_t *dir_node; /* 1-based 3-D node ids                                */
    const double *dir_val;   /* prescribed pressure heads                           */
    const double *neu_time;
    const int32_t *neu_ptr;
    const int32_t *neu_node;
    const double *neu_val;   /* volumetric fluxes                                   */
    const int32_t *neu_n2d;  /* [nneu_rec] NODIN2 (<0: free drainage bottom, SRC/neumann.f) */
    /* --- surface routing inputs, isimgr = 2 only (SRC/datin.f:325-372) */
    const int32_t *qoi;      /* [nrow*ncol] cells in descending elevation order (prepro/qoi_a) */
    const double *dtm_w_1, *dtm_w_2;
    const double *dtm_p_outflow_1, *dtm_p_outflow_2; /* keypad codes 1..9 stored as double */
    const double *dtm_local_slope_1, *dtm_local_slope_2;
    const double *dtm_epl_1, *dtm_epl_2;
    const double *dtm_kss1_sf_1, *dtm_kss1_sf_2;
    const double *dtm_ws1_sf_1, *dtm_ws1_sf_2;
    const double *dtm_b1_sf, *dtm_y1_sf, *dtm_nrc;
    /* --- implementation knobs (not CATHY inputs) */
    int32_t precond;    /* 0: default; see DESIGN.md                                */
    int32_t device;     /* CUDA device ordinal                                      */
    double tolcg_scale; /* multiplies TOLCG for the device PCG (<=0: 1)            */
    /* --- row-block partition of ONE mesh over several GPUs (dd_world > 1): this handle owns the global
     * node rows [dd_row0, dd_row1) of the (nrow+1) DEM node rows (row 0 = north); all arrays above stay
     * GLOBAL.  The handle builds its window (owned rows + 2 ghost node rows per interior side), and the
     * ranks exchange halo rows and reduction scalars through peer memory (cathy_dd_export/connect). */
    int32_t dd_world, dd_rank, dd_row0, dd_row1;
    /* --- moisture-curve parameters of the Huyakorn (IVGHU = 2, 3) and Brooks-Corey (IVGHU = 4) models, read from the header
     * of input/soil (SRC/datin.f:440-458; constants derived in SRC/chparm.f:79-106) */
    double hualfa, hubeta, hugama, hupsia, huswr, hun, hua, hub, bcbeta, bcrmc, bcpsat;
    /* --- stopping rule of the device linear solvers (SYMSLV / NSYSLV, SRC/solscal-extended.f:4669-4699, :3063-3240): the
     * solvers run at most ITMXCG x itmxcg_scale iterations to a relative residual of TOLCG x tolcg_scale (above).
     * <= 0 selects the documented default: itmxcg_scale 20; tolcg_scale 1 (Picard) / 1e-3 (Newton).  The effective
     * values are returned by cathy_solver_limits and printed in the header of output/iter. */
    double itmxcg_scale;
    /* --- seepage faces (input/sfbc, SRC/sfvone.f:42-66; parm line ISFONE ISFCVG DUPUIT, SRC/datin.f:110): face i owns the entries
     * [sf_ptr[i], sf_ptr[i+1]) of sf_node (1-based 3-D node ids, elevations descending along a face).  Only the node set in force
     * at time 0 is supported (a later record of sfbc must lie beyond TMAX).  Actual seepage nodes (SFEX = 1) are Dirichlet nodes at
     * psi = 0 (SRC/bcpic.f:46-58), potential ones carry zero flux; EXTALL (SRC/extall.f:32-66) moves nodes between the two sets
     * after every nonlinear iteration, ISFCVG = 1 makes an unchanged set a condition of convergence (SRC/flow3d.f:237-270). */
    int32_t nsf, isfone, isfcvg, dupuit;
    const int32_t *sf_ptr;   /* [nsf+1]                                             */
    const int32_t *sf_node;  /* [sf_ptr[nsf]]                                       */
    /* --- localized slopes KSLOPE = 3, 4 (parm line PKRL PKRR PSEL PSER, SRC/datin.f:108): inside [psel, pser] the storage term
     * uses CHPIC3's / CHPIC4's forms (SRC/chpic3.f:30-57, SRC/chpic4.f:27-45 with DSETAN of SRC/chtanp.f:22-26) */
    double psel, pser;
} CathyProblem;

/* One nonlinear iteration line of output/iter (SRC/conver.f:44 FORMAT 1070). */
typedef struct CathyIterRecord {
    int32_t niter;   /* linear iterations of this nonlinear iteration */
    int32_t ikmax;   /* 1-based node of the max-norm change           */
    double pl2, pinf, pnew_ik, pold_ik, fl2, finf;
} CathyIterRecord;

/* Everything the time loop prints for one ACCEPTED step: mbeconv FORMAT 1240
 * (SRC/cathy_main.f:3269), cumflowvol 1199 (:3677), hgatmsf 1190, dtcoupling
 * 1170 (:3686), plus control flags. */
typedef struct CathyStepReport {
    int32_t nstep, iter, nitert, kbackt, nsurf, nsurft, noback, finished;
    int32_t n_iter_rec, ponding, klsfai_total, kback_total;
    double deltat, time;
    double store1, store2, dstore;
    double vin, vout, erras, errel;
    double adin, adout, ndin, ndout, anin, anout, nnin, nnout, sfflw;
    double vsfflw, vndin, vndout, vnnin, vnnout;
    double apot, aact, ovflow, reflow;
    double fhort, fdunn, fpond, fsat;
    double next_deltat, next_time;
    double ak_max;
    double q_outlet_1, q_outlet_2; /* Q_OUT_KKP1_SN_1/2 at the outlet cell (SRC/detoutq.f) */
    double gpu_ms; /* device time of this step (CUDA events), 0 for the CPU oracle */
    int64_t launches; /* kernels launched during this step                         */
    double pcg_ms;      /* device time inside the PCG kernel over all solves of this step (incl. back-stepped attempts) */
    int64_t pcg_iters;  /* PCG iterations over all those solves */
    int64_t pcg_solves; /* number of linear solves (nonlinear iterations incl. back-stepped attempts) */
    double aact_prev;   /* AACTP: actual atmospheric flux of the previous time level (dtcoupling's AACTAV, SRC/cathy_main.f:3685) */
    double areatot;     /* AREATOT: total catchment surface area (SRC/inital.f:131-134) */
    int32_t itrtot;     /* ITRTOT: nonlinear iterations so far incl. back-stepped attempts (dtcoupling footer, SRC/cathy_main.f:3876) */
    int32_t hgflag[9];  /* HGFLAG totals so far (output/hgflag, SRC/hgraph.f) */
    CathyIterRecord it[CATHY_MAXIT];
} CathyStepReport;

typedef struct CathySim CathySim; /* opaque */

/* Sizes of the public structs as compiled, so a binding can check its mirror. */
int64_t cathy_sizeof_problem(void);
int64_t cathy_sizeof_report(void);
int32_t cathy_abi_version(void);
const char *cathy_last_error(void);

/* DATIN + GRDSYS + STRPIC/TETPIC + INITAL + CHVELO/STORCAL (SRC/cathy_main.f:2508-2660):
 * builds mesh, sparsity, scatter plan and initial state on the device. */
int32_t cathy_create(const CathyProblem *prob, CathySim **out);
void cathy_destroy(CathySim *sim);

/* Mesh counts: NNOD, N, NT, NTERM (symmetric, upper incl. diagonal), nnz of the full matrix. */
int32_t cathy_get_dims(const CathySim *sim, int64_t dims[5]);
/* GEN3D output (SRC/gen3d.f:28-77): coordinates [N] each; tetra [NT*5] in GEN3D (unsorted)
 * order, 1-based nodes + zone -- what `output/grid3d` and `output/xyz` hold.  Any pointer may be NULL. */
int32_t cathy_get_mesh(const CathySim *sim, double *x, double *y, double *z, int32_t *tetra);
/* STORE0 of SRC/cathy_main.f:2660 (initial water volume) */
double cathy_initial_storage(const CathySim *sim);

/* One pass of the time loop body SRC/cathy_main.f:2882-3829 up to and including TIMUPD:
 * BC update, surface routing, FLOW3D with back-stepping, mass balance, hydrograph terms. */
int32_t cathy_step(CathySim *sim, CathyStepReport *rep);

/* Nonlinear iterations of the FAILED attempts of the last cathy_step: the reference lists them in output/iter before those of the
 * accepted attempt, each under its own "(NSTEP: ..  DELTAT: ..  TIME: ..)" line (SRC/cathy_main.f FORMAT 1060/1065, SRC/conver.f
 * FORMAT 1070).  Attempt a (0-based, in the order they were made) ran nrec[a] iterations with time step deltat[a] ending at time[a];
 * its records are rec[a*CATHY_MAXIT .. a*CATHY_MAXIT + nrec[a]).  Returns the number of failed attempts (= kbackt of the report);
 * at most max_attempts of them are copied. */
int32_t cathy_attempt_log(CathySim *sim, int32_t max_attempts, int32_t *nrec, double *deltat, double *time, CathyIterRecord *rec);

/* State after the last accepted step (what DETOUT prints, SRC/detout.f:29-128).
 * Any pointer may be NULL.  psi,sw,ckrw,qtranie: [N]; pond,atmact,atmpot,ovfl: [NNOD]; ifatm [NNOD]. */
int32_t cathy_get_state(CathySim *sim, double *psi, double *sw, double *ckrw, double *qtranie,
                        double *pond, double *atmact, double *atmpot, double *ovfl, int32_t *ifatm);
/* The same read-back, pipelined: returns at once; the copies land in the caller's buffers (page-locked memory for full overlap)
 * while later cathy_step calls compute.  cathy_state_wait blocks until the last requested read-back is complete; a second
 * cathy_get_state_async waits (on the device) for the first to drain, so two sets of host buffers suffice to overlap every step. */
int32_t cathy_get_state_async(CathySim *sim, double *psi, double *sw, double *ckrw, double *qtranie,
                              double *pond, double *atmact, double *atmpot, double *ovfl, int32_t *ifatm);
int32_t cathy_state_wait(CathySim *sim);
/* Darcy velocities at the current state: VEL3D (SRC/vel3d.f) per element [NT] in the processor's element order (nodes of an
 * element sorted ascending under Picard, SRC/grdsys.f:63) and VNOD3D (SRC/vnod3d.f) per node [N] -- what DETOUT prints to
 * velelt / velnod and VTKRIS3D to vtk/1NN.vtk (SRC/detout.f:35, SRC/vtkris3d.f).  Any pointer may be NULL. */
int32_t cathy_get_velocity(CathySim *sim, double *uu, double *vv, double *ww, double *unod, double *vnod, double *wnod);
/* RECHARGE (SRC/recharge.f): recharge flux to the water table per surface node [NNOD] (any pointer may be NULL) and its sum
 * RECFLOW, from the nodal vertical Darcy velocity (VEL3D + VNOD3D) at the current state -- hgatmsf's REC. FLUX column and
 * output/recharge.  WTDEPTH (SRC/wtdepth.f): water-table elevation under the surface nodes nodvp[numvp] (1-based). */
int32_t cathy_get_recharge(CathySim *sim, double *recnod, double *recflow);
int32_t cathy_get_wtdepth(CathySim *sim, const int32_t *nodvp, int32_t numvp, double *wt);
/* Overwrite the pressure-head state (DA restart; stands for pyCATHY update_ic(INDP=1) +
 * relaunch, pyCATHY/cathy_tools.py:1863-1875).  Only valid before the first step. */
int32_t cathy_set_psi(CathySim *sim, const double *psi);
/* Overwrite record `rec` (0-based) of the atmospheric forcing table with new rates (stands for the
 * reference reading the next (TIME, ATMINP) record of input/atmbc as the run proceeds, SRC/atmnxt.f:34-45).
 * vals: [NNOD] when HSPATM = 0, [1] otherwise.  Host -> device copy on the handle's stream. */
int32_t cathy_upload_atm_record(CathySim *sim, int32_t rec, const double *vals);

/* ---- row-block partition (BASELINE config 5): one process per GPU, one handle per process --------
 * After every rank has created its handle: each exports an opaque 64-byte CUDA IPC handle of its
 * communication box, the caller gathers them (torch.distributed all_gather) and every rank connects.
 * From then on cathy_step must be called by ALL ranks for every step (the kernels rendezvous). */
int32_t cathy_dd_export(CathySim *sim, void *handle64);
int32_t cathy_dd_connect(CathySim *sim, const void *handles /* [dd_world][64] */);
/* Same-process variant (one host thread per handle): wire the ranks by direct pointers, then let every thread call
 * cathy_dd_start concurrently (the collective part of the set-up). */
int32_t cathy_dd_connect_local(CathySim *sim, CathySim *const *all /* [dd_world] */);
int32_t cathy_dd_start(CathySim *sim);
/* info[0..7] = window start (global node row), window rows, owned global rows [a,b), local NNOD, local N,
 * global NNOD, global N. */
int32_t cathy_dd_info(const CathySim *sim, int64_t info[8]);

/* Which linear-solver kernel this handle launches (for bench.py's roofline bookkeeping; no reference counterpart):
 * info[0] = 1 k_pcg (CG vectors streamed), 2 k_pcg2, 3 k_pcg_res / 4 k_pcg_res2 (CG vectors resident in shared memory), 5 k_pcg / 6 k_pcg_tma on the
 * column-major permutation, 7 k_pcg_cl / 8 k_pcg_cl2 (small meshes: one thread-block cluster, matrix and vectors in shared memory; 8 = one cluster barrier per iteration), 10 k_bicgstab / 11 k_bicgstab_res (Newton);
 * info[1] = rows per CTA (k_pcg_res), info[2] = 1 if the solution vector is resident too, info[3] = CTAs of the solver grid. */
int32_t cathy_solver_info(const CathySim *sim, int64_t info[4]);
/* Effective stopping rule of this handle's linear solver: lim[0] = iteration limit (ITMXCG x itmxcg_scale), lim[1] = relative
 * residual tolerance (TOLCG x tolcg_scale), lim[2] = itmxcg_scale, lim[3] = tolcg_scale as applied, lim[4] = preconditioner
 * (1 diagonal, 2 vertical line).  The reference's ISOLV preconditioners IC(0) / ILU(0) (SRC/solscal-extended.f:2142-2267) are sequential
 * sweeps; the device uses parallel ones and therefore other iteration counts -- LSFAIL keeps its meaning. */
int32_t cathy_solver_limits(const CathySim *sim, double lim[5]);
/* How the assembly (ASSPIC, SRC/asspic.f:26-51) finds the elements of a matrix entry: info[0] = 1 when the tet indices of the gather
 * plan are derived from the mesh structure (verified against the stored lists at cathy_create), 0 when they are read from memory;
 * info[1] = width of the per-class offset tables. */
int32_t cathy_plan_info(const CathySim *sim, int64_t info[2]);

/* ---- kernel-level entry points used by parity tests and bench.py -------------------- */
/* Assemble the Picard system at the current state for time step `deltat` without solving
 * (PICUNS+ASSPIC+RHSPIC+CFMATP+RHSGRV+BCPIC, SRC/picard.f:74-154) and export it as
 * symmetric upper CSR in the reference layout (diagonal first; SRC/strpic.f:19-96).
 * topol [N+1], ja [NTERM] are 1-based; coef1 [NTERM]; rhs [N].  Any pointer may be NULL. */
int32_t cathy_debug_assemble(CathySim *sim, double deltat, int32_t *topol, int32_t *ja,
                             double *coef1, double *rhs);
/* y = A x with the currently assembled LHS (the SpMV of GRADDP, SRC/solscal-extended.f:1326-1334).
 * Runs `reps` launches, returns average device ms per launch in *ms (may be NULL). */
int32_t cathy_debug_spmv(CathySim *sim, const double *x, double *y, int32_t reps, double *ms);
/* Solve the currently assembled system (SYMSLV, SRC/solscal-extended.f:4669-4699).
 * sol [N]; niter, err = relative residual as GRADDP defines it. */
int32_t cathy_debug_solve(CathySim *sim, double *sol, int32_t *niter, double *err, double *ms);

/* ---- ensemble data assimilation: dense analysis update (fp64 tensor cores) ------------------
 * Reference arithmetic: pyCATHY/DA/enkf.py:16-224 (enkf_analysis), :225-342
 * (enkf_analysis_localized_with_inflation); pyCATHY/DA/pf.py:3-110,197-211 (particle filter
 * weights + systematic resampling).  Ensemble matrices are row-major [n_state][n_ens], the
 * orientation the reference uses (ensemble.shape = (N_state, N_ens)).
 *
 * Notation: X [n][ne] augmented state (states, then parameters), HX [m][ne] predicted
 * observations, y [m] (or [m][ne] when y_is_matrix), R [m][m] observation error covariance,
 * L [n_loc][m] localisation (Schur) factors applied to the first n_loc rows of the cross
 * covariance (NULL: none), inflate: multiplicative inflation about the analysis mean applied to the
 * first n_infl rows (1.0: none). */
const char *cathy_enkf_last_error(void);
/* Whole analysis with HOST buffers (what DA.run_analysis calls, pyCATHY/DA/cathy_DA.py:86-260):
 * Xa = X + [(X - mean)(HX - mean)^T/(ne-1) o L] B,  B = C^-1 (y - HX) or (y - HX)/diag(R) (sakov).
 * Outputs (any may be NULL): Xa [n][ne], B [m][ne] (inv_data_pert), P [n][m] (ensemble_pert). */
int32_t cathy_enkf_analysis_host(const double *X, int64_t n, int32_t ne, const double *HX, const double *y,
                                 int32_t y_is_matrix, const double *R, int32_t m, int32_t sakov,
                                 const double *L, int64_t n_loc, double inflate, int64_t n_infl,
                                 double inflate2 /* rows n_infl..n, e.g. parameters */,
                                 double *Xa, double *B, double *P, int32_t device, double *device_ms);
/* Stage 1 (tiny, host buffers): S = HX - mean (obs_pert) and B (inv_data_pert); enkf.py:139-171. */
int32_t cathy_enkf_gain(const double *hx, const double *y, int32_t y_is_matrix, const double *R, int32_t m,
                        int32_t ne, int32_t sakov, double *S_out, double *B_out);
/* Stages on DEVICE pointers (member-sharded ensembles; the caller all-reduces rowsum and P between
 * stages).  `stream` is a cudaStream_t passed as an integer (0 = legacy default stream). */
int32_t cathy_enkf_rowsum(const double *dX, int64_t n, int32_t ne, double *d_rowsum, uint64_t stream);
/* d_mean = rowsum / ne_total is formed by the caller after the reduction (or by cathy_enkf_scale). */
int32_t cathy_enkf_scale(double *d_v, int64_t n, double a, uint64_t stream);
/* partial cross covariance of the local members: P = (X - mean) S_local^T / (ne_total - 1); enkf.py:182 */
int32_t cathy_enkf_crosscov(const double *dX, const double *d_mean, const double *dS_local, int64_t n,
                            int32_t ne_local, int32_t m, int32_t ne_total, double *dP, uint64_t stream);
/* Xa = X + (P o L) B_local, then inflation about mean_a = mean + (P o L) bbar; enkf.py:197, :317-324.
 * inflate applies to rows [0,n_infl), inflate2 to rows [n_infl,n). dXa may alias dX. */
int32_t cathy_enkf_update(const double *dX, const double *dP, const double *dL, int64_t n_loc,
                          const double *dB_local, const double *d_mean, const double *d_bbar, double inflate,
                          int64_t n_infl, double inflate2, int64_t n, int32_t ne_local, int32_t m, double *dXa,
                          uint64_t stream);
/* Gaspari-Cohn localisation matrix (pyCATHY/DA/localisation.py:136-188 gaspari_cohn, build_localization_matrix):
 * L[i][k] = gc(|grid_i - obs_k| / radius) with 2-D positions; DEVICE pointers grid_xy [n][2], obs_xy [m][2], L [n][m]. */
int32_t cathy_enkf_localization(const double *d_grid_xy, int64_t n, const double *d_obs_xy, int32_t m, double radius,
                                double *dL, uint64_t stream);
/* Particle filter (pf.py:60-110): normalised weights [ne] and n_eff from HX [m][ne], y [m], obs_std [m];
 * host buffers. */
int32_t cathy_pf_weights(const double *hx, const double *y, const double *obs_std, int32_t m, int32_t ne,
                         double *weights, double *n_eff);
/* Systematic resampling (pf.py:197-211) with the caller's uniform draw u in [0,1): indices [ne] (0-based). */
int32_t cathy_pf_systematic_resample(const double *weights, int32_t ne, double u, int32_t *indices);
/* Xout[:, j] = X[:, idx[j]] on the device (ensemble[:, indices], pf.py:104-106); dXout must not alias dX. */
int32_t cathy_pf_gather_members(const double *dX, int64_t n, int32_t ne, const int32_t *d_idx, double *dXout,
                                uint64_t stream);
/* Move one member's pressure heads between its simulation handle and column `col` of a device-resident
 * ensemble matrix [n][ld] (stands for DA._read_state_ensemble / update_ENS_files' text round trip,
 * pyCATHY/DA/cathy_DA.py:2684, :1863-1875).  which: 0 = psi, 1 = sw. */
int32_t cathy_pack_state(CathySim *sim, int32_t which, double *dX, int64_t ld, int64_t col);
int32_t cathy_unpack_psi(CathySim *sim, const double *dX, int64_t ld, int64_t col);
/* Begin a new run at time 0 from the CURRENT pressure heads with a new TMAX / first DELTAT (<= 0: keep):
 * the relaunch of the processor between assimilation windows (pyCATHY/cathy_tools.py:593-740 after
 * update_ic + update_parm, pyCATHY/DA/cathy_DA.py:1863-1875). */
int32_t cathy_restart(CathySim *sim, double tmax, double deltat);
/* Replace the soil tables ([nstr][nzone] each; SRC/datin.f:510-514; pyCATHY/cathy_tools.py update_soil):
 * the parameter part of the analysis.  Host buffers. */
int32_t cathy_set_soil(CathySim *sim, const double *permx, const double *permy, const double *permz,
                       const double *elstor, const double *poros, const double *vgn, const double *vgrmc,
                       const double *vgpsat);
/* Replace the atmospheric forcing table (input/atmbc rewrite per window; SRC/atmone.f).  Host buffers:
 * times [natm], vals [natm * (HSPATM ? 1 : NNOD)].  Effective at the next cathy_restart. */
int32_t cathy_set_atm_table(CathySim *sim, int32_t natm, const double *times, const double *vals);

#ifdef __cplusplus
}
#endif
#endif /* CATHY_B200_H */
