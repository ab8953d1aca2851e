// host_setup.cuh -- host side, set-up: device buffers, the simulation handle, launch macro, mesh / sparsity / static gather plan (GRDSYS, STRPIC, TETPIC replaced by the plan), surface routing tables.
// Part of the single translation unit cathy_b200.cu (included in dependency order; shares its structs and helpers).
#pragma once

// ==========================================================================================
// host side
// ==========================================================================================
template <class T>
struct DBuf {
    T *p = nullptr;      // logical element 0
    T *base = nullptr;   // allocation start (p - pad)
    size_t n = 0, pad = 0;
    // `halo` zero-filled elements are kept on both sides so stencil kernels can gather without bounds checks
    int alloc(size_t cnt, size_t halo = 0)
    {
        n = cnt; pad = halo;
        size_t tot = std::max<size_t>(cnt + 2 * halo, 1);
        if (cudaMalloc((void **)&base, tot * sizeof(T)) != cudaSuccess) return -1;
        p = base + halo;
        return cudaMemset(base, 0, tot * sizeof(T)) == cudaSuccess ? 0 : -1;
    }
    int upload(const std::vector<T> &h, size_t halo = 0)
    {
        if (base && n == h.size() && pad == halo) {   // refresh of an existing table (cathy_set_soil)
            if (h.empty()) return 0;
            return cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice) == cudaSuccess ? 0 : -1;
        }
        release();
        if (alloc(h.size(), halo)) return -1;
        if (h.empty()) return 0;
        return cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice) == cudaSuccess ? 0 : -1;
    }
    void release() { if (base) cudaFree(base); base = p = nullptr; }
};

// host bookkeeping of one nansfdirbc / nansfneubc record stream: the three-slot window of BCONE/BCNXT/BCBAK
struct HostBc {
    int nrec = 0;
    std::vector<double> time, val;
    std::vector<int> ptr, node, n2d;
    int slot[3] = {-1, -1, -1};
    double tim[3] = {0, 0, 0};
    int next = 0, hti = 0, active = -2;   // active: record currently loaded on the device (-1 none, -2 never)
    int anbc() const { return active >= 0 ? ptr[active + 1] - ptr[active] : 0; }
};

struct DDComm {
    void *base = nullptr;            // [DDBox][inbox 2 x 2 x hcap doubles]
    size_t bytes = 0;
    void *peer_base[DD_MAXW] = {nullptr};
    bool opened[DD_MAXW] = {false};
    bool connected = false;
    DDCtx ctx;
    unsigned int *seq = nullptr;     // device [2]
    int *err = nullptr;              // device [1]
    unsigned int *recv_counter = nullptr;
    cudaIpcMemHandle_t handle;
};

struct CathySim {
    CathyProblem p;
    HostBc dir, neu;
    bool have_dir = false, have_neu = false, free_drain = false, bc_any = false;
    DBuf<unsigned char> contp_flag, contq_flag;
    DBuf<double> contp_val, qneu, qlist, qpnew, qpold, kznod, bcsum;
    DBuf<int> contp_list;
    double ndin = 0, ndout = 0, nnin = 0, nnout = 0, vndin = 0, vndout = 0, vnnin = 0, vnnout = 0;
    // seepage faces (seepage.cuh): flattened node list and its per-node state
    DBuf<RelxPartial> relx_part;      // NLRELX = 2 (RELXOM): block partials, {OMEGA, OMEGAP}
    DBuf<double> d_omega;
    DBuf<double> ptold;      // previous nonlinear iterate of PTNEW, kept for the chord slopes (KSLOPE = 1, 2)
    int sf_n = 0, sfchek = 0, ksfzer = 1, ksfcv = 0, ksfcvt = 0;
    DBuf<int> sf_node, sf_ex, sf_exp, sf_exit;
    DBuf<double> sf_q, sf_qp;
    DBuf<SfOut> d_sf;
    double sfflw = 0, sfflwp = 0, vsfflw = 0;
    // dense Dirichlet flag / value arrays as the kernels see them: prescribed-head nodes (bit 0) and actual seepage nodes (bit 1)
    const unsigned char *flagp() const { return (have_dir || sf_n > 0) ? contp_flag.p : nullptr; }
    const double *valp() const { return (have_dir || sf_n > 0) ? contp_val.p : nullptr; }
    int nrow, ncol, nc1, nstr, nnod, n, ntri, nt, ncell;
    bool surf;
    cudaStream_t st = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evp0 = nullptr, evp1 = nullptr;
    double pcg_ms = 0;
    int64_t pcg_iters = 0, pcg_solves = 0;
    int sms = 148, grid_n = 0, grid_pcg = 0, pcg_block = 1024, pcg_custom = 1, pcg_minb = 0, pcg_prefetch = 1;
    int pcg_cluster = 0;                     // > 0: k_pcg_res2 runs as ONE thread-block cluster of that many CTAs (small meshes)
    int pcl_block = 256;                     // threads per CTA of k_pcg_cl2 (CATHY_PCG_CL_BLOCK)
    bool route_lanes4 = false;               // CATHY_ROUTE_LANES=4: k_route4 (four lanes per cell) instead of k_route -- measured slower, see k_route4
    // CUDA-graph replay of one Picard iteration (small meshes, see picard_iteration): [0] = later iterations of a step, [1] = the first
    // (it also evaluates Sw at the previous time level); the step-dependent scalars {DELTAT, 1/DELTAT} are read from d_dt
    int graph_mode = 0, graph_capturing = 0;
    cudaGraphExec_t gexec[2] = {nullptr, nullptr};
    int64_t glaunches[2] = {0, 0};
    DBuf<double> d_dt;
    double *h_dt = nullptr, dt_uploaded = -1.0;
    struct HostReadback { SfOut sf; int pond; int pad; double bc[4]; } *h_rb = nullptr;      // page-locked targets of the per-iteration read-backs
    const double *graph_dt() const { return graph_capturing ? d_dt.p : nullptr; }
    void graph_drop() { for (auto &g : gexec) { if (g) cudaGraphExecDestroy(g); g = nullptr; } }
    int pcl_c = 0, pcl_rows = 0, pcl_v2 = 0; // > 0: k_pcg_cl / k_pcg_cl2 (pcg_cluster.cuh): cluster size, rows per CTA, single-barrier variant
    size_t pcl_smem = 0;
    unsigned int barrier_epoch = 0;
    cudaStream_t st_copy = nullptr;          // cathy_get_state_async: drain stream, snapshot buffers
    cudaEvent_t ev_snap = nullptr, ev_drained = nullptr;
    DBuf<double> snap;
    DBuf<int> snap_i;
    DBuf<unsigned int> d_counter;
    int64_t launches = 0;
    cudaError_t launch_err = cudaSuccess;    // first failed kernel launch (LAUNCH macro), reported by launch_check()
    const char *launch_err_kernel = "";
    int launch_err_line = 0;
    // host mesh kept for export
    std::vector<double> hx, hy, hz, harenod;
    std::vector<double> h_dem, h_root, h_zratio, h_veg;   // owned copies of the caller's mesh inputs (cathy_set_soil rebuilds from them)
    std::vector<int32_t> h_zone;
    CurveModel cm;                  // Huyakorn / Brooks-Corey constants (ivghu = 0: unused)
    double areatot = 0.0;           // AREATOT (SRC/inital.f:131-134), sequential sum over the (global) surface nodes
    std::vector<double> h_perm;     // permx | permy | permz tables as last built ([nstr][nzone] each)
    std::vector<int> htri;       // [ntri*4] sorted nodes + zone
    std::vector<unsigned char> hexist; // [NDIAG*n] structural mask of the upper diagonals
    int64_t nterm = 0;
    int off[NDIAG];
    // static device data
    DBuf<double> vgn, vgm, vgpsat, vgpnot, rr, snodi, pnodi, vgn1, vgnr, vgpsn, vgmr, vgm52, vgmm1, volnod, arenod, z, m4, vegpar;
    DBuf<int> veg;
    DBuf<int4> tet;
    DBuf<unsigned char> ell_loc; // Newton: (local row node | local column node << 2) of every diagonal-family ELL entry
    DBuf<double> tet_k0, tet_gz, tet_vol;   // Newton: per-tet unit-kr stiffness [10][nt], Kz*IVOL*d_k [4][nt], volume [nt]
    size_t fam_off[NDIAG] = {0};
    DBuf<int> ell_tet;           // ELL-transposed gather lists (see k_assemble)
    DBuf<int> plan_rel;          // k_assemble_a: 27 classes x NDIAG x wrel tet offsets
    PlanGeom geom{};             // rel == nullptr: stored indices (k_assemble)
    DBuf<double> ell_coef, ell_coef2;
    EllPlan plan;
    size_t ld = 0, halo = 0;     // leading dimension of the diagonals / halo of gathered vectors
    DBuf<StepPartial> spart;
    // matrices / vectors
    DBuf<double> A;              // 8 diagonals, [NDIAG][n]
    DBuf<double> diag_true, diag_bc, grav, m2, krt, e1t;
    DBuf<double> pnew, pold, ptimep, ptnew, pdiff, sw, ckrw, ckrwp, et1, et2, swnew, swtimep, rhs, xt5, qtranie;
    DBuf<double> wr, wz, wp0, wp1, wbv, partial, store_part;
    DBuf<double> dis, wq0, wq1;      // k_pcg2: 1/sqrt(diag), two more work vectors
    bool scaled = false;             // off-diagonals of A currently hold the symmetrically scaled matrix
    int pcg_algo = 4;                // 4: k_pcg_res2 (CG vectors resident in shared memory, paired rows; default, falls back to 1 when they do not fit),
                                     // 3: k_pcg_res (first resident version, one row per thread),
                                     // 1: k_pcg (vectors streamed from HBM/L2), 2: k_pcg2 (scaled, single reduction); CATHY_PCG_ALGO
    int res_rows = 0, res_x = 0, res_prefetch = 0;   // k_pcg_res: rows per CTA (0 = does not fit), x resident too, L2 prefetch of the diagonals
    int bicg_line = 0;               // Newton: 0 = point Jacobi (default), 1 = vertical-line preconditioner (opt-in, CATHY_BICG_LINE=1: -36 % iterations but
                                     // +46 % per iteration on the config-3 storm, no gain on unsaturated systems; profiles/r1_precond_experiment.md)
    DBuf<double> widn, wcp;          // its Thomas factors
    bool l2_reset = true;
    size_t l2_window = 0, l2_persist = 0, l2_maxwin = 0;   // bytes of the Jacobian covered by the access-policy window / L2 set-aside for persisting lines
    DBuf<double> Ju, Jl, dinv, dckrw, detai, ts, s1, ws, wsh, wt;   // Newton: Jacobian diagonals, Jacobi scaling, derivative curves, element factors, BiCGSTAB vectors
    // Picard, streaming PCG in the column-major permutation (meshes too large for the resident kernels): cm_on, permuted arrays
    bool cm_on = false;
    int cm_off[NDIAG] = {0};
    size_t cm_halo = 0;
    DBuf<double> cm_A, cm_diag, cm_rhs, cm_x, cm_r, cm_z, cm_p0, cm_p1, cm_bv;
    bool tma_on = false;             // k_pcg_tma (pcg_tma.cuh) instead of k_pcg on the permuted arrays; cm_p1 holds the reciprocal diagonal
    size_t tma_smem = 0;
    double *tma_zpeer_n = nullptr, *tma_zpeer_s = nullptr;
    long long tma_ndst0 = 0, tma_sdst0 = 0;
    bool newton = false;
    // Newton, resident solver (bicg_res.cuh): permuted Jacobian + vectors, line factors; bres_rows = 0: not used (does not fit / opted out)
    int bres_rows = 0, bres_cols = 0, bres_off[NDIAG] = {0};
    size_t bres_halo = 0;
    DBuf<double> bres_u, bres_l, bres_rhs, bres_dinv, bres_x, bres_ph, bres_sh, bres_rt, bres_p;
    size_t bres_smem = 0;
    DBuf<unsigned char> bres_symf;           // k_bres_sym_flags: one byte per (CTA, pass, warp) group of 64 rows
    DBuf<unsigned long long> bres_prof;      // CATHY_BRES_PROF=1: per-phase nanoseconds of CTA 0, printed at cathy_destroy
    // ---- row-block partition of one large mesh over several GPUs (BASELINE config 5) ----
    bool dd = false, pcg_shared_gpu = false;
    int dd_world = 1, dd_rank = 0;
    int gnrow = 0;            // global number of DEM rows
    int grow0 = 0;            // global node row of local node row 0 (window start)
    int own_a = 0, own_b = 0; // owned LOCAL node rows [own_a, own_b)
    int gnnod = 0;            // global surface node count
    std::vector<double> ovr_z; std::vector<int> ovr_veg; double ovr_zmin = 0.0;
    DBuf<unsigned char> own;  // [n] 1 = row owned by this rank (reductions count owned rows only)
    struct DDComm *comm = nullptr;
    DBuf<NormPartial> npart;
    DBuf<IterOut> d_iter;
    DBuf<StepOut> d_step;
    IterOut *h_iter = nullptr;
    StepOut *h_step = nullptr;
    DBuf<int> ifatm, ifatmp, d_flags; // d_flags[0]=ponding, [1]=etran error
    DBuf<double> atmpot, atmact, atmold, atmtab, pondnod, ovflnod, ovflp, scal3;
    // atmospheric stream (host bookkeeping of the three-slot window, SRC/atmone.f / atmnxt.f)
    double atmtim[3] = {0, 0, 0};
    int atmrec[3] = {-1, -1, -1};
    int atm_next = 0, htiatm = 0;
    // surface routing
    DBuf<int> lv_ptr, lv_cell, seqpos, don_ptr, don_cell, don_code;
    DBuf<unsigned char> don_dir;
    DBuf<double> r_w1, r_w2, r_sl1, r_sl2, r_epl1, r_epl2, r_ks1, r_ks2, r_ws1, r_ws2, r_b1, r_y1, r_nrc, r_ckf1, r_ckf2, r_dhd1, r_dhd2;
    bool route_static_done = false;
    DBuf<RouteS> r_rs;               // k_route_wave: per-cell records in level order, overflow donor codes, histories
    DBuf<int> r_dcx, r_handled;
    DBuf<double> r_qo, r_qin_ring, r_vol_ring, r_best;
    DBuf<unsigned long long> r_prof;
    bool route_wave = false;
    int route_last_nsurf = 1;        // sub-steps of the previous routing call (sizes the next launch)
    int route_cluster = 8;           // CTAs of the k_route_wave cluster (16 when the device allows the non-portable size)
    DBuf<double> sw_sn, q_in_kk, q_in_kkp1, q_out_kk_1, q_out_kk_2, q_out_kkp1_1, q_out_kkp1_2, volume_kk, volume_kkp1, h_water;
    DBuf<double> q_in_kk_sav, q_out_kk_1_sav, q_out_kk_2_sav, volume_kk_sav, q_in_kk_p, q_out_kk_1_p, q_out_kk_2_p, volume_kk_p;
    DBuf<double> d_akmax;   // [3]: ak_max, ak_max_p, ak_max_sav
    DBuf<int> d_nsurf;
    int nlevel = 0, outlet_cell = 0;
    // time stepping state (host)
    double time = 0, timep = 0, deltat = 0, dtmin = 0, dtmax = 0, tmax = 0, tetaf = 1;
    int dtgmin = 1, nstep = 1, iter = 1, nitert = 0, itlin = 0, itrtot = 0, kbackt = 0, kback = 0, klsfai = 0, nsurft = 0;
    int finished = 0, lsfail = 0, ponding = 0, pondp = 0, timep_dirty = 1;
    double adinp = 0, adoutp = 0, ndinp = 0, ndoutp = 0, aninp = 0, anoutp = 0, nninp = 0, nnoutp = 0, aactp = 0;
    double adin = 0, adout = 0, anin = 0, anout = 0, vin = 0, vout = 0, dstore = 0, erras = 0, errel = 0;
    double store0 = 0, store1 = 0, store2 = 0;
    int hgflag[9] = {0};
    CathyIterRecord itrec[CATHY_MAXIT];
    struct Attempt { double deltat, time; int n; CathyIterRecord rec[CATHY_MAXIT]; };
    std::vector<Attempt> attempts;   // failed attempts of the step being made (cathy_attempt_log)
    int itmax_dev = 0;
    double tol_dev = 0, itmxcg_scale = 0, tolcg_scale = 0;
};

static inline int nblk(long long n, int cap) { long long b = (n + RED_BLOCK - 1) / RED_BLOCK; return (int)std::max<long long>(1, std::min<long long>(b, cap)); }
// a launch that fails for a non-sticky reason (bad configuration, too many resources) must not pass silently: the first such error
// is kept in the handle and turned into a failed cathy_step / cathy_create by launch_check()
#define LAUNCH(S, kern, grid, block, ...)                                   \
    do {                                                                    \
        kern<<<(grid), (block), 0, (S)->st>>>(__VA_ARGS__);                 \
        (S)->launches++;                                                    \
        cudaError_t le_ = cudaPeekAtLastError();                            \
        if (le_ != cudaSuccess && (S)->launch_err == cudaSuccess) {         \
            (S)->launch_err = le_; (S)->launch_err_kernel = #kern; (S)->launch_err_line = __LINE__; \
            cudaGetLastError();                                             \
        }                                                                   \
    } while (0)

static int launch_check(CathySim *S)
{
    if (S->launch_err == cudaSuccess) return 0;
    FAIL(-100, "kernel launch %s failed (%s, %s:%d)", S->launch_err_kernel, cudaGetErrorString(S->launch_err), __FILE__, S->launch_err_line);
}
static Diag make_diag(CathySim *S, double *base)
{
    Diag D;
    for (int d = 0; d < NDIAG; ++d) { D.d[d] = base + (size_t)d * S->ld; D.off[d] = S->off[d]; }
    return D;
}
static Soil make_soil(CathySim *S)
{
    Soil s;
    s.vgn = S->vgn.p; s.vgm = S->vgm.p; s.vgpsat = S->vgpsat.p; s.vgpnot = S->vgpnot.p; s.rr = S->rr.p; s.snodi = S->snodi.p;
    s.pnodi = S->pnodi.p; s.vgn1 = S->vgn1.p; s.vgnr = S->vgnr.p; s.vgpsn = S->vgpsn.p; s.vgmr = S->vgmr.p;
    s.vgm52 = S->vgm52.p; s.vgmm1 = S->vgmm1.p;
    return s;
}

// ---- host mesh + static tables -----------------------------------------------------------
static void sort4(int *e)
{
    for (int k = 0; k < 3; ++k) for (int j = k + 1; j < 4; ++j) if (e[k] > e[j]) std::swap(e[k], e[j]);
}
static void gen_tets_of_prism(const int *tri, int top, int bot, int out[3][4])
{   // SRC/gen3d.f:31-45 (0-based)
    out[0][0] = top + tri[0]; out[0][1] = top + tri[1]; out[0][2] = top + tri[2]; out[0][3] = bot + tri[0];
    out[1][0] = bot + tri[0]; out[1][1] = bot + tri[1]; out[1][2] = bot + tri[2]; out[1][3] = top + tri[2];
    out[2][0] = top + tri[1]; out[2][1] = top + tri[2]; out[2][2] = bot + tri[1]; out[2][3] = bot + tri[0];
}

static int build_static(CathySim *S)
{
    const CathyProblem &p = S->p;
    const int nrow = S->nrow, ncol = S->ncol, nc1 = S->nc1, nnod = S->nnod, n = S->n, nstr = S->nstr, ntri = S->ntri;
    const size_t nt = (size_t)S->nt;
    // --- surface mesh (SRC/triangoli.f, SRC/tpnodi2d.f, SRC/area2d.f)
    S->hx.assign(n, 0.0); S->hy.assign(n, 0.0); S->hz.assign(n, 0.0); S->harenod.assign(nnod, 0.0);
    S->htri.resize(4 * (size_t)ntri);
    std::vector<int> cnt(nnod, 0);
    for (int i = 0; i <= nrow; ++i)
        for (int j = 0; j <= ncol; ++j) {
            int k = i * nc1 + j;
            S->hx[k] = p.west + j * p.dx;
            S->hy[k] = S->dd ? p.south + (S->gnrow - (i + S->grow0)) * p.dy : p.south + (nrow - i) * p.dy;   // row-block window: global row index
        }
    for (int i = 0, it = 0; i < nrow; ++i)
        for (int j = 0; j < ncol; ++j) {
            int n00 = i * nc1 + j, n10 = n00 + nc1, n11 = n10 + 1, n01 = n00 + 1, zn = p.zone[i * ncol + j];
            double e = p.dem[i * ncol + j] * p.factor;
            int t1[3] = {n00, n10, n11}, t2[3] = {n00, n11, n01};
            for (int q = 0; q < 3; ++q) { S->hz[t1[q]] += e; cnt[t1[q]]++; }
            for (int q = 0; q < 3; ++q) { S->hz[t2[q]] += e; cnt[t2[q]]++; }
            int *a = &S->htri[4 * (size_t)it++]; a[0] = n00; a[1] = n10; a[2] = n11; a[3] = zn;
            int *b = &S->htri[4 * (size_t)it++]; b[0] = n00; b[1] = n01; b[2] = n11; b[3] = zn;
        }
    for (int k = 0; k < nnod; ++k) S->hz[k] /= cnt[k];
    if (S->dd) for (int k = 0; k < nnod; ++k) S->hz[k] = S->ovr_z[k];   // node elevations from the GLOBAL DEM (window edges lack cells)
    for (int t = 0; t < ntri; ++t) {
        const int *T = &S->htri[4 * (size_t)t];
        double a3 = 0, a2 = 0;
        for (int ii = 0; ii < 3; ++ii) {
            int I = T[ii], J = T[(ii + 1) % 3], M = T[(ii + 2) % 3];
            a3 = S->hx[I] * S->hy[J] + a3; a2 = S->hx[I] * S->hy[M] + a2;
        }
        double are3 = std::fabs(0.5 * (a3 - a2)) * (1.0 / 3.0);
        S->harenod[T[0]] += are3; S->harenod[T[1]] += are3; S->harenod[T[2]] += are3;
    }
    S->areatot = 0.0;
    for (int k = 0; k < nnod; ++k) S->areatot = S->areatot + S->harenod[k];
    // --- vertical discretisation (SRC/gen3d.f:52-77)
    double zmin = RMAX_;
    for (int i = 0; i < nnod; ++i) zmin = std::min(zmin, S->hz[i]);
    if (S->dd) zmin = S->ovr_zmin;
    for (int i = 0; i < nnod; ++i) {
        double zthick = (S->hz[i] - zmin) + p.base, zrsum = 0.0;
        for (int j = 1; j <= nstr; ++j) {
            size_t kk = (size_t)j * nnod + i;
            S->hx[kk] = S->hx[i]; S->hy[kk] = S->hy[i];
            zrsum = zrsum + p.zratio[j - 1];
            double zz;
            switch (p.ivert) {
            case 0: zz = S->hz[i] - zrsum * p.base; break;
            case 1: zz = S->hz[i] - zrsum * zthick; break;
            case 2: zz = zmin - zrsum * p.base; break;
            default: zz = S->hz[i] - zrsum * p.base; if (j == nstr) zz = zmin - p.base; break;
            }
            S->hz[kk] = zz;
        }
    }
    // --- stencil offsets of the 8 upper diagonals
    int offs[NDIAG] = {0, 1, nc1, nc1 + 1, nnod - nc1 - 1, nnod - nc1, nnod - 1, nnod};
    for (int d = 0; d < NDIAG; ++d) S->off[d] = offs[d];
    if (!(nc1 + 1 < nnod - nc1 - 1)) FAIL(-3, "DEM too small for the diagonal layout (need at least 2 rows)");
    auto diag_of = [&](int dlt) -> int { for (int d = 0; d < NDIAG; ++d) if (offs[d] == dlt) return d; return -1; };
    // --- per-tet geometry, nodal soil averages, contribution lists
    std::vector<int4> tet(nt);
    std::vector<double> volnod(n, 0.0), pnodi(n, 0.0), snodi(n, 0.0), vgn(n, 0.0), vgrmc(n, 0.0), vgpsat(n, 0.0), kznod(n, 0.0);
    std::vector<int> tp(n, 0);
    const size_t nslots = (size_t)NDIAG * n;
    std::vector<int> s_cnt(nslots + 1, 0), n_cnt(n + 1, 0);
    struct TetGeo { double c[10]; double g[4]; double vol; };
    // pass 1: geometry is recomputed in pass 2 to keep memory low; here only counts + nodal sums
    auto tet_nodes = [&](size_t e, int T[4]) {
        size_t lay = e / ((size_t)ntri * 3), rem = e - lay * (size_t)ntri * 3;
        int tri = (int)(rem / 3), which = (int)(rem % 3), pr[3][4];
        gen_tets_of_prism(&S->htri[4 * (size_t)tri], (int)lay * nnod, ((int)lay + 1) * nnod, pr);
        for (int q = 0; q < 4; ++q) T[q] = pr[which][q];
        if (p.iopt == 1) sort4(T);
    };
    static const double amen[5] = {-1.0, 1.0, -1.0, 1.0, -1.0};
    auto geometry = [&](const int T[4], double b[4], double c[4], double d[4], double &vol) {
        const double *X = S->hx.data(), *Y = S->hy.data(), *Z = S->hz.data();
        vol = 0.0;
        for (int nn = 0; nn < 4; ++nn) {
            int o3[3] = {(nn + 1) & 3, (nn + 2) & 3, (nn + 3) & 3};
            double a2, a3;
            a2 = a3 = 0.0;
            for (int ii = 0; ii < 3; ++ii) { int I = T[o3[ii]], J = T[o3[(ii + 1) % 3]], M = T[o3[(ii + 2) % 3]]; a3 = Y[I] * Z[J] + a3; a2 = Y[I] * Z[M] + a2; }
            vol = vol + X[T[nn]] * amen[nn] * (a3 - a2) / 6.0;
            b[nn] = amen[nn] * (a3 - a2) / 6.0;
            a2 = a3 = 0.0;
            for (int ii = 0; ii < 3; ++ii) { int I = T[o3[ii]], J = T[o3[(ii + 1) % 3]], M = T[o3[(ii + 2) % 3]]; a3 = X[I] * Z[J] + a3; a2 = X[I] * Z[M] + a2; }
            c[nn] = amen[nn + 1] * (a3 - a2) / 6.0;
            a2 = a3 = 0.0;
            for (int ii = 0; ii < 3; ++ii) { int I = T[o3[ii]], J = T[o3[(ii + 1) % 3]], M = T[o3[(ii + 2) % 3]]; a3 = X[I] * Y[J] + a3; a2 = X[I] * Y[M] + a2; }
            d[nn] = amen[nn] * (a3 - a2) / 6.0;
        }
    };
    for (size_t e = 0; e < nt; ++e) {
        int T[4];
        tet_nodes(e, T);
        tet[e] = make_int4(T[0], T[1], T[2], T[3]);
        int lay = (int)(e / ((size_t)ntri * 3));
        int zn = S->htri[4 * ((e % ((size_t)ntri * 3)) / 3) + 3] - 1;
        int idx = lay * p.nzone + zn;
        for (int q = 0; q < 4; ++q) {
            int nd = T[q];
            pnodi[nd] += p.poros[idx]; snodi[nd] += p.elstor[idx]; vgn[nd] += p.vgn[idx]; vgrmc[nd] += p.vgrmc[idx]; vgpsat[nd] += p.vgpsat[idx];
            kznod[nd] += p.permz[idx];
            tp[nd]++;
            n_cnt[nd + 1]++;
        }
        for (int k = 0; k < 4; ++k)
            for (int l = k; l < 4; ++l) {   // Newton keeps the GEN3D node order (SRC/grdsys.f:63 sorts for Picard only)
                int lo = std::min(T[k], T[l]), hi = std::max(T[k], T[l]);
                int dg = diag_of(hi - lo);
                if (dg < 0) FAIL(-3, "unexpected node pair offset %d in tetrahedron %zu", hi - lo, e);
                s_cnt[(size_t)dg * n + lo + 1]++;
            }
    }
    for (int k = 0; k < n; ++k) {
        if (tp[k] == 0) FAIL(-3, "node %d is not connected to any element", k + 1);
        pnodi[k] /= tp[k]; snodi[k] /= tp[k]; vgn[k] /= tp[k]; vgpsat[k] /= tp[k]; vgrmc[k] /= tp[k]; kznod[k] /= tp[k];
    }
    // ELL widths per diagonal / for the node family, then transposed fill (entry c of row k at [c][k])
    int wd[NDIAG], wnode = 0;
    for (int d = 0; d < NDIAG; ++d) { wd[d] = 0; for (int k = 0; k < n; ++k) wd[d] = std::max(wd[d], s_cnt[(size_t)d * n + k + 1]); }
    for (int k = 0; k < n; ++k) wnode = std::max(wnode, n_cnt[k + 1]);
    const size_t ld = S->ld;
    size_t wtot = wnode;
    for (int d = 0; d < NDIAG; ++d) wtot += wd[d];
    const bool newton = p.iopt == 2;
    std::vector<int> e_tet(wtot * ld, 0);
    std::vector<double> e_coef(wtot * ld, 0.0), e_coef2((size_t)wnode * ld, 0.0), m4(n, 0.0);
    // Newton extras: local node indices of (row, column) inside each listed tet, and per-tet unit-kr stiffness / gravity / volume
    std::vector<unsigned char> e_loc(newton ? (wtot - wnode) * ld : 0, 0);
    std::vector<double> tet_k0(newton ? 10 * nt : 0), tet_gz(newton ? 4 * nt : 0), tet_vol(newton ? nt : 0);
    size_t fam_off[NDIAG + 1];
    fam_off[0] = 0;
    for (int d = 0; d < NDIAG; ++d) fam_off[d + 1] = fam_off[d] + (size_t)wd[d] * ld;   // node family starts at fam_off[NDIAG]
    std::vector<int> s_fill(nslots, 0), n_fill(n, 0);
    for (size_t e = 0; e < nt; ++e) {
        int T[4] = {tet[e].x, tet[e].y, tet[e].z, tet[e].w};
        double b[4], c[4], d[4], vol;
        geometry(T, b, c, d, vol);
        if (vol == 0.0) FAIL(-3, "zero volume at element %zu", e + 1);
        int ivol = vol < 0.0 ? -1 : 1;
        double V = std::fabs(vol), VR = 1.0 / V;
        int lay = (int)(e / ((size_t)ntri * 3));
        int zn = S->htri[4 * ((e % ((size_t)ntri * 3)) / 3) + 3] - 1;
        int idx = lay * p.nzone + zn;
        double kx = p.permx[idx] * VR, ky = p.permy[idx] * VR, kz = p.permz[idx] * VR;
        double pel = (((pnodi[T[0]] + pnodi[T[1]]) + pnodi[T[2]]) + pnodi[T[3]]) * 0.25;   // PICUNS' NODELT(PNODI,PEL)
        for (int q = 0; q < 4; ++q) {
            volnod[T[q]] += V * 0.25;
            size_t pos = fam_off[NDIAG] + (size_t)(n_fill[T[q]]++) * ld + T[q];
            e_tet[pos] = (int)e;
            e_coef[pos] = p.permz[idx] * d[q] * ivol;
            e_coef2[pos - fam_off[NDIAG]] = V * 0.25;
            m4[T[q]] += (V * pel) * 0.25;
        }
        for (int k = 0, pr = 0; k < 4; ++k)
            for (int l = k; l < 4; ++l, ++pr) {
                int lo = std::min(T[k], T[l]), hi = std::max(T[k], T[l]);
                int dg = diag_of(hi - lo);
                size_t pos = fam_off[dg] + (size_t)(s_fill[(size_t)dg * n + lo]++) * ld + lo;
                double kk = (kx * b[k]) * b[l] + (ky * c[k]) * c[l] + (kz * d[k]) * d[l];
                e_tet[pos] = (int)e;
                e_coef[pos] = kk;
                if (newton) {
                    int la = T[k] <= T[l] ? k : l, lb = T[k] <= T[l] ? l : k;   // local index of the row node (lo) and of the column node (hi)
                    e_loc[pos] = (unsigned char)(la | (lb << 2) | 16);   // bit 4: real (non-padding) entry
                    tet_k0[(size_t)pr * nt + e] = kk;
                }
            }
        if (newton) {
            for (int q = 0; q < 4; ++q) tet_gz[(size_t)q * nt + e] = p.permz[idx] * ivol * d[q];
            tet_vol[e] = V;
        }
    }
    {
        const size_t nsz = (size_t)nstr * p.nzone;
        S->h_perm.resize(3 * nsz);
        for (size_t q = 0; q < nsz; ++q) { S->h_perm[q] = p.permx[q]; S->h_perm[nsz + q] = p.permy[q]; S->h_perm[2 * nsz + q] = p.permz[q]; }
    }
    S->hexist.assign(nslots, 0);
    S->nterm = 0;
    for (size_t s = 0; s < nslots; ++s) if (s_cnt[s + 1] > 0) { S->hexist[s] = 1; S->nterm++; }
    // --- derived VG constants (SRC/chparm.f:22-35)
    std::vector<double> vgm(n), vgn1(n), vgnr(n), vgpsn(n), vgmr(n), vgpnot(n), rr(n), vgm52(n), vgmm1(n);
    for (int k = 0; k < n; ++k) {
        vgm[k] = (vgn[k] - 1.0) / vgn[k]; vgn1[k] = vgn[k] - 1.0; vgnr[k] = 1.0 / vgn[k];
        vgpsn[k] = std::pow(std::fabs(vgpsat[k]), vgn[k]); vgmr[k] = 1.0 / vgm[k];
        vgpnot[k] = (pnodi[k] - vgrmc[k]) / pnodi[k]; rr[k] = vgrmc[k] / pnodi[k];
        vgmm1[k] = vgm[k] - 1.0; vgm52[k] = 2.5 * vgm[k];
    }
    if (p.ivghu == 1) {
        // extended van Genuchten (SRC/chparm.f:36-78): vgpnot <- PNOT, the head between the curve's inflexion point and 0 at which
        // d(theta)/d(psi) = SS (interval halving with the reference's stopping rule: half-width < 1e-14 or an exact root); rr <- VGRMC
        for (int k = 0; k < n; ++k) {
            const double m1 = vgm[k] + 1.0, ss = snodi[k], tsr = pnodi[k] - vgrmc[k], target = ss * vgpsn[k] / (vgn1[k] * tsr);
            const double dmcmax = -vgm[k] * vgn[k] * tsr * std::pow(vgm[k], vgm[k]) / (vgpsat[k] * std::pow(m1, m1));
            if (ss >= dmcmax) FAIL(-2, "IVGHU=1: SNODI = %g at node %d must be smaller than DMCMAX = %g (SRC/chparm.f:48-52)", ss, k + 1, dmcmax);
            auto g = [&](double h) { return std::pow(std::fabs(h), vgn1[k]) / std::pow(1.0 + std::pow(h / vgpsat[k], vgn[k]), m1) - target; };
            double lo = vgpsat[k] * std::pow(vgm[k], 1.0 / vgn[k]), hi = 0.0, mid = 0.0;
            bool found = false;
            for (int it = 0; it < 500 && !found; ++it) {
                const double half = (hi - lo) / 2.0;
                mid = lo + half;
                const double gm = g(mid);
                if (gm == 0.0 || half < 1.0e-14) found = true;
                else if (g(lo) * gm > 0.0) lo = mid;
                else hi = mid;
            }
            if (!found) FAIL(-2, "IVGHU=1: the bisection for PNOT did not converge at node %d (SRC/chparm.f:71-73)", k + 1);
            vgpnot[k] = mid; rr[k] = vgrmc[k];
        }
    }
    // --- vegetation type per surface node (SRC/datin.f:236-246)
    std::vector<int> veg(nnod);
    {
        std::vector<double> acc(nnod, 0.0);
        std::vector<int> c2(nnod, 0);
        for (int i = 0; i < nrow; ++i)
            for (int j = 0; j < ncol; ++j) {
                int n00 = i * nc1 + j, n10 = n00 + nc1, n11 = n10 + 1, n01 = n00 + 1;
                double e = p.root_map[i * ncol + j] * p.factor;
                int t1[3] = {n00, n10, n11}, t2[3] = {n00, n11, n01};
                for (int q = 0; q < 3; ++q) { acc[t1[q]] += e; c2[t1[q]]++; acc[t2[q]] += e; c2[t2[q]]++; }
            }
        for (int k = 0; k < nnod; ++k) { int v = (int)(acc[k] / c2[k]); veg[k] = std::min(std::max(v, 1), p.nveg) - 1; }
        if (S->dd) veg = S->ovr_veg;
    }
    std::vector<double> vegpar((size_t)6 * p.nveg);
    for (int v = 0; v < p.nveg; ++v) {
        vegpar[6 * v + 0] = p.pcana[v]; vegpar[6 * v + 1] = p.pcref[v]; vegpar[6 * v + 2] = p.pcwlt[v];
        vegpar[6 * v + 3] = p.zroot[v]; vegpar[6 * v + 4] = p.pz[v]; vegpar[6 * v + 5] = p.omgc[v];
    }
    // --- upload
    int rc = 0;
    rc |= S->vgn.upload(vgn); rc |= S->vgm.upload(vgm); rc |= S->vgpsat.upload(vgpsat); rc |= S->vgpnot.upload(vgpnot);
    rc |= S->rr.upload(rr); rc |= S->snodi.upload(snodi); rc |= S->pnodi.upload(pnodi); rc |= S->vgn1.upload(vgn1);
    rc |= S->vgnr.upload(vgnr); rc |= S->vgpsn.upload(vgpsn); rc |= S->vgmr.upload(vgmr); rc |= S->volnod.upload(volnod);
    if (newton || p.ivghu == 1) { rc |= S->vgm52.upload(vgm52); rc |= S->vgmm1.upload(vgmm1); }   // FXVKR needs VGM52 under Picard too
    rc |= S->arenod.upload(S->harenod); rc |= S->z.upload(S->hz); rc |= S->m4.upload(m4); rc |= S->veg.upload(veg);
    rc |= S->vegpar.upload(vegpar); rc |= S->tet.upload(tet);
    if (S->bc_any) rc |= S->kznod.upload(kznod);
    rc |= S->ell_tet.upload(e_tet); rc |= S->ell_coef.upload(e_coef); rc |= S->ell_coef2.upload(e_coef2);
    if (newton) { rc |= S->ell_loc.upload(e_loc); rc |= S->tet_k0.upload(tet_k0); rc |= S->tet_gz.upload(tet_gz); rc |= S->tet_vol.upload(tet_vol); }
    for (int d = 0; d < NDIAG; ++d) S->fam_off[d] = fam_off[d];
    for (int d = 0; d < NDIAG; ++d) { S->plan.diag[d].tet = S->ell_tet.p + fam_off[d]; S->plan.diag[d].coef = S->ell_coef.p + fam_off[d]; S->plan.diag[d].coef2 = nullptr; S->plan.diag[d].w = wd[d]; S->plan.diag[d].pad = 0; }
    S->plan.node.tet = S->ell_tet.p + fam_off[NDIAG]; S->plan.node.coef = S->ell_coef.p + fam_off[NDIAG]; S->plan.node.coef2 = S->ell_coef2.p; S->plan.node.w = wnode;
    S->plan.node.pad = wnode == wd[0] && std::memcmp(e_tet.data() + fam_off[0], e_tet.data() + fam_off[NDIAG], (size_t)wnode * ld * sizeof(int)) == 0;
    // --- tet indices as base(k) + per-class offset (k_assemble_a): build the 27 tables and verify every stored entry against them
    S->geom = PlanGeom{};
    if (S->plan.node.pad && !getenv("CATHY_PLAN_STORED") && (long long)nt < (1LL << 30)) {
        int wrel = 0;
        for (int d = 0; d < NDIAG; ++d) wrel = std::max(wrel, wd[d]);
        const int UNSET = INT32_MIN;
        std::vector<int> rel((size_t)27 * NDIAG * wrel, UNSET);
        bool ok = true;
        for (int k = 0; k < n && ok; ++k) {
            const int l = k / nnod, sidx = k - l * nnod, i = sidx / nc1, j = sidx - i * nc1;
            const int cls = ((l == 0 ? 0 : l == nstr ? 2 : 1) * 3 + (i == 0 ? 0 : i == nrow ? 2 : 1)) * 3 + (j == 0 ? 0 : j == ncol ? 2 : 1);
            const long long base = 3LL * ntri * l + 6LL * ((long long)i * ncol + j);
            for (int d = 0; d < NDIAG && ok; ++d) {
                const int cnt = s_fill[(size_t)d * n + k];
                for (int c = 0; c < cnt; ++c) {
                    const long long r = (long long)e_tet[fam_off[d] + (size_t)c * ld + k] - base;
                    int &slot = rel[((size_t)cls * NDIAG + d) * wrel + c];
                    if (slot == UNSET) slot = (int)r; else if (slot != r) { ok = false; break; }
                }
            }
        }
        if (ok) {
            for (int &v : rel) if (v == UNSET) v = 0;     // padding positions (coefficient 0): any valid tet, the kernel clamps
            if (S->plan_rel.upload(rel)) FAIL(-101, "device allocation of the tet offset tables failed");
            S->geom = PlanGeom{S->plan_rel.p, wrel, nnod, nc1, ncol, nrow, nstr, 3 * ntri, (int)nt};
        }
    }
    S->plan.ld = (long long)ld;
    if (rc) FAIL(-101, "device allocation/upload of static tables failed: %s", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

// file raster (north row first) -> routing linearisation I_BASIN = (col)*NROW + (row from south)
static std::vector<double> to_route(const CathySim *S, const double *north_first)
{
    std::vector<double> d(S->ncell);
    for (int fr = 0; fr < S->nrow; ++fr)
        for (int c = 0; c < S->ncol; ++c) d[(size_t)c * S->nrow + (S->nrow - 1 - fr)] = north_first[(size_t)fr * S->ncol + c];
    return d;
}

static int build_surface(CathySim *S)
{
    const CathyProblem &p = S->p;
    const int nrow = S->nrow, ncol = S->ncol, nc = S->ncell;
    std::vector<double> w1 = to_route(S, p.dtm_w_1), w2 = to_route(S, p.dtm_w_2), p1 = to_route(S, p.dtm_p_outflow_1), p2 = to_route(S, p.dtm_p_outflow_2);
    std::vector<int> seq(nc, -1), qoi(nc);
    for (int q = 0; q < nc; ++q) {
        int ib = p.qoi[q] - 1;
        if (ib < 0 || ib >= nc || seq[ib] != -1) FAIL(-4, "qoi_a entry %d is out of range or repeated", q + 1);
        seq[ib] = q; qoi[q] = ib;
    }
    // donors of every cell in sequential (QOI) order, direction 1 before direction 2 (SRC/route.f:165-166,247-248)
    std::vector<std::vector<std::pair<int, int>>> don(nc);
    std::vector<int> level(nc, 0);
    for (int q = 0; q < nc; ++q) {
        int ib = qoi[q], J = ib % nrow + 1, I = ib / nrow + 1;
        for (int dir = 0; dir < 2; ++dir) {
            double w = dir ? w2[ib] : w1[ib];
            if (w == 0.0) continue;
            int pout = (int)(dir ? p2[ib] : p1[ib]);
            int iii = (int)std::lround((float)(pout - 5) / 3.0f), jjj = pout - 5 - 3 * iii;
            int icv = I + iii, jcv = J + jjj;
            if (dir == 0 && q == nc - 1) continue;              // the outlet keeps its direction-1 outflow
            if (icv < 1 || icv > ncol || jcv < 1 || jcv > nrow) continue;
            int tgt = (icv - 1) * nrow + jcv - 1;
            if (seq[tgt] <= q) FAIL(-4, "drainage pointer of cell %d goes to a cell that is not later in qoi_a", ib + 1);
            don[tgt].push_back({ib, dir});
            level[tgt] = std::max(level[tgt], level[ib] + 1);   // donors precede receivers in QOI order
        }
    }
    int nlev = 0;
    for (int c = 0; c < nc; ++c) nlev = std::max(nlev, level[c] + 1);
    std::vector<int> lptr(nlev + 1, 0), lcell(nc), dptr(nc + 1, 0), dcell;
    std::vector<unsigned char> ddir;
    for (int c = 0; c < nc; ++c) lptr[level[c] + 1]++;
    for (int l = 0; l < nlev; ++l) lptr[l + 1] += lptr[l];
    {
        std::vector<int> fill(lptr.begin(), lptr.end() - 1);
        for (int q = 0; q < nc; ++q) { int ib = qoi[q]; lcell[fill[level[ib]]++] = ib; }
    }
    for (int c = 0; c < nc; ++c) {
        dptr[c + 1] = dptr[c] + (int)don[c].size();
        for (auto &pr : don[c]) { dcell.push_back(pr.first); ddir.push_back((unsigned char)pr.second); }
    }
    std::vector<int> dcode(dcell.size());
    {
        std::vector<int> slot(nc);
        for (int l = 0; l < nlev; ++l) for (int q = lptr[l]; q < lptr[l + 1]; ++q) slot[lcell[q]] = q - lptr[l];
        if ((long long)nc >= (1LL << 28)) FAIL(-2, "surface routing: more than 2^28 cells");
        for (int c = 0; c < nc; ++c)
            for (int dn = dptr[c]; dn < dptr[c + 1]; ++dn) {
                const int dc = dcell[dn], dr = ddir[dn];
                if (level[dc] < level[c] - 1) dcode[dn] = (dc << 3) | (dr << 2) | 0;
                else if (slot[dc] < ROUTE_BLOCK) dcode[dn] = (slot[dc] << 3) | (dr << 2) | 1;
                else dcode[dn] = (dc << 3) | (dr << 2) | 2;
            }
    }
    S->nlevel = nlev; S->outlet_cell = qoi[nc - 1];
    if (getenv("CATHY_ROUTE_DEBUG")) {
        int big = 0, mx = 0; long long sq = 0;
        for (int l = 0; l < nlev; ++l) { const int c = lptr[l + 1] - lptr[l]; mx = std::max(mx, c); if (c > 1024) ++big; sq += (long long)c * c; }
        fprintf(stderr, "routing: %d cells, %d levels, largest level %d cells, %d levels > 1024 cells, level 0..7:", nc, nlev, mx, big);
        for (int l = 0; l < std::min(nlev, 8); ++l) fprintf(stderr, " %d", lptr[l + 1] - lptr[l]);
        fprintf(stderr, " ... last 4:");
        for (int l = std::max(0, nlev - 4); l < nlev; ++l) fprintf(stderr, " %d", lptr[l + 1] - lptr[l]);
        fprintf(stderr, "\n");
    }
    int rc = 0;
    {   // k_route_wave: one record per cell in level order; donors referenced by level-order position
        std::vector<int> posof(nc);
        for (int q = 0; q < nc; ++q) posof[lcell[q]] = q;
        std::vector<double> epl1 = to_route(S, p.dtm_epl_1), epl2 = to_route(S, p.dtm_epl_2), nrcv = to_route(S, p.dtm_nrc), b1v = to_route(S, p.dtm_b1_sf), y1v = to_route(S, p.dtm_y1_sf);
        std::vector<RouteS> rs(nc);
        std::vector<int> dcx;
        for (int q = 0; q < nc; ++q) {
            const int ib = lcell[q];
            RouteS &R = rs[q];
            memset(&R, 0, sizeof R);
            R.w[0] = w1[ib]; R.w[1] = w2[ib]; R.epl[0] = epl1[ib]; R.epl[1] = epl2[ib]; R.nrc = nrcv[ib]; R.b1 = b1v[ib]; R.y1 = y1v[ib];
            R.ib = ib; R.seq = seq[ib]; R.nd = dptr[ib + 1] - dptr[ib]; R.d0 = (int)dcx.size();
            for (int j = 0; j < R.nd; ++j) {
                const int dn = dptr[ib] + j, code = (posof[dcell[dn]] << 1) | ddir[dn];
                if (j < 4) R.dc[j] = code; else dcx.push_back(code);
            }
        }
        if (dcx.empty()) dcx.push_back(0);
        const char *e = getenv("CATHY_ROUTE_WAVE");
        S->route_wave = !(e && atoi(e) == 0) && (long long)nc < (1LL << 30);
        if (S->route_wave) {
            int rw = 0;
            rw |= S->r_rs.upload(rs); rw |= S->r_dcx.upload(dcx); rw |= S->r_handled.alloc(1);
            rw |= S->r_qo.alloc((size_t)2 * ROUTE_NSMAX * nc); rw |= S->r_qin_ring.alloc((size_t)2 * nc); rw |= S->r_vol_ring.alloc((size_t)2 * nc);
            rw |= S->r_best.alloc(3 * 16);
            if (rw) { cudaGetLastError(); S->route_wave = false; }      // no memory for the histories: k_route alone
            else {
                // cluster of 16 CTAs if this device schedules it (non-portable size), else 8
                S->route_cluster = 8;
                if (const char *ec = getenv("CATHY_ROUTE_CLUSTER")) S->route_cluster = std::max(1, std::min(16, atoi(ec)));
                else if (cudaFuncSetAttribute((const void *)k_route_wave, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
                    cudaLaunchConfig_t cfg = {};
                    cudaLaunchAttribute at[1];
                    cfg.gridDim = dim3(16); cfg.blockDim = dim3(ROUTE_WBLOCK);
                    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                    cfg.attrs = at; cfg.numAttrs = 1;
                    int ncl = 0;
                    if (cudaOccupancyMaxActiveClusters(&ncl, (const void *)k_route_wave, &cfg) == cudaSuccess && ncl >= 1) S->route_cluster = 16;
                    else cudaGetLastError();
                } else cudaGetLastError();
                if (S->route_cluster > 8) cudaFuncSetAttribute((const void *)k_route_wave, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            }
        }
    }
    rc |= S->don_code.upload(dcode);
    rc |= S->lv_ptr.upload(lptr); rc |= S->lv_cell.upload(lcell); rc |= S->seqpos.upload(seq); rc |= S->don_ptr.upload(dptr);
    rc |= S->don_cell.upload(dcell); rc |= S->don_dir.upload(ddir);
    rc |= S->r_w1.upload(w1); rc |= S->r_w2.upload(w2);
    rc |= S->r_sl1.upload(to_route(S, p.dtm_local_slope_1)); rc |= S->r_sl2.upload(to_route(S, p.dtm_local_slope_2));
    rc |= S->r_epl1.upload(to_route(S, p.dtm_epl_1)); rc |= S->r_epl2.upload(to_route(S, p.dtm_epl_2));
    rc |= S->r_ks1.upload(to_route(S, p.dtm_kss1_sf_1)); rc |= S->r_ks2.upload(to_route(S, p.dtm_kss1_sf_2));
    rc |= S->r_ws1.upload(to_route(S, p.dtm_ws1_sf_1)); rc |= S->r_ws2.upload(to_route(S, p.dtm_ws1_sf_2));
    rc |= S->r_b1.upload(to_route(S, p.dtm_b1_sf)); rc |= S->r_y1.upload(to_route(S, p.dtm_y1_sf)); rc |= S->r_nrc.upload(to_route(S, p.dtm_nrc));
    DBuf<double> *bufs[] = {&S->sw_sn, &S->q_in_kk, &S->q_in_kkp1, &S->q_out_kk_1, &S->q_out_kk_2, &S->q_out_kkp1_1, &S->q_out_kkp1_2,
                            &S->volume_kk, &S->volume_kkp1, &S->h_water, &S->q_in_kk_sav, &S->q_out_kk_1_sav, &S->q_out_kk_2_sav,
                            &S->volume_kk_sav, &S->q_in_kk_p, &S->q_out_kk_1_p, &S->q_out_kk_2_p, &S->volume_kk_p};
    for (auto *b : bufs) rc |= b->alloc(nc);
    rc |= S->r_ckf1.alloc(nc); rc |= S->r_ckf2.alloc(nc); rc |= S->r_dhd1.alloc(nc); rc |= S->r_dhd2.alloc(nc);
    S->route_static_done = false;
    rc |= S->d_akmax.alloc(3); rc |= S->d_nsurf.alloc(1);
    if (rc) FAIL(-101, "device allocation of surface routing tables failed");
    return 0;
}
