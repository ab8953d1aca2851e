// host_step.cuh -- host side, time loop: atmospheric and BC record bookkeeping (ATMNXT, BCNXT), system assembly and the linear-solver launches, one nonlinear iteration with its CUDA-graph replay, FLOW3D's decision logic, SURF_FLOWTRA, BKSTEP.
// Part of the single translation unit cathy_b200.cu (included in dependency order; shares its structs and helpers).
#pragma once

// ---- atmospheric stream bookkeeping (host) ------------------------------------------------
static void atm_shift_read(CathySim *S, double time)
{   // label 200 of ATMONE / ATMNXT
    while (!(time <= S->atmtim[2])) {
        S->atmtim[0] = S->atmtim[1]; S->atmtim[1] = S->atmtim[2];
        S->atmrec[0] = S->atmrec[1]; S->atmrec[1] = S->atmrec[2];
        if (S->atm_next >= S->p.natm) { S->htiatm = 1; break; }
        S->atmtim[2] = S->p.atm_time[S->atm_next];
        S->atmrec[2] = S->atm_next++;
    }
}
static void atm_interp_launch(CathySim *S, int slot_a, int slot_b, double time, int set_act)
{
    int up = S->atmtim[slot_b] > S->atmtim[slot_a];
    LAUNCH(S, k_atm_interp, nblk(S->nnod, S->grid_n), RED_BLOCK, S->nnod, S->atmtab.p, S->p.hspatm == 0 ? 1 : 0, S->atmrec[slot_a],
           S->atmrec[slot_b], !up, S->atmtim[slot_a], S->atmtim[slot_b], time, S->p.ieto, S->p.scf, S->arenod.p, S->ifatm.p, set_act,
           S->atmpot.p, S->atmact.p);
}
// ---- non-atmospheric BC record streams (SRC/bcone.f, bcnxt.f, bcbak.f, rdndbc.f, neumann.f) ----------------
static void bc_advance(HostBc &b, double time, int &want)
{   // label 200..300: shift the window while TIME > BCTIM(3); piecewise-constant values
    while (!(time <= b.tim[2])) {
        b.tim[0] = b.tim[1]; b.tim[1] = b.tim[2];
        b.slot[0] = b.slot[1]; b.slot[1] = b.slot[2];
        if (b.next >= b.nrec) { b.hti = 1; break; }
        b.tim[2] = b.time[b.next]; b.slot[2] = b.next; b.next++;
    }
    want = b.tim[2] > b.tim[1] ? b.slot[1] : b.slot[2];
}
static void bc_one(HostBc &b, double time, int &want)
{
    b.hti = 0; b.tim[0] = b.tim[1] = b.tim[2] = 0.0; b.slot[0] = b.slot[1] = b.slot[2] = -1; b.next = 0;
    if (b.nrec > 0) { b.tim[2] = b.time[0]; b.slot[2] = 0; b.next = 1; }
    bc_advance(b, time, want);
}
// make record `want` of both streams the active one on the device (dense flag/value arrays + lists)
static int bc_upload(CathySim *S, int want_dir, int want_neu)
{
    const int n = S->n;
    if (want_dir != S->dir.active || want_neu != S->neu.active) S->graph_drop();      // launch sizes follow the node lists
    if (want_dir != S->dir.active) {
        S->dir.active = want_dir;
        int m = S->dir.anbc();
        S->have_dir = m > 0;
        std::vector<unsigned char> flag(n, 0);
        std::vector<double> val(n, 0.0), lv(std::max(m, 1), 0.0);
        std::vector<int> list(std::max(m, 1), 0);
        for (int q = 0; q < m; ++q) {
            int nd = S->dir.node[S->dir.ptr[want_dir] + q] - 1;
            if (nd < 0 || nd >= n) FAIL(-4, "nansfdirbc node %d out of range", nd + 1);
            flag[nd] = 1; val[nd] = S->dir.val[S->dir.ptr[want_dir] + q]; list[q] = nd;
        }
        CK(cudaMemcpyAsync(S->contp_flag.p, flag.data(), n, cudaMemcpyHostToDevice, S->st));
        CK(cudaMemcpyAsync(S->contp_val.p, val.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, S->st));
        CK(cudaMemcpyAsync(S->contp_list.p, list.data(), list.size() * sizeof(int), cudaMemcpyHostToDevice, S->st));
        CK(cudaStreamSynchronize(S->st));   // the staging vectors go out of scope
    }
    if (want_neu != S->neu.active) {
        S->neu.active = want_neu;
        int m = S->neu.anbc();
        S->have_neu = m > 0;
        std::vector<unsigned char> flag(n, 0);
        std::vector<double> q(n, 0.0), ql(std::max(m, 1), 0.0);
        for (int i = 0; i < m; ++i) {
            int nd = S->neu.node[S->neu.ptr[want_neu] + i] - 1;
            if (nd < 0 || nd >= n) FAIL(-4, "nansfneubc node %d out of range", nd + 1);
            flag[nd] = 1; q[nd] += S->neu.val[S->neu.ptr[want_neu] + i]; ql[i] = S->neu.val[S->neu.ptr[want_neu] + i];
        }
        CK(cudaMemcpyAsync(S->contq_flag.p, flag.data(), n, cudaMemcpyHostToDevice, S->st));
        CK(cudaMemcpyAsync(S->qneu.p, q.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, S->st));
        CK(cudaMemcpyAsync(S->qlist.p, ql.data(), ql.size() * sizeof(double), cudaMemcpyHostToDevice, S->st));
        CK(cudaStreamSynchronize(S->st));
    }
    return 0;
}
// NEUMANN (SRC/neumann.f): acts only when the slot-2 record is a free-drainage one (NODIN2 < 0)
static void neumann_device(CathySim *S, const double *ckrw)
{
    int r = S->neu.slot[1];
    if (r < 0 || S->neu.n2d[r] >= 0 || S->neu.active != r) return;
    LAUNCH(S, k_free_drain_list, nblk(S->nnod, S->grid_n), RED_BLOCK, S->nnod, S->nstr, S->arenod.p, ckrw, S->kznod.p, S->qlist.p, S->qneu.p);
}
static int bc_next_both(CathySim *S, bool back)
{
    if (!S->bc_any) return 0;
    int wd = S->dir.active, wn = S->neu.active;
    if (!back) {
        if (S->dir.hti == 0) bc_advance(S->dir, S->time, wd);
        if (S->neu.hti == 0) bc_advance(S->neu, S->time, wn);
    } else {   // BKSTEP: BCNXT if TIME > BCTIM(2) else BCBAK (slot 1 when the window holds an older record)
        if (S->time > S->dir.tim[1]) { if (S->dir.hti == 0) bc_advance(S->dir, S->time, wd); }
        else if (S->dir.tim[0] < S->dir.tim[1]) wd = S->dir.slot[0];
        if (S->time > S->neu.tim[1]) { if (S->neu.hti == 0) bc_advance(S->neu, S->time, wn); }
        else if (S->neu.tim[0] < S->neu.tim[1]) wn = S->neu.slot[0];
    }
    return bc_upload(S, wd, wn);
}

static void atmnxt(CathySim *S)
{
    if (S->htiatm == 0) {
        atm_shift_read(S, S->time);
        atm_interp_launch(S, 1, 2, S->time, 1);
    }
    if (S->have_dir || S->have_neu)
        LAUNCH(S, k_mark_nonatm, nblk(S->nnod, S->grid_n), RED_BLOCK, S->nnod, S->have_dir ? S->contp_flag.p : nullptr,
               S->have_neu ? S->contq_flag.p : nullptr, S->ifatm.p, (int *)nullptr);
    if (S->sf_n > 0) LAUNCH(S, k_sf_mark_nonatm, nblk(S->sf_n, S->grid_n), RED_BLOCK, S->sf_n, S->sf_node.p, S->nnod, S->ifatm.p, (int *)nullptr);
}
static void atmbak(CathySim *S)
{
    if (S->atmtim[0] >= S->atmtim[1]) return;
    atm_interp_launch(S, 0, 1, S->time, 1);   // ATMBAK always interpolates between slots 1 and 2 of the shifted window
}

static void weight_and_copy(CathySim *S, bool iterate = false)
{   // POLD <- PNEW ; PTNEW = WEIGHT (SRC/weight.f); PTOLD (kept for the chord slopes, KSLOPE != 0) is the previous iterate's PTNEW
    // inside the nonlinear loop (SRC/flow3d.f:248-250) and the new PTNEW at the start of a step / after a back-step
    size_t b = (size_t)S->n * sizeof(double);
    cudaMemcpyAsync(S->pold.p, S->pnew.p, b, cudaMemcpyDeviceToDevice, S->st);
    if (S->ptold.p && iterate) cudaMemcpyAsync(S->ptold.p, S->ptnew.p, b, cudaMemcpyDeviceToDevice, S->st);
    if (S->tetaf == 1.0) cudaMemcpyAsync(S->ptnew.p, S->pnew.p, b, cudaMemcpyDeviceToDevice, S->st);
    else LAUNCH(S, k_weight, nblk(S->n, S->grid_n), RED_BLOCK, S->n, S->tetaf, S->pnew.p, S->ptimep.p, S->ptnew.p);
    if (S->ptold.p && !iterate) cudaMemcpyAsync(S->ptold.p, S->ptnew.p, b, cudaMemcpyDeviceToDevice, S->st);
}

// chvelo + storage sum -> returns STORE1 through h_step later; here just launches
static void chvelo_launch(CathySim *S, const double *psi)
{
    if (S->cm.ivghu == 1)
        LAUNCH(S, k_chvelo_xvg, S->grid_n, RED_BLOCK, S->n, make_soil(S), psi, S->volnod.p, S->sw.p, S->ckrw.p, S->store_part.p, S->dd ? S->own.p : (const unsigned char *)nullptr);
    else if (S->cm.ivghu != 0)
        LAUNCH(S, k_chvelo_alt, S->grid_n, RED_BLOCK, S->n, S->cm, S->pnodi.p, psi, S->volnod.p, S->sw.p, S->ckrw.p, S->store_part.p, S->dd ? S->own.p : (const unsigned char *)nullptr);
    else
    LAUNCH(S, k_chvelo, S->grid_n, RED_BLOCK, S->n, make_soil(S), psi, S->volnod.p, S->sw.p, S->ckrw.p, S->store_part.p, S->dd ? S->own.p : (const unsigned char *)nullptr);
}
static int step_final_sync(CathySim *S, double *extra3 = nullptr)
{
    int nbs = nblk(S->nnod, S->grid_n);
    LAUNCH(S, k_step_partial, nbs, RED_BLOCK, S->nnod, S->nstr, S->p.pmin, S->p.pondh_min, S->ifatm.p, S->atmpot.p, S->atmact.p, S->pnew.p, S->spart.p, S->dd ? S->own.p : (const unsigned char *)nullptr);
    LAUNCH(S, k_step_final, 1, RED_BLOCK, nbs, S->spart.p, S->grid_n, S->store_part.p, S->d_step.p);
    if (S->dd) LAUNCH(S, k_dd_combine_step, 1, 32, S->comm->ctx, S->d_step.p, extra3);
    CK(cudaMemcpyAsync(S->h_step, S->d_step.p, sizeof(StepOut), cudaMemcpyDeviceToHost, S->st));
    CK(cudaStreamSynchronize(S->st));
    return 0;
}

static void dd_exchange(CathySim *S, double *vec)
{
    const long long E = (long long)DD_W * (S->nstr + 1) * S->nc1;
    int blocks = (int)std::max<long long>(1, std::min<long long>((E + RED_BLOCK - 1) / RED_BLOCK, S->sms));
    LAUNCH(S, k_dd_send, blocks, RED_BLOCK, S->comm->ctx, vec);
    LAUNCH(S, k_dd_recv, blocks, RED_BLOCK, S->comm->ctx, vec, S->comm->recv_counter);
}
// ---- one Picard iteration on the device: SRC/picard.f:74-198 + MASBAL + NORMS ------------
static int assemble_system(CathySim *S, double deltat)
{
    const int n = S->n;
    S->scaled = false;
    if (S->sf_n > 0) LAUNCH(S, k_sf_apply, nblk(S->sf_n, S->grid_n), RED_BLOCK, S->sf_n, S->sf_node.p, S->sf_ex.p, S->contp_flag.p, S->contp_val.p);
    Diag A = make_diag(S, S->A.p);
    if (S->cm.ivghu == 1)
        LAUNCH(S, k_curves_xvg, nblk(n, S->grid_n), RED_BLOCK, n, make_soil(S), S->ptnew.p, S->pnew.p, S->ptimep.p, S->timep_dirty, S->sw.p, S->ckrw.p, S->et1.p, S->et2.p,
               S->swnew.p, S->swtimep.p);
    else if (S->cm.ivghu != 0)
        LAUNCH(S, k_curves_alt, nblk(n, S->grid_n), RED_BLOCK, n, S->cm, S->snodi.p, S->pnodi.p, S->ptnew.p, S->pnew.p, S->ptimep.p, S->timep_dirty, S->sw.p, S->ckrw.p,
               S->et1.p, S->et2.p, S->swnew.p, S->swtimep.p);
    else if (S->p.kslope != 0)
        LAUNCH(S, k_curves_chord, nblk(n, S->grid_n), RED_BLOCK, n, make_soil(S), S->p.kslope, S->p.tolksl, S->p.psel, S->p.pser, S->ptnew.p, S->ptold.p, S->pnew.p, S->ptimep.p, S->timep_dirty,
               S->sw.p, S->ckrw.p, S->et1.p, S->et2.p, S->swnew.p, S->swtimep.p);
    else
    LAUNCH(S, k_curves, nblk(n, S->grid_n), RED_BLOCK, n, make_soil(S), S->ptnew.p, S->pnew.p, S->ptimep.p, S->timep_dirty, S->sw.p, S->ckrw.p, S->et1.p, S->et2.p, S->swnew.p, S->swtimep.p);
    S->timep_dirty = 0;
    LAUNCH(S, k_tet_avg, nblk(S->nt, 4 * S->grid_n), RED_BLOCK, S->nt, S->tet.p, S->ckrw.p, S->et1.p, S->krt.p, S->e1t.p);
    if (S->geom.rel) LAUNCH(S, k_assemble_a, nblk(n, 4 * S->grid_n), RED_BLOCK, n, S->plan, S->geom, S->krt.p, S->e1t.p, A, S->grav.p, S->m2.p);
    else LAUNCH(S, k_assemble, nblk(n, 4 * S->grid_n), RED_BLOCK, n, S->plan, S->krt.p, S->e1t.p, A, S->grav.p, S->m2.p);
    LAUNCH(S, k_rhs_lhs, nblk(n, S->grid_n), RED_BLOCK, n, S->nnod, A, S->tetaf, 1.0 / deltat, S->ptnew.p, S->pnew.p, S->ptimep.p, S->swnew.p,
           S->swtimep.p, S->m2.p, S->m4.p, S->et2.p, S->grav.p, S->ifatm.p, S->flagp(),
           S->have_neu ? S->qneu.p : (const double *)nullptr, S->atmact.p, S->atmold.p, S->qtranie.p, S->rhs.p, S->xt5.p, S->diag_true.p, S->diag_bc.p, S->graph_dt());
    if (S->tetaf != 1.0)   // off-diagonals of the LHS are TETAF * stiffness (SRC/cfmatp.f:24-26)
        LAUNCH(S, k_scale, nblk((long long)(NDIAG - 1) * S->ld, 8 * S->grid_n), RED_BLOCK, (long long)(NDIAG - 1) * S->ld, S->tetaf, S->A.p + S->ld);
    return 0;
}
static int solve_system2(CathySim *S)
{
    const int n = S->n;
    Diag A = make_diag(S, S->A.p);
    if (!S->scaled) {
        LAUNCH(S, k_sym_scale, nblk(n, S->grid_n), RED_BLOCK, n, A, S->diag_bc.p, S->dis.p);
        LAUNCH(S, k_sym_scale2, nblk(n, S->grid_n), RED_BLOCK, n, A, S->dis.p);
        S->scaled = true;
    }
    Pcg2Args a;
    a.n = n; a.nnod = S->nnod; a.itmax = S->itmax_dev; a.prefetch = S->pcg_prefetch; a.tol = S->tol_dev;
    a.A = A; a.dis = S->dis.p; a.rhs = S->rhs.p;
    a.y = S->pdiff.p; a.p = S->wbv.p; a.r0 = S->wr.p; a.r1 = S->wz.p; a.w0 = S->wp0.p; a.w1 = S->wp1.p; a.s0 = S->wq0.p; a.s1 = S->wq1.p;
    a.ifatm = S->ifatm.p; a.contp_flag = S->flagp(); a.partial = S->partial.p; a.out = S->d_iter.p;
    a.counter = S->d_counter.p; a.epoch0 = S->barrier_epoch;
    void *args[] = {&a};
    CK(cudaEventRecord(S->evp0, S->st));
    if (S->pcg_shared_gpu) { k_pcg2<1024><<<S->grid_pcg, 1024, 0, S->st>>>(a); CK(cudaGetLastError()); }
    else CK(cudaLaunchCooperativeKernel((void *)k_pcg2<1024>, dim3(S->sms), dim3(1024), args, 0, S->st));
    CK(cudaEventRecord(S->evp1, S->st));
    S->launches += 1;
    return 0;
}
static const void *pcg_res2_fn(const CathySim *S)
{
    static const void *const fn[8] = {(const void *)k_pcg_res2<1024, 0>, (const void *)k_pcg_res2<1024, 1>, (const void *)k_pcg_res2<1024, 2>, (const void *)k_pcg_res2<1024, 3>,
                                      (const void *)k_pcg_res2<1024, 4>, (const void *)k_pcg_res2<1024, 5>, (const void *)k_pcg_res2<1024, 6>, (const void *)k_pcg_res2<1024, 7>};
    static const void *const fc[8] = {(const void *)k_pcg_res2<1024, 0, true>, (const void *)k_pcg_res2<1024, 1, true>, (const void *)k_pcg_res2<1024, 2, true>, (const void *)k_pcg_res2<1024, 3, true>,
                                      (const void *)k_pcg_res2<1024, 4, true>, (const void *)k_pcg_res2<1024, 5, true>, (const void *)k_pcg_res2<1024, 6, true>, (const void *)k_pcg_res2<1024, 7, true>};
    const int nc1 = S->ncol + 1, o2 = nc1, o4 = S->nnod - nc1 - 1, o6 = S->nnod - 1;    // = off[2], off[4], off[6] (set later, by the mesh builder)
    return (S->pcg_cluster > 0 ? fc : fn)[(o2 & 1) | ((o4 & 1) << 1) | ((o6 & 1) << 2)];
}
// streaming PCG on the column-major permutation of the system (see create_impl); the same kernel, other offsets
static int solve_system_cm(CathySim *S)
{
    const int n = S->n, NN = S->nnod, L = S->nstr + 1;
    Diag A = make_diag(S, S->A.p), P;
    for (int d = 0; d < NDIAG; ++d) { P.d[d] = S->cm_A.p + (size_t)d * S->ld; P.off[d] = S->cm_off[d]; }
    // old family d -> permuted family; 4, 5, 6 point to a LOWER permuted index: the (symmetric) entry is stored at its other end
    static const int newd[NDIAG] = {0, 3, 5, 7, 6, 4, 2, 1};
    PermArgs pa;
    int q = 0;
    for (int d = 1; d < NDIAG; ++d) {
        const bool swp = d >= 4 && d <= 6;
        pa.src[q] = A.d[d]; pa.dst[q] = P.d[newd[d]]; pa.shift[q] = swp ? S->cm_off[newd[d]] : 0; ++q;
    }
    pa.src[q] = S->diag_bc.p; pa.dst[q] = S->cm_diag.p; pa.shift[q] = 0; ++q;
    pa.src[q] = S->rhs.p; pa.dst[q] = S->cm_rhs.p; pa.shift[q] = 0; ++q;
    pa.nnod = NN; pa.nl = L; pa.n = n;
    const size_t tile = (size_t)L * 33 * sizeof(double);
    k_permute_cols<<<dim3((NN + 31) / 32, q), 256, tile, S->st>>>(pa);
    CK(cudaGetLastError());
    S->launches++;
    PcgArgs a;
    a.rows_cta = 0; a.xres = 0; a.cm = 1;
    a.n = n; a.nnod = NN; a.itmax = S->itmax_dev; a.tol = S->tol_dev;
    a.A = P; a.diag = S->cm_diag.p; a.rhs = S->cm_rhs.p;
    a.x = S->cm_x.p; a.r = S->cm_r.p; a.z = S->cm_z.p; a.p0 = S->cm_p0.p; a.p1 = S->cm_p1.p; a.bv = S->cm_bv.p;
    a.ifatm = nullptr; a.contp_flag = nullptr; a.partial = S->partial.p; a.out = S->d_iter.p;
    a.counter = S->d_counter.p; a.epoch0 = S->barrier_epoch; a.own = nullptr; a.prefetch = S->pcg_prefetch;
    void *args[] = {&a};
    CK(cudaEventRecord(S->evp0, S->st));
    if (S->pcg_shared_gpu) { k_pcg<1024, true, false><<<S->grid_pcg, 1024, 0, S->st>>>(a); CK(cudaGetLastError()); }
    else CK(cudaLaunchCooperativeKernel((void *)k_pcg<1024, true, false>, dim3(S->sms), dim3(1024), args, 0, S->st));
    CK(cudaEventRecord(S->evp1, S->st));
    S->launches++;
    k_unpermute_cols<<<(NN + 31) / 32, 256, tile, S->st>>>(NN, L, S->cm_x.p, S->pdiff.p);
    CK(cudaGetLastError());
    S->launches++;
    return 0;
}
// k_pcg_tma on the permuted system (pcg_tma.cuh): TMA-staged tiles, also the partitioned solver
static int solve_system_tma(CathySim *S)
{
    const int n = S->n, NN = S->nnod, L = S->nstr + 1;
    Diag A = make_diag(S, S->A.p), P;
    for (int d = 0; d < NDIAG; ++d) { P.d[d] = S->cm_A.p + (size_t)d * S->ld; P.off[d] = S->cm_off[d]; }
    static const int newd[NDIAG] = {0, 3, 5, 7, 6, 4, 2, 1};
    PermArgs pa;
    int q = 0;
    for (int d = 1; d < NDIAG; ++d) {
        const bool swp = d >= 4 && d <= 6;
        pa.src[q] = A.d[d]; pa.dst[q] = P.d[newd[d]]; pa.shift[q] = swp ? S->cm_off[newd[d]] : 0; ++q;
    }
    pa.src[q] = S->diag_bc.p; pa.dst[q] = S->cm_diag.p; pa.shift[q] = 0; ++q;
    pa.src[q] = S->rhs.p; pa.dst[q] = S->cm_rhs.p; pa.shift[q] = 0; ++q;
    pa.nnod = NN; pa.nl = L; pa.n = n;
    const size_t tile = (size_t)L * 33 * sizeof(double);
    k_permute_cols<<<dim3((NN + 31) / 32, q), 256, tile, S->st>>>(pa);
    CK(cudaGetLastError());
    S->launches++;
    TmaArgs a;
    const int rowlen = S->nc1 * L;
    a.n = n; a.lo = S->dd ? S->own_a * rowlen : 0; a.hi = S->dd ? S->own_b * rowlen : n;
    a.itmax = S->itmax_dev; a.tol = S->tol_dev; a.A = P; a.dg = S->cm_diag.p; a.rhs = S->cm_rhs.p;
    a.dinv = S->cm_p1.p; a.x = S->cm_x.p; a.r = S->cm_r.p; a.z = S->cm_z.p; a.p = S->cm_p0.p; a.bv = S->cm_bv.p;
    a.partial = S->partial.p; a.counter = S->d_counter.p; a.epoch0 = S->barrier_epoch; a.out = S->d_iter.p;
    const int g = S->pcg_shared_gpu ? S->grid_pcg : S->sms;
    const int ntile = (a.hi - a.lo + TMA_T - 1) / TMA_T;
    a.tiles_cta = std::max(1, (ntile + g - 1) / g);
    a.nl = L;
    a.dd_on = S->dd ? 1 : 0;
    a.zpeer_n = S->tma_zpeer_n; a.zpeer_s = S->tma_zpeer_s; a.ndst0 = S->tma_ndst0; a.sdst0 = S->tma_sdst0; a.nbr = DD_W * rowlen;
    if (S->dd) a.dd = S->comm->ctx; else memset(&a.dd, 0, sizeof a.dd);
    a.prof = nullptr;
    if (getenv("CATHY_TMA_PROF")) {
        if (!S->bres_prof.p && S->bres_prof.alloc(16)) FAIL(-101, "profile buffer allocation failed");
        a.prof = S->bres_prof.p;
    }
    void *args[] = {&a};
    const void *fn = S->dd ? (const void *)k_pcg_tma<true> : (const void *)k_pcg_tma<false>;
    CK(cudaEventRecord(S->evp0, S->st));
    if (S->pcg_shared_gpu) CK(cudaLaunchKernel(fn, dim3(g), dim3(TMA_BLOCK), args, S->tma_smem, S->st));
    else CK(cudaLaunchCooperativeKernel(fn, dim3(g), dim3(TMA_BLOCK), args, S->tma_smem, S->st));
    CK(cudaEventRecord(S->evp1, S->st));
    S->launches++;
    k_unpermute_cols<<<(NN + 31) / 32, 256, tile, S->st>>>(NN, L, S->cm_x.p, S->pdiff.p);
    CK(cudaGetLastError());
    S->launches++;
    return 0;
}
// small meshes: the whole solve in one thread-block cluster, matrix and vectors in shared memory (pcg_cluster.cuh)
static int solve_system_cl(CathySim *S)
{
    PclArgs a;
    a.n = S->n; a.nnod = S->nnod; a.itmax = S->itmax_dev; a.tol = S->tol_dev; a.R = S->pcl_rows; a.H = S->nnod;
    a.A = make_diag(S, S->A.p); a.diag = S->diag_bc.p; a.rhs = S->rhs.p; a.x = S->pdiff.p; a.z = S->wz.p;
    a.ifatm = S->ifatm.p; a.contp_flag = S->flagp(); a.out = S->d_iter.p; a.epoch0 = S->barrier_epoch;
    void *args[] = {&a};
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    cfg.gridDim = dim3(S->pcl_c); cfg.blockDim = dim3(S->pcl_v2 ? S->pcl_block : 1024); cfg.dynamicSmemBytes = S->pcl_smem; cfg.stream = S->st;
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = S->pcl_c; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    // (inside a captured graph the pair would become event-record nodes, which cudaEventElapsedTime does not accept: the replayed
    // iterations are not timed per solve -- CATHY_GRAPH=0 for a PCG time split)
    if (!S->graph_capturing) CK(cudaEventRecord(S->evp0, S->st));
    CK(cudaLaunchKernelExC(&cfg, S->pcl_v2 ? (const void *)k_pcg_cl2 : (const void *)k_pcg_cl, args));
    if (!S->graph_capturing) CK(cudaEventRecord(S->evp1, S->st));
    S->launches++;
    return 0;
}
static int solve_system(CathySim *S)
{
    if (S->pcl_c > 0) return solve_system_cl(S);
    if (!S->dd && S->pcg_algo == 2) return solve_system2(S);
    if (S->tma_on) return solve_system_tma(S);
    if (S->cm_on) return solve_system_cm(S);
    PcgArgs a;
    a.rows_cta = 0; a.xres = 0; a.cm = 0;
    a.n = S->n; a.nnod = S->nnod; a.itmax = S->itmax_dev; a.tol = S->tol_dev;
    a.A = make_diag(S, S->A.p); a.diag = S->diag_bc.p; a.rhs = S->rhs.p;
    a.x = S->pdiff.p; a.r = S->wr.p; a.z = S->wz.p; a.p0 = S->wp0.p; a.p1 = S->wp1.p; a.bv = S->wbv.p;
    a.ifatm = S->ifatm.p; a.contp_flag = S->flagp(); a.partial = S->partial.p; a.out = S->d_iter.p;
    a.counter = S->d_counter.p; a.epoch0 = S->barrier_epoch;
    if (S->graph_mode) {      // frozen launch arguments: the barrier counter restarts from zero in every replayed solve (the memset is part of the graph)
        CK(cudaMemsetAsync(S->d_counter.p, 0, sizeof(unsigned int), S->st));
        a.epoch0 = 0;
    }
    void *args[] = {&a};
    if (!S->graph_capturing) CK(cudaEventRecord(S->evp0, S->st));
    void *fn = nullptr;
    const bool cu = S->pcg_custom != 0;
    switch (S->pcg_block) {
    case 256: fn = cu ? (void *)k_pcg<256, true, false> : (void *)k_pcg<256, false, false>; break;
    case 512: fn = cu ? (void *)k_pcg<512, true, false> : (void *)k_pcg<512, false, false>; break;
    default: fn = cu ? (void *)k_pcg<1024, true, false> : (void *)k_pcg<1024, false, false>; break;
    }
    if (S->pcg_minb == 1 && S->pcg_block == 512) fn = (void *)k_pcg<512, true, false, 1>;      // 128 registers/thread: all stencil loads in flight
    if (S->pcg_minb == 1 && S->pcg_block == 768) fn = (void *)k_pcg<768, true, false, 1>;
    if (S->pcg_minb == 1 && S->pcg_block == 256) fn = (void *)k_pcg<256, true, false, 1>;
    a.own = nullptr;
    a.prefetch = S->pcg_prefetch;
    if (S->dd) { fn = (void *)k_pcg<1024, true, true>; a.own = S->own.p; a.dd = S->comm->ctx; }
    if (!S->dd && (S->pcg_algo == 3 || S->pcg_algo == 4) && S->res_rows > 0) {
        // CG vectors resident in shared memory (k_pcg_res2 / k_pcg_res): one 1024-thread CTA per SM owns res_rows consecutive rows
        a.rows_cta = S->res_rows; a.xres = S->res_x; a.prefetch = S->res_prefetch;
        const size_t smem = (size_t)(3 + S->res_x) * S->res_rows * sizeof(double);
        const void *fres = S->pcg_algo == 4 ? pcg_res2_fn(S) : (const void *)k_pcg_res<1024>;
        if (S->pcg_cluster > 0) {     // the whole solve in one thread-block cluster
            cudaLaunchConfig_t cfg = {};
            cudaLaunchAttribute at[1];
            cfg.gridDim = dim3(S->pcg_cluster); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = smem; cfg.stream = S->st;
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = S->pcg_cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            CK(cudaLaunchKernelExC(&cfg, fres, args));
        } else
        if (S->pcg_shared_gpu) CK(cudaLaunchKernel(fres, dim3(S->grid_pcg), dim3(1024), args, smem, S->st));
        else CK(cudaLaunchCooperativeKernel(fres, dim3(S->grid_pcg), dim3(1024), args, smem, S->st));
        if (!S->graph_capturing) CK(cudaEventRecord(S->evp1, S->st));
        S->launches++;
        return 0;
    }
    if (S->pcg_shared_gpu) {
        // Several handles share this GPU (partition ranks in tests, concurrent ensemble members): the driver runs cooperative
        // launches one at a time, which would serialise members and deadlock ranks that wait for each other inside the kernel.
        // The custom grid barrier only needs co-residency: the caller keeps (handles in flight) x CATHY_PCG_GRID <= #SMs.
        if (S->dd) k_pcg<1024, true, true><<<S->grid_pcg, 1024, 0, S->st>>>(a);
        else k_pcg<1024, true, false><<<S->grid_pcg, 1024, 0, S->st>>>(a);
        CK(cudaGetLastError());
    } else
    CK(cudaLaunchCooperativeKernel(fn, dim3(S->grid_pcg), dim3(S->pcg_block), args, 0, S->st));
    CK(cudaEventRecord(S->evp1, S->st));
    S->launches++;
    return 0;
}

// ---- one Newton iteration's system on the device: SRC/newton.f:52-123 ------------------------
static int assemble_system_newton(CathySim *S, double deltat)
{
    const int n = S->n;
    Diag A = make_diag(S, S->A.p), Ju = make_diag(S, S->Ju.p), Jl = make_diag(S, S->Jl.p);
    // the previous solve's persisting L2 lines go back to normal, so that the gathers below have the whole cache (the device is idle
    // here: the host has just read the previous iteration's scalars)
    if (S->l2_window && S->l2_reset) cudaCtxResetPersistingL2Cache();
    if (S->sf_n > 0) LAUNCH(S, k_sf_apply, nblk(S->sf_n, S->grid_n), RED_BLOCK, S->sf_n, S->sf_node.p, S->sf_ex.p, S->contp_flag.p, S->contp_val.p);
    if (S->cm.ivghu != 0)
        LAUNCH(S, k_curves_newton_alt, nblk(n, S->grid_n), RED_BLOCK, n, S->cm, make_soil(S), S->ptnew.p, S->sw.p, S->ckrw.p, S->et1.p, S->dckrw.p, S->detai.p);
    else
    LAUNCH(S, k_curves_newton, nblk(n, S->grid_n), RED_BLOCK, n, make_soil(S), S->ptnew.p, S->sw.p, S->ckrw.p, S->et1.p, S->dckrw.p, S->detai.p);
    LAUNCH(S, k_tet_newton, nblk(S->nt, 4 * S->grid_n), RED_BLOCK, S->nt, S->tet.p, S->ckrw.p, S->et1.p, S->ptnew.p, S->pnew.p, S->ptimep.p,
           S->tet_k0.p, S->tet_gz.p, S->tet_vol.p, S->tetaf, 1.0 / deltat, S->krt.p, S->e1t.p, S->ts.p, S->s1.p);
    if (S->geom.rel) LAUNCH(S, k_assemble_newton<true>, nblk(n, 4 * S->grid_n), RED_BLOCK, n, S->nt, S->plan, S->geom, S->ell_loc.p, S->krt.p, S->e1t.p, S->ts.p, S->s1.p,
           S->dckrw.p, S->detai.p, A, Ju, Jl, S->grav.p, S->m2.p);
    else LAUNCH(S, k_assemble_newton<false>, nblk(n, 4 * S->grid_n), RED_BLOCK, n, S->nt, S->plan, S->geom, S->ell_loc.p, S->krt.p, S->e1t.p, S->ts.p, S->s1.p,
           S->dckrw.p, S->detai.p, A, Ju, Jl, S->grav.p, S->m2.p);
    LAUNCH(S, k_rhs_lhs_newton, nblk(n, S->grid_n), RED_BLOCK, n, S->nnod, A, Ju, Jl, S->tetaf, 1.0 / deltat, S->ptnew.p, S->pnew.p, S->ptimep.p,
           S->m2.p, S->grav.p, S->ifatm.p, S->flagp(),
           S->have_neu ? S->qneu.p : (const double *)nullptr, S->atmact.p, S->atmold.p, S->rhs.p, S->xt5.p, S->diag_true.p, S->dinv.p);
    return 0;
}
static const void *bicg_res_fn(const int *off)
{
    static const void *const fn[8] = {(const void *)k_bicgstab_res<1024, 0>, (const void *)k_bicgstab_res<1024, 1>, (const void *)k_bicgstab_res<1024, 2>, (const void *)k_bicgstab_res<1024, 3>,
                                      (const void *)k_bicgstab_res<1024, 4>, (const void *)k_bicgstab_res<1024, 5>, (const void *)k_bicgstab_res<1024, 6>, (const void *)k_bicgstab_res<1024, 7>};
    return fn[(off[2] & 1) | ((off[4] & 1) << 1) | ((off[6] & 1) << 2)];
}
// resident-vector, line-preconditioned BiCGSTAB in the column-major permutation (bicg_res.cuh)
static int solve_system_newton_res(CathySim *S)
{
    const int n = S->n, NN = S->nnod, L = S->nstr + 1;
    Diag Ju = make_diag(S, S->Ju.p), Jl = make_diag(S, S->Jl.p);
    Diag U, Lw;
    for (int d = 0; d < NDIAG; ++d) { U.d[d] = S->bres_u.p + (size_t)d * S->ld; Lw.d[d] = S->bres_l.p + (size_t)d * S->ld; U.off[d] = Lw.off[d] = S->bres_off[d]; }
    // old family d (direction in (layer, row, column)) -> permuted family; families 4, 5, 6 point to a LOWER permuted index, so
    // their upper and lower parts swap roles and are indexed by the other end of the entry (shift = permuted offset)
    static const int newd[NDIAG] = {0, 3, 5, 7, 6, 4, 2, 1};
    PermArgs pa;
    int q = 0;
    for (int d = 0; d < NDIAG; ++d) {
        const bool swp = d >= 4 && d <= 6;
        pa.src[q] = Ju.d[d]; pa.dst[q] = swp ? Lw.d[newd[d]] : U.d[newd[d]]; pa.shift[q] = swp ? S->bres_off[newd[d]] : 0; ++q;
        if (d > 0) { pa.src[q] = Jl.d[d]; pa.dst[q] = swp ? U.d[newd[d]] : Lw.d[newd[d]]; pa.shift[q] = swp ? S->bres_off[newd[d]] : 0; ++q; }
    }
    pa.src[q] = S->rhs.p; pa.dst[q] = S->bres_rhs.p; pa.shift[q] = 0; ++q;
    pa.src[q] = S->dinv.p; pa.dst[q] = S->bres_dinv.p; pa.shift[q] = 0; ++q;
    pa.nnod = NN; pa.nl = L; pa.n = n;
    const size_t tile = (size_t)L * 33 * sizeof(double);
    k_permute_cols<<<dim3((NN + 31) / 32, q), 256, tile, S->st>>>(pa);
    CK(cudaGetLastError());
    S->launches++;
    {   // symmetric groups of 64 rows (CATHY_BRES_SYM=0: never use the upper arrays for the lower triangle)
        const char *e = getenv("CATHY_BRES_SYM");
        const int npass = (S->bres_rows + 2047) / 2048, gg = S->pcg_shared_gpu ? S->grid_pcg : S->sms;
        if (e && atoi(e) == 0) CK(cudaMemsetAsync(S->bres_symf.p, 0, (size_t)gg * npass * 32, S->st));
        else LAUNCH(S, k_bres_sym_flags, S->grid_n, RED_BLOCK, n, S->bres_rows, npass, U, Lw, S->bres_symf.p);
    }
    BresArgs a;
    a.symf = S->bres_symf.p;
    a.n = n; a.itmax = S->itmax_dev; a.tol = S->tol_dev; a.U = U; a.L = Lw;
    a.rhs = S->bres_rhs.p; a.dinv = S->bres_dinv.p; a.x = S->bres_x.p; a.ph = S->bres_ph.p; a.sh = S->bres_sh.p; a.rt = S->bres_rt.p; a.p = S->bres_p.p;
    a.partial = S->partial.p; a.counter = S->d_counter.p; a.epoch0 = S->barrier_epoch; a.out = S->d_iter.p;
    a.rows_cta = S->bres_rows; a.nl = L; a.cols_cta = S->bres_cols;
    {   // CATHY_BICG_ZIGZAG=0/1 overrides; default: on when the Jacobian does not fit the L2
        const char *e = getenv("CATHY_BICG_ZIGZAG");
        a.zigzag = e ? atoi(e) != 0 : (size_t)n * 240 > ((size_t)100 << 20);
    }
    { const char *e = getenv("CATHY_BRES_POINT"); a.point = e ? atoi(e) != 0 : 0; }
    {   // opt-in: measured on B200 at config 3 the products already run at ~80 % of the HBM copy peak (DRAM-bound, ncu) and the extra
        // prefetch instructions cost more than they hide (P1 24.2 -> 27.0 us; a TMA bulk prefetch issued by one thread: 76 -> 84 us/iteration)
        const char *e = getenv("CATHY_BRES_PREFETCH");
        a.prefetch = e ? atoi(e) != 0 : 0;
    }
    a.prof = nullptr;
    if (getenv("CATHY_BRES_PROF")) {
        if (!S->bres_prof.p && S->bres_prof.alloc(16)) FAIL(-101, "profile buffer allocation failed");
        a.prof = S->bres_prof.p;
    }

    void *args[] = {&a};
    const void *fn = bicg_res_fn(S->bres_off);
    const size_t smem = S->bres_smem;
    const int g = S->pcg_shared_gpu ? S->grid_pcg : S->sms;
    // CATHY_L2_PERSIST (opt-in): the permuted Jacobian's lines are marked persisting for this launch (as many as the set-aside holds)
    cudaStreamAttrValue av = {};
    if (S->l2_persist) {
        const size_t jbytes = ((size_t)2 * NDIAG * S->ld + 4 * S->bres_halo) * sizeof(double);
        av.accessPolicyWindow.base_ptr = S->bres_u.base;
        av.accessPolicyWindow.num_bytes = std::min(jbytes, S->l2_maxwin);
        av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)S->l2_persist / (double)av.accessPolicyWindow.num_bytes);
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        CK(cudaStreamSetAttribute(S->st, cudaStreamAttributeAccessPolicyWindow, &av));
    }
    CK(cudaEventRecord(S->evp0, S->st));
    if (S->pcg_shared_gpu) CK(cudaLaunchKernel(fn, dim3(g), dim3(1024), args, smem, S->st));
    else CK(cudaLaunchCooperativeKernel(fn, dim3(g), dim3(1024), args, smem, S->st));
    CK(cudaEventRecord(S->evp1, S->st));
    if (S->l2_persist) {
        av.accessPolicyWindow.num_bytes = 0;
        CK(cudaStreamSetAttribute(S->st, cudaStreamAttributeAccessPolicyWindow, &av));
    }
    S->launches++;
    k_unpermute_cols<<<(NN + 31) / 32, 256, tile, S->st>>>(NN, L, S->bres_x.p, S->pdiff.p);
    CK(cudaGetLastError());
    S->launches++;
    return 0;
}
static int solve_system_newton(CathySim *S)
{
    if (S->bres_rows > 0) return solve_system_newton_res(S);
    BicgArgs a;
    a.n = S->n; a.itmax = S->itmax_dev; a.tol = S->tol_dev;
    a.U = make_diag(S, S->Ju.p); a.L = make_diag(S, S->Jl.p); a.dinv = S->dinv.p; a.rhs = S->rhs.p;
    a.x = S->pdiff.p; a.r = S->wr.p; a.rt = S->wz.p; a.p = S->wp0.p; a.ph = S->wp1.p; a.v = S->wbv.p; a.s = S->ws.p; a.sh = S->wsh.p; a.t = S->wt.p;
    a.partial = S->partial.p; a.out = S->d_iter.p; a.counter = S->d_counter.p; a.epoch0 = S->barrier_epoch;
    a.prefetch = S->pcg_prefetch && (size_t)S->n * 240 > ((size_t)64 << 20);     // the Jacobian (2 x 15 diagonals) does not stay in L2
    {   // CATHY_BICG_ZIGZAG=0/1 overrides; default: on when the Jacobian does not fit the L2
        const char *e = getenv("CATHY_BICG_ZIGZAG");
        a.zigzag = e ? atoi(e) != 0 : (size_t)S->n * 240 > ((size_t)100 << 20);
    }
    a.line = S->bicg_line; a.nnod = S->nnod; a.nl = S->nstr + 1; a.idn = S->widn.p; a.cp = S->wcp.p;
    void *args[] = {&a};
    // The Jacobian (15 diagonals, config 3: 102 MB) is read twice per iteration and would fit the 126 MB L2, but the nine work
    // vectors streaming past it evict it every time.  An access-policy window marks its lines PERSISTING for this launch (as many
    // as the device's set-aside holds: hitRatio = set-aside / window) so that the vectors stream through the rest of the cache.
    cudaStreamAttrValue av = {};
    if (S->l2_window) {
        av.accessPolicyWindow.base_ptr = S->Ju.base;
        av.accessPolicyWindow.num_bytes = S->l2_window;
        av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)S->l2_persist / (double)S->l2_window);
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        CK(cudaStreamSetAttribute(S->st, cudaStreamAttributeAccessPolicyWindow, &av));
    }
    CK(cudaEventRecord(S->evp0, S->st));
    if (S->pcg_shared_gpu) { k_bicgstab<1024><<<S->grid_pcg, 1024, 0, S->st>>>(a); CK(cudaGetLastError()); }
    else CK(cudaLaunchCooperativeKernel((void *)k_bicgstab<1024>, dim3(S->sms), dim3(1024), args, 0, S->st));
    if (S->l2_window) {
        av.accessPolicyWindow.num_bytes = 0;
        CK(cudaStreamSetAttribute(S->st, cudaStreamAttributeAccessPolicyWindow, &av));
    }
    CK(cudaEventRecord(S->evp1, S->st));
    S->launches++;
    return 0;
}
// atmospheric switching (SRC/switch_old.f / SRC/switch.f), evaluated inside CONVER (SRC/conver.f:58-71)
static void launch_switch(CathySim *S)
{
    if (!S->surf) LAUNCH(S, k_switch_old, nblk(S->nnod, S->grid_n), RED_BLOCK, S->nnod, S->p.pmin, S->atmpot.p, S->ifatm.p, S->atmact.p, S->pnew.p);
    else {
        cudaMemsetAsync(S->d_flags.p, 0, sizeof(int), S->st);
        LAUNCH(S, k_switch, nblk(S->nnod, S->grid_n), RED_BLOCK, S->nnod, S->deltat, S->p.pmin, S->p.pondh_min, S->arenod.p, S->pondnod.p,
               S->atmpot.p, S->qtranie.p, S->ifatm.p, S->atmact.p, S->pnew.p, S->ovflnod.p, S->d_flags.p, S->graph_dt());
    }
}
// everything one nonlinear iteration puts on the stream, up to and including the read-backs (no synchronisation, no host decision)
static int enqueue_iteration(CathySim *S)
{
    const int n = S->n;
    int rc = S->newton ? assemble_system_newton(S, S->deltat) : assemble_system(S, S->deltat);
    if (rc) return rc;
    rc = S->newton ? solve_system_newton(S) : solve_system(S);
    if (rc) return rc;
    Diag A = make_diag(S, S->A.p);
    const bool fuse_update = !S->newton && S->p.nlrelx != 2;     // Picard: PNEW += PDIFF happens inside k_norms (not with NLRELX = 2: RELXOM needs the new heads first)
    if (!fuse_update)
    LAUNCH(S, k_update, nblk(n, S->grid_n), RED_BLOCK, n, S->nnod, S->pdiff.p, S->pold.p, S->ifatm.p, S->flagp(),
           S->valp(), S->pnew.p);
    if (S->newton) {
        Diag Ju = make_diag(S, S->Ju.p), Jl = make_diag(S, S->Jl.p);
        LAUNCH(S, k_bkflux_n, nblk(S->nnod, S->grid_n), RED_BLOCK, S->nnod, Ju, Jl, S->pdiff.p, S->xt5.p, S->ifatm.p, S->tetaf, S->atmold.p, S->atmact.p);
        if (S->have_dir) {
            int m = S->dir.anbc();
            LAUNCH(S, k_bkflux_list_n, nblk(m, S->grid_n), RED_BLOCK, m, S->contp_list.p, Ju, Jl, S->pdiff.p, S->xt5.p, S->tetaf, S->qpold.p, S->qpnew.p);
            LAUNCH(S, k_flux_sums, 1, RED_BLOCK, m, S->qpnew.p, S->bcsum.p);
        }
        if (S->cm.ivghu != 0)
            LAUNCH(S, k_sw_pair_alt, nblk(n, S->grid_n), RED_BLOCK, n, S->cm, make_soil(S), S->pnew.p, S->ptimep.p, S->timep_dirty, S->swnew.p, S->swtimep.p);
        else
        LAUNCH(S, k_sw_pair, nblk(n, S->grid_n), RED_BLOCK, n, make_soil(S), S->pnew.p, S->ptimep.p, S->timep_dirty, S->swnew.p, S->swtimep.p);
        S->timep_dirty = 0;
    } else {
    const double *dis = S->scaled ? S->dis.p : (const double *)nullptr;
    LAUNCH(S, k_bkflux, nblk(S->nnod, S->grid_n), RED_BLOCK, n, S->nnod, A, S->diag_true.p, S->pdiff.p, S->xt5.p, S->ifatm.p, S->tetaf,
           S->atmold.p, S->atmact.p, dis);
    if (S->have_dir) {
        int m = S->dir.anbc();
        LAUNCH(S, k_bkflux_list, nblk(m, S->grid_n), RED_BLOCK, n, m, S->contp_list.p, A, S->diag_true.p, S->pdiff.p, S->xt5.p, S->tetaf, S->qpold.p, S->qpnew.p, dis);
        LAUNCH(S, k_flux_sums, 1, RED_BLOCK, m, S->qpnew.p, S->bcsum.p);
    }
    }
    if (S->sf_n > 0) {   // SFQ of the actual seepage nodes and their sum SFFLW (BKPIC / BKNEW, FLUXMB)
        if (S->newton) {
            Diag Ju = make_diag(S, S->Ju.p), Jl = make_diag(S, S->Jl.p);
            LAUNCH(S, k_sf_flux<true>, 1, RED_BLOCK, S->sf_n, n, S->sf_node.p, S->sf_ex.p, Ju, Jl, (const double *)nullptr, (const double *)nullptr, S->pdiff.p,
                   S->xt5.p, S->tetaf, S->sf_qp.p, S->sf_q.p, S->d_sf.p);
        } else
            LAUNCH(S, k_sf_flux<false>, 1, RED_BLOCK, S->sf_n, n, S->sf_node.p, S->sf_ex.p, A, A, S->diag_true.p, S->scaled ? S->dis.p : (const double *)nullptr,
                   S->pdiff.p, S->xt5.p, S->tetaf, S->sf_qp.p, S->sf_q.p, S->d_sf.p);
    }
    if (S->have_neu) LAUNCH(S, k_flux_sums, 1, RED_BLOCK, S->neu.anbc(), S->qlist.p, S->bcsum.p + 2);
    const double *omd = nullptr;
    if (S->p.nlrelx == 2) {   // RELXOM (SRC/flow3d.f:168): OMEGA of this iteration from the unrelaxed head change, kept on the device
        const int nb = nblk(n, S->grid_n);
        LAUNCH(S, k_relxom_partial, nb, RED_BLOCK, n, S->pnew.p, S->pold.p, S->relx_part.p);
        LAUNCH(S, k_relxom_final, 1, 1, nb, S->relx_part.p, S->iter, S->d_iter.p, S->d_omega.p);
        omd = S->d_omega.p;
    }
    LAUNCH(S, k_norms, S->grid_n, RED_BLOCK, n, S->nnod, S->pnew.p, S->pold.p, S->rhs.p, S->ptimep.p, S->swnew.p, S->swtimep.p, S->volnod.p,
           S->snodi.p, S->pnodi.p, S->ifatm.p, S->atmact.p, S->npart.p, S->dd ? S->own.p : (const unsigned char *)nullptr,
           S->p.nlrelx == 1 ? S->p.omega : (S->p.nlrelx == 2 ? 0.5 : 1.0), fuse_update ? S->pdiff.p : (const double *)nullptr,
           S->flagp(), S->valp(), omd);
    if (S->p.nlrelx != 0) LAUNCH(S, k_relax, nblk(n, S->grid_n), RED_BLOCK, n, S->p.omega, S->pold.p, S->pnew.p, omd);
    LAUNCH(S, k_norms_final, 1, RED_BLOCK, S->grid_n, S->npart.p, S->pnew.p, S->pold.p, S->d_iter.p);
    if (S->dd) LAUNCH(S, k_dd_combine_iter, 1, 32, S->comm->ctx, S->d_iter.p, S->gnnod, S->grow0 * S->nc1);
    // atmospheric switching is evaluated every iteration when TOLSWI is large (SRC/conver.f:58-71); when it is
    // conditional the host decides after the read-back below.
    const bool switch_always = S->p.tolswi >= 1.0e29;
    if (switch_always) launch_switch(S);
    // EXTALL after the switch (SRC/conver.f:76-99); the two touch disjoint nodes (potential seepage nodes on the surface are IFATM = -1),
    // so it may also run ahead of a switch that the host decides on after the read-back
    if (S->sf_n > 0) {
        LAUNCH(S, k_sf_extall, 1, RED_BLOCK, S->sf_n, S->sf_node.p, S->sf_ex.p, S->sf_exit.p, S->sf_q.p, S->pnew.p, S->d_sf.p);
        CK(cudaMemcpyAsync(&S->h_rb->sf, S->d_sf.p, sizeof(SfOut), cudaMemcpyDeviceToHost, S->st));
    }
    CK(cudaMemcpyAsync(S->h_iter, S->d_iter.p, sizeof(IterOut), cudaMemcpyDeviceToHost, S->st));
    if (switch_always && S->surf) CK(cudaMemcpyAsync(&S->h_rb->pond, S->d_flags.p, sizeof(int), cudaMemcpyDeviceToHost, S->st));
    if (S->have_dir || S->have_neu) CK(cudaMemcpyAsync(S->h_rb->bc, S->bcsum.p, 4 * sizeof(double), cudaMemcpyDeviceToHost, S->st));
    return 0;
}

// One nonlinear iteration (SRC/picard.f:74-198 / SRC/newton.f:52-123, then MASBAL, NORMS, CONVER's switching): the device work is
// enqueue_iteration; the host reads a few scalars back and takes FLOW3D's decisions.
// Small Picard meshes (the cluster solvers of pcg_cluster.cuh): the ~12 launches and copies of an iteration are captured ONCE into
// a CUDA graph and replayed with one cudaGraphLaunch per iteration -- the arguments are frozen, the step-dependent scalars
// {DELTAT, 1/DELTAT} are read from device memory instead.  On large meshes the whole linear solve already is one persistent launch
// and device time is 99 % of the step, so nothing is captured there (and the grid-barrier kernels take a per-launch epoch argument).
// CATHY_GRAPH=0 switches the replay off.  A new BC record drops the graphs (its node lists change launch sizes).
static int picard_iteration(CathySim *S, CathyIterRecord *rec)
{
    const bool switch_always = S->p.tolswi >= 1.0e29;
    S->h_rb->sf = SfOut{0.0, 0, 0}; S->h_rb->pond = 0; S->h_rb->bc[0] = S->h_rb->bc[1] = S->h_rb->bc[2] = S->h_rb->bc[3] = 0.0;
    bool replayed = false;
    if (S->graph_mode) {
        const int v = S->timep_dirty ? 1 : 0;
        if (!S->gexec[v]) {
            cudaGraph_t g = nullptr;
            const int64_t l0 = S->launches;
            const int dirty0 = S->timep_dirty;
            cudaError_t e_begin = cudaStreamBeginCapture(S->st, cudaStreamCaptureModeThreadLocal), e_end = cudaSuccess, e_inst = cudaSuccess;
            bool ok = e_begin == cudaSuccess;
            int rce = 0;
            if (ok) {
                S->graph_capturing = 1;
                rce = enqueue_iteration(S);
                S->graph_capturing = 0;
                e_end = cudaStreamEndCapture(S->st, &g);
                ok = e_end == cudaSuccess && rce == 0 && g != nullptr && S->launch_err == cudaSuccess;
            }
            if (ok) { e_inst = cudaGraphInstantiate(&S->gexec[v], g, 0); ok = e_inst == cudaSuccess; }
            if (getenv("CATHY_GRAPH_DEBUG")) fprintf(stderr, "cathy graph stages: begin %d, enqueue rc %d (%s), end %d, launch_err %d, instantiate %d\n", (int)e_begin, rce, g_err, (int)e_end, (int)S->launch_err, (int)e_inst);
            if (g) cudaGraphDestroy(g);
            S->glaunches[v] = S->launches - l0;
            S->launches = l0; S->timep_dirty = dirty0;         // nothing ran yet
            if (getenv("CATHY_GRAPH_DEBUG")) fprintf(stderr, "cathy graph capture (variant %d): %s, %lld launches, last error %s\n", v, ok ? "ok" : "FAILED",
                                                     (long long)S->glaunches[v], cudaGetErrorString(cudaPeekAtLastError()));
            if (!ok) {      // not capturable on this driver / configuration: run the plain path from now on
                cudaGetLastError(); S->launch_err = cudaSuccess;
                S->graph_drop(); S->graph_mode = 0;
            }
        }
        if (S->graph_mode) {
            if (S->dt_uploaded != S->deltat) {
                S->h_dt[0] = S->deltat; S->h_dt[1] = 1.0 / S->deltat;
                CK(cudaMemcpyAsync(S->d_dt.p, S->h_dt, 2 * sizeof(double), cudaMemcpyHostToDevice, S->st));
                S->dt_uploaded = S->deltat;
            }
            CK(cudaGraphLaunch(S->gexec[v], S->st));
            S->launches += S->glaunches[v];
            S->timep_dirty = 0; S->scaled = false;
            replayed = true;
        }
    }
    if (!replayed) { int rc = enqueue_iteration(S); if (rc) return rc; }
    CK(cudaStreamSynchronize(S->st));
    const SfOut h_sf = S->h_rb->sf;
    const double *h_bc = S->h_rb->bc;
    int h_pond = 0;
    if (switch_always && S->surf) S->h_iter->ponding = S->h_rb->pond;
    const IterOut &o = *S->h_iter;
    S->barrier_epoch = S->graph_mode ? 0u : (unsigned int)o.pad;
    {   // per-launch device time of the PCG kernel (events sit on the launching stream)
        float pm = 0.f;
        if (!replayed) { if (cudaEventElapsedTime(&pm, S->evp0, S->evp1) == cudaSuccess) S->pcg_ms += pm; else cudaGetLastError(); }
        S->pcg_iters += o.pcg_niter; S->pcg_solves++;
    }
    rec->niter = o.pcg_niter; rec->ikmax = o.ikmax + 1; rec->pl2 = o.pl2; rec->pinf = o.pinf; rec->pnew_ik = o.pnew_ik;
    rec->pold_ik = o.pold_ik; rec->fl2 = o.fl2; rec->finf = o.finf;
    if (!switch_always) {
        bool sw = (S->p.l2norm == 0 && o.pinf <= S->p.tolswi) || (S->p.l2norm != 0 && o.pl2 <= S->p.tolswi);
        if (sw) {
            launch_switch(S);
            if (S->surf) { CK(cudaMemcpyAsync(&h_pond, S->d_flags.p, sizeof(int), cudaMemcpyDeviceToHost, S->st)); CK(cudaStreamSynchronize(S->st)); S->ponding = h_pond; }
        }
    } else if (S->surf) S->ponding = o.ponding;
    if (S->dd) {   // the new heads (after SHLPIC and the atmospheric switch) go to the neighbours' ghost rows
        dd_exchange(S, S->pnew.p);
        int h_err = 0;
        CK(cudaMemcpyAsync(&h_err, S->comm->err, sizeof(int), cudaMemcpyDeviceToHost, S->st));
        CK(cudaStreamSynchronize(S->st));
        if (h_err) FAIL(-6, "row-block partition: a peer did not answer within the time-out (rank %d of %d, wait site %d, sequence %d, PCG iterations %d)",
                        S->dd_rank, S->dd_world, h_err % 10, h_err / 10, S->h_iter->pcg_niter);
    }
    // MASBAL scalars (SRC/masbal.f:69-109); fluxes were summed BEFORE the switch, as in the reference
    S->adin = o.adin; S->adout = o.adout; S->anin = o.anin; S->anout = o.anout; S->dstore = o.dstore;
    double dm = 0.5 * S->deltat;
    double vadin = (S->adin + S->adinp) * dm, vadout = (S->adout + S->adoutp) * dm, vanin = (S->anin + S->aninp) * dm, vanout = (S->anout + S->anoutp) * dm;
    S->ndin = S->have_dir ? h_bc[0] : 0.0; S->ndout = S->have_dir ? h_bc[1] : 0.0;
    S->nnin = S->have_neu ? h_bc[2] : 0.0; S->nnout = S->have_neu ? h_bc[3] : 0.0;
    S->vndin = (S->ndin + S->ndinp) * dm; S->vndout = (S->ndout + S->ndoutp) * dm;
    S->vnnin = (S->nnin + S->nninp) * dm; S->vnnout = (S->nnout + S->nnoutp) * dm;
    S->sfflw = h_sf.sfflw;
    S->vsfflw = (S->sfflw + S->sfflwp) * dm;
    if (S->sf_n > 0) { if (h_sf.ksf > 0) { S->ksfcv++; S->ksfcvt += h_sf.ksf; S->ksfzer = 0; } else S->ksfzer = 1; }
    S->vin = vadin + S->vndin + vanin + S->vnnin + 0.0;
    S->vout = vadout + S->vndout + vanout + S->vnnout + S->vsfflw + 0.0;
    S->erras = S->vin + S->vout - S->dstore;
    S->errel = (S->vin + S->vout) != 0.0 ? 100.0 * S->erras / (S->vin + S->vout) : 0.0;
    S->itlin += o.pcg_niter; S->nitert += o.pcg_niter;
    if (o.pcg_niter >= S->itmax_dev) { S->lsfail = 1; S->klsfai++; } else S->lsfail = 0;
    return 0;
}

// FLOW3D's nonlinear loop and decision logic (SRC/flow3d.f:93-294): 0 converged, 1 back-step, 2 no back-step possible
static int flow3d(CathySim *S, int *status)
{
    const CathyProblem &p = S->p;
    for (;;) {
        CathyIterRecord *r = &S->itrec[std::min(S->iter - 1, CATHY_MAXIT - 1)];
        int rc = picard_iteration(S, r);
        if (rc) return rc;
        if (!(r->pinf == r->pinf) || !(r->pl2 == r->pl2) || !(S->h_iter->pcg_err == S->h_iter->pcg_err)) {
            if (!S->lsfail) { S->lsfail = 1; S->klsfai++; }   // NaN guard: a broken-down linear solve is a solver failure -> back-step
        }
        bool itagen = S->iter < p.ituns;
        bool errgmx = (r->pl2 >= p.ernlmx || r->pinf >= p.ernlmx || r->fl2 >= p.ernlmx || r->finf >= p.ernlmx);
        bool normcv = p.l2norm == 0 ? (r->pinf <= p.toluns) : (r->pl2 <= p.toluns);
        const bool sfwait = S->sf_n > 0 && S->sfchek && !S->ksfzer;   // ISFCVG = 1: the exit points must have settled too (SRC/flow3d.f:237-270)
        if ((!S->lsfail && !errgmx && !normcv && itagen) || (!S->lsfail && !errgmx && itagen && sfwait)) {
            if (S->sf_n > 0) cudaMemcpyAsync(S->sf_exit.p, S->sf_ex.p, (size_t)S->sf_n * sizeof(int), cudaMemcpyDeviceToDevice, S->st);
            weight_and_copy(S, true);
            S->iter++;
            continue;
        }
        S->itrtot += S->iter;
        if (!S->lsfail && !errgmx && normcv && !sfwait) { *status = 0; return 0; }
        *status = S->dtgmin ? 1 : 2;
        return 0;
    }
}

static int surf_flowtra(CathySim *S)
{
    LAUNCH(S, k_div_area, nblk(S->nnod, S->grid_n), RED_BLOCK, S->nnod, S->arenod.p, S->ovflnod.p);
    LAUNCH(S, k_nod_cell, nblk(S->ncell, S->grid_n), RED_BLOCK, S->nrow, S->ncol, S->p.dx, S->p.dy, S->ovflnod.p, S->sw_sn.p);
    RouteArgs a;
    a.ncell = S->ncell; a.nlevel = S->nlevel; a.level_ptr = S->lv_ptr.p; a.level_cell = S->lv_cell.p; a.seq = S->seqpos.p;
    a.don_ptr = S->don_ptr.p; a.don_cell = S->don_cell.p; a.don_dir = S->don_dir.p; a.don_code = S->don_code.p;
    a.w1 = S->r_w1.p; a.w2 = S->r_w2.p; a.sl1 = S->r_sl1.p; a.sl2 = S->r_sl2.p; a.epl1 = S->r_epl1.p; a.epl2 = S->r_epl2.p;
    a.ks1 = S->r_ks1.p; a.ks2 = S->r_ks2.p; a.ws1 = S->r_ws1.p; a.ws2 = S->r_ws2.p; a.b1 = S->r_b1.p; a.y1 = S->r_y1.p; a.nrc = S->r_nrc.p;
    a.sw_sn = S->sw_sn.p; a.q_in_kk = S->q_in_kk.p; a.q_in_kkp1 = S->q_in_kkp1.p; a.q_out_kk_1 = S->q_out_kk_1.p; a.q_out_kk_2 = S->q_out_kk_2.p;
    a.q_out_kkp1_1 = S->q_out_kkp1_1.p; a.q_out_kkp1_2 = S->q_out_kkp1_2.p; a.volume_kk = S->volume_kk.p; a.volume_kkp1 = S->volume_kkp1.p;
    a.h_water = S->h_water.p; a.ak_max = S->d_akmax.p; a.nsurf_out = S->d_nsurf.p; a.deltat = S->deltat; a.cellarea = S->p.dx * S->p.dy;
    a.ckf1 = S->r_ckf1.p; a.ckf2 = S->r_ckf2.p; a.dhd1 = S->r_dhd1.p; a.dhd2 = S->r_dhd2.p;
    if (!S->route_static_done) {
        LAUNCH(S, k_route_static, nblk(S->ncell, S->grid_n), RED_BLOCK, a);
        if (S->route_wave) LAUNCH(S, k_route_fill_static, nblk(S->ncell, S->grid_n), RED_BLOCK, S->ncell, S->r_rs.p, S->r_ckf1.p, S->r_ckf2.p, S->r_dhd1.p, S->r_dhd2.p);
        S->route_static_done = true;
    }
    const bool wave = S->route_wave && (S->route_last_nsurf >= 2 || getenv("CATHY_ROUTE_WAVE_ALWAYS"));
    if (S->route_wave) CK(cudaMemsetAsync(S->r_handled.p, 0, sizeof(int), S->st));
    if (wave) {
        RouteWArgs wa;
        wa.r = a; wa.rs = S->r_rs.p; wa.dcx = S->r_dcx.p; wa.qo = S->r_qo.p; wa.qin_ring = S->r_qin_ring.p; wa.vol_ring = S->r_vol_ring.p;
        wa.handled = S->r_handled.p; wa.nsmax = ROUTE_NSMAX; wa.best = S->r_best.p;
        wa.prof = nullptr;
        if (getenv("CATHY_ROUTE_DEBUG")) {
            if (!S->r_prof.p) S->r_prof.alloc(1024);
            wa.prof = S->r_prof.p;
        }
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute at[1];
        // tasks per wavefront ~ sub-steps x cells per level x 4 lanes: one CTA while they fit it (no cluster barrier, L1 prefetch),
        // else as many CTAs of the cluster as they fill
        const long long lanes = 4LL * std::max(1, S->route_last_nsurf) * ((S->ncell + S->nlevel - 1) / std::max(1, S->nlevel));
        int ccl = (int)std::min<long long>(S->route_cluster, std::max<long long>(1, (lanes + ROUTE_WBLOCK - 1) / ROUTE_WBLOCK));
        while (ccl & (ccl - 1)) ++ccl;               // power of two
        ccl = std::min(ccl, S->route_cluster);
        cfg.gridDim = dim3(ccl); cfg.blockDim = dim3(ROUTE_WBLOCK); cfg.stream = S->st;
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = ccl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, k_route_wave, wa));
        S->launches++;
    }
    if (S->route_lanes4) LAUNCH(S, k_route4, 1, ROUTE4_BLOCK, a, S->route_wave ? S->r_handled.p : (const int *)nullptr);
    else LAUNCH(S, k_route, 1, ROUTE_BLOCK, a, S->route_wave ? S->r_handled.p : (const int *)nullptr);
    LAUNCH(S, k_cell_nod, nblk(S->nnod, S->grid_n), RED_BLOCK, S->nrow, S->ncol, S->h_water.p, S->pondnod.p);
    cudaMemsetAsync(S->d_flags.p, 0, sizeof(int), S->st);
    LAUNCH(S, k_pondupd, nblk(S->nnod, S->grid_n), RED_BLOCK, S->nnod, S->p.pondh_min, 1.0 / S->deltat, S->pondnod.p, S->arenod.p, S->atmpot.p,
           S->ifatm.p, S->atmact.p, S->pnew.p, S->d_flags.p);
    int h[2];
    CK(cudaMemcpyAsync(&h[0], S->d_flags.p, sizeof(int), cudaMemcpyDeviceToHost, S->st));
    CK(cudaMemcpyAsync(&h[1], S->d_nsurf.p, sizeof(int), cudaMemcpyDeviceToHost, S->st));
    CK(cudaStreamSynchronize(S->st));
    S->ponding = h[0];
    S->route_last_nsurf = std::max(1, h[1]);
    return h[1];
}
static void copy_cells(CathySim *S, DBuf<double> &dst, DBuf<double> &src) { cudaMemcpyAsync(dst.p, src.p, (size_t)S->ncell * sizeof(double), cudaMemcpyDeviceToDevice, S->st); }
static void zero_cells(CathySim *S, DBuf<double> &dst) { cudaMemsetAsync(dst.p, 0, (size_t)S->ncell * sizeof(double), S->st); }

// BKSTEP (SRC/bkstep.f:54-166)
static void bkstep(CathySim *S)
{
    const CathyProblem &p = S->p;
    size_t bn = (size_t)S->n * sizeof(double), bs = (size_t)S->nnod * sizeof(double);
    cudaMemcpyAsync(S->pnew.p, S->ptimep.p, bn, cudaMemcpyDeviceToDevice, S->st);
    cudaMemcpyAsync(S->ifatm.p, S->ifatmp.p, (size_t)S->nnod * sizeof(int), cudaMemcpyDeviceToDevice, S->st);
    cudaMemcpyAsync(S->atmact.p, S->atmold.p, bs, cudaMemcpyDeviceToDevice, S->st);
    S->time = S->timep;
    S->deltat = S->deltat * p.dtredm - p.dtreds;
    if (S->deltat <= S->dtmin) { S->deltat = S->dtmin; S->dtgmin = 0; } else S->dtgmin = 1;
    S->time = S->time + S->deltat;
    S->kbackt++; S->kback++; S->iter = 1; S->nitert = 0;
    if (S->have_dir) cudaMemcpyAsync(S->qpnew.p, S->qpold.p, (size_t)S->dir.anbc() * sizeof(double), cudaMemcpyDeviceToDevice, S->st);
    if (S->sf_n > 0) {   // SFEX = SFEXIT = SFEXP, SFQ = SFQP (SRC/bkstep.f:56-66)
        cudaMemcpyAsync(S->sf_ex.p, S->sf_exp.p, (size_t)S->sf_n * sizeof(int), cudaMemcpyDeviceToDevice, S->st);
        cudaMemcpyAsync(S->sf_exit.p, S->sf_exp.p, (size_t)S->sf_n * sizeof(int), cudaMemcpyDeviceToDevice, S->st);
        cudaMemcpyAsync(S->sf_q.p, S->sf_qp.p, (size_t)S->sf_n * sizeof(double), cudaMemcpyDeviceToDevice, S->st);
    }
    bc_next_both(S, true);
    if (S->time > S->atmtim[1]) atmnxt(S); else atmbak(S);
    if (!S->surf) LAUNCH(S, k_switch_old, nblk(S->nnod, S->grid_n), RED_BLOCK, S->nnod, p.pmin, S->atmpot.p, S->ifatm.p, S->atmact.p, S->pnew.p);
    else LAUNCH(S, k_adrstn, nblk(S->nnod, S->grid_n), RED_BLOCK, S->nnod, p.pmin, S->atmpot.p, S->ifatm.p, S->atmact.p, S->pnew.p);
    weight_and_copy(S);
    if (S->surf) {
        S->ponding = S->pondp;
        cudaMemcpyAsync(S->ovflnod.p, S->ovflp.p, bs, cudaMemcpyDeviceToDevice, S->st);
        cudaMemcpyAsync(S->d_akmax.p, S->d_akmax.p + 1, sizeof(double), cudaMemcpyDeviceToDevice, S->st);   // AK_MAX = AK_MAX_P
        copy_cells(S, S->q_in_kk, S->q_in_kk_p); copy_cells(S, S->q_out_kk_1, S->q_out_kk_1_p);
        copy_cells(S, S->q_out_kk_2, S->q_out_kk_2_p); copy_cells(S, S->volume_kk, S->volume_kk_p);
    }
}
