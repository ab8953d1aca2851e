// cathy_b200.cu -- B200 (sm_100a) implementation of the CATHY Richards hot path.
//
// Boundary: include/cathy_b200.h.  Reference routines are cited per kernel as
// SRC/<file>:<line> (SRC = pyCATHY/tests/weil_exemple/my_cathy_prj/src).
//
// Data layout: nodes are numbered layer-major like the reference
// (k = layer*NNOD + s, SRC/gen3d.f:31-45), and the prism-split DEM mesh gives
// every matrix row the same 15-point stencil, so the symmetric system matrix is
// stored as 8 dense "upper" diagonals (offsets 0, 1, NC1, NC1+1, NNOD-NC1-1,
// NNOD-NC1, NNOD-1, NNOD) -- a SELL/DIA layout with implicit column indices:
// no index traffic, every load coalesced.  The element -> nonzero scatter map of
// the reference (TETJA, SRC/tetpic.f) becomes a static, per-slot sorted gather
// list, so assembly is atomic-free and deterministic.
//
// Everything numerical runs on the device; the host code in this file is the
// control flow of the time step (BC stream bookkeeping, back-stepping, step
// size control) and reads back a few dozen scalars per nonlinear iteration.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/cathy_b200.h"

namespace cg = cooperative_groups;

#define RMAX_ 1.7e100
#define NDIAG 8
#define RED_BLOCK 256

static thread_local char g_err[1024] = "";
#define FAIL(code, ...)                          \
    do {                                         \
        snprintf(g_err, sizeof g_err, __VA_ARGS__); \
        return (code);                           \
    } while (0)
#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) FAIL(-100, "CUDA error %s at %s:%d", cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#include "richards_kernels.cuh"
#include "partition_dev.cuh"
#include "pcg_kernels.cuh"
#include "newton_kernels.cuh"
#include "bicg_res.cuh"
#include "pcg_tma.cuh"

#include "flow_kernels.cuh"
#include "routing_kernels.cuh"
#include "step_output_kernels.cuh"
#include "host_setup.cuh"
#include "host_step.cuh"

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

const char *cathy_last_error(void) { return g_err; }
int64_t cathy_sizeof_problem(void) { return (int64_t)sizeof(CathyProblem); }
int64_t cathy_sizeof_report(void) { return (int64_t)sizeof(CathyStepReport); }
int32_t cathy_abi_version(void) { return CATHY_ABI_VERSION; }
int32_t cathy_attempt_log(CathySim *S, int32_t max_attempts, int32_t *nrec, double *deltat, double *time, CathyIterRecord *rec)
{
    if (!S) return -1;
    const int na = (int)S->attempts.size();
    for (int a = 0; a < na && a < max_attempts; ++a) {
        const CathySim::Attempt &t = S->attempts[a];
        if (nrec) nrec[a] = t.n;
        if (deltat) deltat[a] = t.deltat;
        if (time) time[a] = t.time;
        if (rec) memcpy(rec + (size_t)a * CATHY_MAXIT, t.rec, sizeof(CathyIterRecord) * t.n);
    }
    return na;
}

void cathy_destroy(CathySim *S)
{
    if (!S) return;
    cudaSetDevice(S->p.device);
    if (S->st) cudaStreamSynchronize(S->st);
    if (S->r_prof.p) {
        std::vector<unsigned long long> h(1024);
        if (cudaMemcpy(h.data(), S->r_prof.p, 1024 * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess) {
            fprintf(stderr, "k_route_wave, last launch: ns per wavefront:");
            for (int w = 1; w < 1024 && h[w] > h[w - 1]; ++w) if (w < 12 || w % 50 == 0) fprintf(stderr, " [%d] %llu", w, h[w] - h[w - 1]);
            fprintf(stderr, "\n");
        }
        S->r_prof.release();
    }
    if (S->bres_prof.p) {
        unsigned long long h[16];
        if (cudaMemcpy(h, S->bres_prof.p, sizeof h, cudaMemcpyDeviceToHost) == cudaSuccess && h[15] > 0 && S->tma_on)
            fprintf(stderr, "k_pcg_tma phases of CTA 0, us per iteration over %llu iterations: phase A %.2f; reduce %.2f; phase B %.2f; reduce %.2f\n", h[15],
                    1e-3 * h[0] / h[15], 1e-3 * h[1] / h[15], 1e-3 * h[2] / h[15], 1e-3 * h[3] / h[15]);
        else if (h[15] > 0) {
            static const char *nm[11] = {"setup", "P1 product", "reduce1", "s update", "Thomas(s)", "sh out + barrier", "P3 product", "reduce4", "P4 updates", "Thomas(p)", "ph out + reduce"};
            fprintf(stderr, "k_bicgstab_res phases of CTA 0, us per iteration over %llu iterations:", h[15]);
            for (int q = 1; q < 11; ++q) fprintf(stderr, " %s %.2f;", nm[q], 1e-3 * (double)h[q] / (double)h[15]);
            fprintf(stderr, " setup total %.1f us\n", 1e-3 * (double)h[0]);
        }
        S->bres_prof.release();
    }
    // DBuf members are plain pointers: release them explicitly
    DBuf<double> *dd[] = {&S->vgn, &S->vgm, &S->vgpsat, &S->vgpnot, &S->rr, &S->snodi, &S->pnodi, &S->vgn1, &S->vgnr, &S->vgpsn, &S->vgmr,
                          &S->volnod, &S->arenod, &S->z, &S->m4, &S->vegpar, &S->ell_coef, &S->ell_coef2, &S->A, &S->diag_true, &S->diag_bc,
                          &S->grav, &S->m2, &S->krt, &S->e1t, &S->pnew, &S->pold, &S->ptimep, &S->ptnew, &S->pdiff, &S->sw, &S->ckrw, &S->ckrwp,
                          &S->et1, &S->et2, &S->swnew, &S->swtimep, &S->rhs, &S->xt5, &S->qtranie, &S->wr, &S->wz, &S->wp0, &S->wp1, &S->wbv,
                          &S->partial, &S->store_part, &S->atmpot, &S->atmact, &S->atmold, &S->atmtab, &S->pondnod, &S->ovflnod, &S->ovflp,
                          &S->scal3, &S->r_w1, &S->r_w2, &S->r_sl1, &S->r_sl2, &S->r_epl1, &S->r_epl2, &S->r_ks1, &S->r_ks2, &S->r_ws1, &S->r_ws2,
                          &S->r_b1, &S->r_y1, &S->r_nrc, &S->r_ckf1, &S->r_ckf2, &S->r_dhd1, &S->r_dhd2, &S->sw_sn, &S->q_in_kk, &S->q_in_kkp1, &S->q_out_kk_1, &S->q_out_kk_2, &S->q_out_kkp1_1,
                          &S->q_out_kkp1_2, &S->volume_kk, &S->volume_kkp1, &S->h_water, &S->q_in_kk_sav, &S->q_out_kk_1_sav, &S->q_out_kk_2_sav,
                          &S->volume_kk_sav, &S->q_in_kk_p, &S->q_out_kk_1_p, &S->q_out_kk_2_p, &S->volume_kk_p, &S->d_akmax};
    for (auto *b : dd) b->release();
    DBuf<int> *di[] = {&S->veg, &S->ell_tet, &S->ifatm, &S->ifatmp, &S->d_flags, &S->lv_ptr, &S->lv_cell,
                       &S->seqpos, &S->don_ptr, &S->don_cell, &S->don_code, &S->d_nsurf};
    for (auto *b : di) b->release();
    { DBuf<double> *nn[] = {&S->widn, &S->wcp, &S->Ju, &S->Jl, &S->dinv, &S->dckrw, &S->detai, &S->ts, &S->s1, &S->ws, &S->wsh, &S->wt, &S->tet_k0, &S->tet_gz, &S->tet_vol, &S->vgm52, &S->vgmm1};
      for (auto *b : nn) b->release(); S->ell_loc.release(); }
    S->dis.release(); S->wq0.release(); S->wq1.release();
    S->snap.release(); S->snap_i.release(); S->plan_rel.release();
    { DBuf<double> *cc[] = {&S->cm_A, &S->cm_diag, &S->cm_rhs, &S->cm_x, &S->cm_r, &S->cm_p0, &S->cm_p1, &S->cm_bv};
      for (auto *b : cc) b->release();
      if (!S->dd) S->cm_z.release(); }
    S->bres_symf.release();
    { DBuf<double> *bb[] = {&S->bres_u, &S->bres_rhs, &S->bres_dinv, &S->bres_x, &S->bres_ph, &S->bres_sh, &S->bres_rt, &S->bres_p};
      for (auto *b : bb) b->release(); }
    S->contp_flag.release(); S->contq_flag.release(); S->contp_val.release(); S->qneu.release(); S->qlist.release(); S->qpnew.release();
    S->qpold.release(); S->kznod.release(); S->bcsum.release(); S->contp_list.release();
    S->ptold.release(); S->relx_part.release(); S->d_omega.release();
    S->sf_node.release(); S->sf_ex.release(); S->sf_exp.release(); S->sf_exit.release(); S->sf_q.release(); S->sf_qp.release(); S->d_sf.release();
    S->r_rs.release(); S->r_dcx.release(); S->r_handled.release(); S->r_qo.release(); S->r_qin_ring.release(); S->r_vol_ring.release(); S->r_best.release();
    S->d_counter.release(); S->tet.release(); S->don_dir.release(); S->npart.release(); S->spart.release(); S->d_iter.release(); S->d_step.release();
    if (S->comm) {
        for (int r = 0; r < DD_MAXW; ++r) if (S->comm->opened[r]) cudaIpcCloseMemHandle(S->comm->peer_base[r]);
        if (S->comm->base) cudaFree(S->comm->base);
        if (S->comm->seq) cudaFree(S->comm->seq);
        if (S->comm->err) cudaFree(S->comm->err);
        if (S->comm->recv_counter) cudaFree(S->comm->recv_counter);
        delete S->comm;
    }
    S->own.release();
    if (S->h_iter) cudaFreeHost(S->h_iter);
    if (S->h_rb) cudaFreeHost(S->h_rb);
    if (S->h_dt) cudaFreeHost(S->h_dt);
    S->graph_drop(); S->d_dt.release();
    if (S->h_step) cudaFreeHost(S->h_step);
    if (S->ev0) cudaEventDestroy(S->ev0);
    if (S->ev1) cudaEventDestroy(S->ev1);
    if (S->evp0) cudaEventDestroy(S->evp0);
    if (S->evp1) cudaEventDestroy(S->evp1);
    if (S->st_copy) { cudaStreamSynchronize(S->st_copy); cudaStreamDestroy(S->st_copy); cudaEventDestroy(S->ev_snap); cudaEventDestroy(S->ev_drained); }
    if (S->st) cudaStreamDestroy(S->st);
    delete[] S->p.atm_time;
    delete S;
}


// CUDA loads kernels lazily at their first launch, and a load may wait for running kernels to drain -- fatal when a running
// kernel is itself waiting for a peer whose next kernel still has to be loaded (partitioned runs).  Touch every kernel once.
static int preload_kernels()
{
    static bool done = false;
    if (done) return 0;
    cudaFuncAttributes at;
    const void *fns[] = {(const void *)k_curves, (const void *)k_chvelo, (const void *)k_tet_avg, (const void *)k_assemble, (const void *)k_assemble_a, (const void *)k_snapshot, (const void *)k_rhs_lhs,
                         (const void *)k_scale, (const void *)k_spmv, (const void *)k_dd_send, (const void *)k_dd_recv, (const void *)k_dd_combine_iter,
                         (const void *)k_dd_combine_step, (const void *)k_pcg<1024, true, true>, (const void *)k_pcg<1024, true, false>, (const void *)k_pcg2<1024>, (const void *)k_sym_scale, (const void *)k_sym_scale2,
                         (const void *)k_pcg<1024, false, false>, (const void *)k_pcg<512, true, false>, (const void *)k_pcg<512, false, false>,
                         (const void *)k_pcg<256, true, false>, (const void *)k_pcg<256, false, false>, (const void *)k_curves_newton,
                         (const void *)k_sw_pair, (const void *)k_tet_newton, (const void *)k_assemble_newton<true>, (const void *)k_assemble_newton<false>, (const void *)k_rhs_lhs_newton,
                         (const void *)k_bkflux_n, (const void *)k_bkflux_list_n, (const void *)k_bicgstab<1024>, (const void *)k_update,
                         (const void *)k_bkflux, (const void *)k_bkflux_list, (const void *)k_mark_nonatm, (const void *)k_flux_sums,
                         (const void *)k_free_drain_list, (const void *)k_norms, (const void *)k_norms_final, (const void *)k_switch,
                         (const void *)k_switch_old, (const void *)k_adrstn, (const void *)k_pondupd, (const void *)k_atm_interp, (const void *)k_etran,
                         (const void *)k_div_area, (const void *)k_nod_cell, (const void *)k_cell_nod, (const void *)k_route, (const void *)k_route_static, (const void *)k_pond_zero,
                         (const void *)k_step_partial, (const void *)k_step_final, (const void *)k_weight, (const void *)k_atmone, (const void *)k_mbinit,
                         (const void *)k_pack_col, (const void *)k_unpack_col, (const void *)k_vel3d, (const void *)k_vnod3d, (const void *)k_recharge, (const void *)k_wtdepth, (const void *)k_curves_alt, (const void *)k_chvelo_alt, (const void *)k_curves_xvg, (const void *)k_chvelo_xvg,
                         (const void *)k_curves_newton_alt, (const void *)k_sw_pair_alt, (const void *)k_relax, (const void *)k_pcg_res<1024>, (const void *)k_permute_cols, (const void *)k_unpermute_cols, (const void *)k_pcg_tma<true>, (const void *)k_pcg_tma<false>, (const void *)k_route_wave, (const void *)k_route_fill_static, (const void *)k_bres_sym_flags, (const void *)k_curves_chord, (const void *)k_route4};
    for (const void *f : fns) CK(cudaFuncGetAttributes(&at, f));
    done = true;
    return 0;
}

static int create_impl(const CathyProblem *prob, CathySim *S)
{
    S->p = *prob;
    S->p.atm_time = nullptr;   // re-pointed below at an owned copy (cathy_destroy frees it)
    CathyProblem &p = S->p;
    std::vector<double> w_dem, w_root, w_ic, w_atm;   // window copies (row-block partition)
    std::vector<int32_t> w_zone;
    if (prob->dd_world > 1) {
        // ---- row-block partition: cut this rank's window out of the GLOBAL rasters ----
        const int gnrow = prob->nrow, ncol = prob->ncol, nc1 = ncol + 1, W = DD_W;
        if (prob->dd_world > DD_MAXW) FAIL(-2, "dd_world = %d > %d", prob->dd_world, DD_MAXW);
        if (prob->dd_rank < 0 || prob->dd_rank >= prob->dd_world || prob->dd_row0 < 0 || prob->dd_row1 > gnrow + 1 || prob->dd_row1 - prob->dd_row0 < W)
            FAIL(-2, "bad row-block partition: rank %d of %d owns node rows [%d,%d) of %d (each rank needs >= %d rows)", prob->dd_rank, prob->dd_world,
                 prob->dd_row0, prob->dd_row1, gnrow + 1, W);
        if (prob->isimgr != 1) FAIL(-2, "row-block partition: surface routing (ISIMGR=2) is not partitioned yet");
        if (prob->iopt != 1) FAIL(-2, "row-block partition: Picard scheme only");
        if ((prob->ndir_rec > 0 && prob->dir_ptr[prob->ndir_rec] > 0) || (prob->nneu_rec > 0 && prob->neu_ptr[prob->nneu_rec] > 0))
            FAIL(-2, "row-block partition: nansfdirbc / nansfneubc records are not partitioned yet");
        const int lo = std::max(0, prob->dd_row0 - W), hi = std::min(gnrow + 1, prob->dd_row1 + W);   // node rows [lo, hi)
        const int lrow = hi - lo - 1;                                                                    // local cell rows [lo, hi-1)
        S->dd = true; S->dd_world = prob->dd_world; S->dd_rank = prob->dd_rank; S->gnrow = gnrow; S->grow0 = lo;
        S->own_a = prob->dd_row0 - lo; S->own_b = prob->dd_row1 - lo; S->gnnod = (gnrow + 1) * nc1;
        // global surface elevations / vegetation classes exactly as the unpartitioned build forms them (same summation order)
        std::vector<double> gz((size_t)S->gnnod, 0.0), gv((size_t)S->gnnod, 0.0);
        std::vector<int> gc((size_t)S->gnnod, 0);
        for (int i = 0; i < gnrow; ++i)
            for (int j = 0; j < ncol; ++j) {
                int n00 = i * nc1 + j, n10 = n00 + nc1, n11 = n10 + 1, n01 = n00 + 1;
                double e = prob->dem[(size_t)i * ncol + j] * prob->factor, r = prob->root_map[(size_t)i * ncol + j] * prob->factor;
                int t1[3] = {n00, n10, n11}, t2[3] = {n00, n11, n01};
                for (int q = 0; q < 3; ++q) { gz[t1[q]] += e; gv[t1[q]] += r; gc[t1[q]]++; }
                for (int q = 0; q < 3; ++q) { gz[t2[q]] += e; gv[t2[q]] += r; gc[t2[q]]++; }
            }
        double zmin = RMAX_;
        for (int k = 0; k < S->gnnod; ++k) { gz[k] /= gc[k]; zmin = std::min(zmin, gz[k]); }
        S->ovr_zmin = zmin;
        const int lnnod = (lrow + 1) * nc1;
        S->ovr_z.assign(gz.begin() + (size_t)lo * nc1, gz.begin() + (size_t)lo * nc1 + lnnod);
        S->ovr_veg.resize(lnnod);
        for (int k = 0; k < lnnod; ++k) { int v = (int)(gv[(size_t)lo * nc1 + k] / gc[(size_t)lo * nc1 + k]); S->ovr_veg[k] = std::min(std::max(v, 1), prob->nveg) - 1; }
        w_dem.assign(prob->dem + (size_t)lo * ncol, prob->dem + (size_t)(lo + lrow) * ncol);
        w_zone.assign(prob->zone + (size_t)lo * ncol, prob->zone + (size_t)(lo + lrow) * ncol);
        w_root.assign(prob->root_map + (size_t)lo * ncol, prob->root_map + (size_t)(lo + lrow) * ncol);
        p.dem = w_dem.data(); p.zone = w_zone.data(); p.root_map = w_root.data();
        const long long gN = (long long)S->gnnod * (prob->nstr + 1);
        (void)gN;
        if ((prob->indp == 0 || prob->indp == 1) && prob->ic_psi) {
            w_ic.resize((size_t)lnnod * (prob->nstr + 1));
            for (int l = 0; l <= prob->nstr; ++l)
                for (int k = 0; k < lnnod; ++k) w_ic[(size_t)l * lnnod + k] = prob->ic_psi[(size_t)l * S->gnnod + (size_t)lo * nc1 + k];
            p.ic_psi = w_ic.data();
        }
        if (prob->ipond != 0) FAIL(-2, "row-block partition: IPOND != 0 is not partitioned yet");
        if (prob->hspatm == 0 && prob->natm > 0) {
            w_atm.resize((size_t)prob->natm * lnnod);
            for (int r = 0; r < prob->natm; ++r)
                for (int k = 0; k < lnnod; ++k) w_atm[(size_t)r * lnnod + k] = prob->atm_val[(size_t)r * S->gnnod + (size_t)lo * nc1 + k];
            p.atm_val = w_atm.data();
        }
        p.nrow = lrow;
    }
    S->nrow = p.nrow; S->ncol = p.ncol; S->nc1 = p.ncol + 1; S->nstr = p.nstr;
    long long nnod = (long long)(p.nrow + 1) * (p.ncol + 1), n = nnod * (p.nstr + 1), nt = 6LL * p.nrow * p.ncol * p.nstr;
    if (n > 2000000000LL || nt * 10 > 2147000000LL) FAIL(-2, "mesh too large for 32-bit indexing in this build (N=%lld, NT=%lld)", n, nt);
    S->nnod = (int)nnod; S->n = (int)n; S->ntri = 2 * p.nrow * p.ncol; S->nt = (int)nt; S->ncell = p.nrow * p.ncol;
    S->surf = p.isimgr == 2;
    S->newton = p.iopt == 2;
    {   // SRC/chparm.f:79-106
        CurveModel &c = S->cm;
        c.ivghu = p.ivghu; c.hupsia = p.hupsia; c.hubeta = p.hubeta; c.hugama = p.hugama; c.huswr = p.huswr; c.huswr1 = 1.0 - p.huswr;
        c.hualb = std::pow(p.hualfa, (double)(int)p.hubeta); c.hugam1 = p.hugama + 1.0; c.hugb = p.hugama * p.hubeta; c.hun = p.hun; c.hua = p.hua;
        c.hub2a = p.hub - 2.0 * p.hua; c.huab = p.hua - p.hub;
        c.bcpsat = p.bcpsat; c.bcbeta = p.bcbeta; c.bcrmc = p.bcrmc; c.bcb1 = p.bcbeta + 1.0; c.bcbps = p.bcbeta / std::fabs(p.bcpsat); c.bc23b = 2.0 + (3.0 * p.bcbeta);
    }
    {
        double *t = new double[std::max(prob->natm, 1)];
        for (int i = 0; i < prob->natm; ++i) t[i] = prob->atm_time[i];
        p.atm_time = t;
    }
    {
        size_t nc = (size_t)p.nrow * p.ncol;
        S->h_dem.assign(p.dem, p.dem + nc); S->h_zone.assign(p.zone, p.zone + nc);      // p.*: the window when partitioned
        S->h_root.assign(p.root_map, p.root_map + nc); S->h_zratio.assign(prob->zratio, prob->zratio + p.nstr);
        const double *vp[6] = {prob->pcana, prob->pcref, prob->pcwlt, prob->zroot, prob->pz, prob->omgc};
        S->h_veg.resize((size_t)6 * p.nveg);
        for (int q = 0; q < 6; ++q) for (int v = 0; v < p.nveg; ++v) S->h_veg[(size_t)q * p.nveg + v] = vp[q][v];
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) FAIL(-102, "no CUDA device available: the CATHY B200 path has no CPU fallback");
    CK(cudaSetDevice(p.device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, p.device));
    S->sms = prop.multiProcessorCount;
    if (!prop.cooperativeLaunch) FAIL(-102, "device does not support cooperative launches");
    { int rcp = preload_kernels(); if (rcp) return rcp; }
    CK(cudaStreamCreateWithFlags(&S->st, cudaStreamNonBlocking));
    CK(cudaEventCreate(&S->ev0)); CK(cudaEventCreate(&S->ev1)); CK(cudaEventCreate(&S->evp0)); CK(cudaEventCreate(&S->evp1));
    if (const char *e = getenv("CATHY_ROUTE_LANES")) S->route_lanes4 = atoi(e) == 4;
    if (const char *e = getenv("CATHY_PCG_BLOCK")) S->pcg_block = atoi(e);
    if (const char *e = getenv("CATHY_PCG_CUSTOM_BARRIER")) S->pcg_custom = atoi(e);
    if (const char *e = getenv("CATHY_PCG_MINB")) S->pcg_minb = atoi(e);
    if (const char *e = getenv("CATHY_PCG_PREFETCH")) S->pcg_prefetch = atoi(e);
    if (const char *e = getenv("CATHY_PCG_ALGO")) S->pcg_algo = atoi(e);
    if (const char *e = getenv("CATHY_BICG_LINE")) S->bicg_line = atoi(e);
    if (S->pcg_block != 256 && S->pcg_block != 512 && S->pcg_block != 1024 && !(S->pcg_block == 768 && S->pcg_minb == 1)) FAIL(-2, "CATHY_PCG_BLOCK must be 256, 512 or 1024");
    S->grid_pcg = S->sms * (1024 / S->pcg_block);   // one full SM worth of threads per SM, persistent
    if (S->pcg_minb == 1) S->grid_pcg = S->sms;
    if (S->dd) { S->pcg_block = 1024; S->pcg_minb = 0; S->grid_pcg = S->sms; }
    if (const char *e = getenv("CATHY_PCG_GRID")) {   // several handles sharing one GPU
        int g = atoi(e);
        if (g >= 1 && g <= S->grid_pcg) { S->grid_pcg = g; S->pcg_shared_gpu = true; S->pcg_block = 1024; S->pcg_custom = 1; S->pcg_minb = 0; }
    }
    if (S->d_counter.alloc(1)) FAIL(-101, "barrier counter allocation failed");
    S->grid_n = S->sms * 8;                      // grid-stride kernels: a multiple of the SM count
    // Effective stopping rule of the device solvers = (ITMXCG x itmxcg_scale, TOLCG x tolcg_scale), both explicit CathyProblem
    // fields, reported by cathy_solver_limits and written into the header of output/iter by the processor.  Defaults
    // (field <= 0): itmxcg_scale = 20 -- the device preconditioners (diagonal / vertical line) need more, cheaper iterations
    // than the reference's IC(0) / ILU(0), and LSFAIL must keep its meaning "did not reach TOLCG"; tolcg_scale = 1 under
    // Picard and 1e-3 under Newton, where the reference tests the ILU(0)-preconditioned residual, a much tighter bound on the
    // error of the ill-conditioned saturated systems than the true residual the device BiCGSTAB tests (measured on the coupled
    // storm fixture: heads agree with the reference to 4e-9 m after 150 steps, +11 % linear iterations).
    S->itmxcg_scale = p.itmxcg_scale > 0.0 ? p.itmxcg_scale : 20.0;
    S->tolcg_scale = p.tolcg_scale > 0.0 ? p.tolcg_scale : (p.iopt == 2 ? 1.0e-3 : 1.0);
    S->itmax_dev = (int)std::min(2.0e9, std::ceil((double)p.itmxcg * S->itmxcg_scale));
    S->tol_dev = p.tolcg * S->tolcg_scale;
    {   // own copies of the BC record tables (the caller's arrays are not kept)
        auto fill = [](HostBc &b, int nrec, const double *t, const int32_t *ptr, const int32_t *node, const double *val, const int32_t *n2d) {
            b.nrec = nrec;
            if (nrec <= 0) return;
            b.time.assign(t, t + nrec); b.ptr.assign(ptr, ptr + nrec + 1);
            b.node.assign(node, node + ptr[nrec]); b.val.assign(val, val + ptr[nrec]);
            if (n2d) b.n2d.assign(n2d, n2d + nrec); else b.n2d.assign(nrec, 0);
        };
        fill(S->dir, prob->ndir_rec, prob->dir_time, prob->dir_ptr, prob->dir_node, prob->dir_val, nullptr);
        fill(S->neu, prob->nneu_rec, prob->neu_time, prob->neu_ptr, prob->neu_node, prob->neu_val, prob->neu_n2d);
        S->bc_any = (S->dir.nrec > 0 && S->dir.ptr[S->dir.nrec] > 0) || (S->neu.nrec > 0 && S->neu.ptr[S->neu.nrec] > 0);
        for (int r = 0; r < S->neu.nrec; ++r) if (S->neu.n2d[r] < 0) S->free_drain = true;
    }
    S->sf_n = (p.nsf > 0 && p.sf_ptr) ? p.sf_ptr[p.nsf] : 0;
    if (S->sf_n > 0) {   // SFVONE (SRC/sfvone.f:42-66), SFINIT's DUPUIT stop (SRC/sfinit.f:32-36), CONVER's ISFONE stop (SRC/conver.f:77-90)
        if (S->dd) FAIL(-2, "row-block partition: seepage faces are not partitioned yet");
        if (p.dupuit != 0) FAIL(-2, "ONLY DUPUIT = 0 FOR THIS CODE (SRC/sfinit.f:36)");
        if (p.isfone != 0) FAIL(-2, "ISFONE = 1 is disabled in the reference (SRC/conver.f:89)");
        S->bc_any = true;
    }
    S->ld = ((size_t)S->n + 31) / 32 * 32;
    if (!S->dd && S->pcg_algo == 4) {
        // k_pcg_res2 pairs the stencil offsets (o, o+1); the prism-split DEM mesh always yields {1 | NC1, NC1+1 | NNOD-NC1-1, NNOD-NC1 | NNOD-1, NNOD}
        const int nc1 = S->ncol + 1, o[NDIAG] = {0, 1, nc1, nc1 + 1, S->nnod - nc1 - 1, S->nnod - nc1, S->nnod - 1, S->nnod};
        if (!(o[3] == o[2] + 1 && o[5] == o[4] + 1 && o[7] == o[6] + 1 && S->grid_pcg <= 160)) S->pcg_algo = 3;
    }
    // Small meshes: one cluster instead of the whole grid (see k_pcg_res2<.., CL>).  Default: meshes up to 32 Ki rows, with the
    // smallest power-of-two cluster that gives every thread at most one pair of rows (2048 rows per CTA); CATHY_PCG_CLUSTER = 0 / C
    // switches it off / forces C CTAs (1..16).
    if (!S->dd && !S->newton && S->pcg_algo == 4) {
        int c = 0;      // opt-in (measured on config 1: 7.0 -> 5.4 us per iteration with 8 CTAs; k_pcg_cl below does 4x better)
        if (const char *e = getenv("CATHY_PCG_CLUSTER")) { c = atoi(e); if (c < 0 || c > PCG_CL_MAX || (c & (c - 1))) FAIL(-2, "CATHY_PCG_CLUSTER must be 0, 1, 2, 4, 8 or 16"); }
        if (c > 0) { S->pcg_cluster = c; S->grid_pcg = c; S->pcg_block = 1024; S->pcg_custom = 1; S->pcg_minb = 0; }
    }
    // k_pcg_cl (pcg_cluster.cuh): default for Picard meshes of up to 16 Ki rows whose cluster-resident working set fits -- one row per
    // thread, the smallest power-of-two cluster with <= 1024 rows per CTA.  CATHY_PCG_CL=0 switches it off; an explicit
    // CATHY_PCG_ALGO or CATHY_PCG_CLUSTER keeps the kernel it names.
    if (!S->dd && !S->newton && S->pcg_cluster == 0 && !getenv("CATHY_PCG_ALGO") && !(getenv("CATHY_PCG_CL") && atoi(getenv("CATHY_PCG_CL")) == 0)) {
        int c = 1;
        while (c < PCG_CL_MAX && (long long)c * 1024 < S->n) c *= 2;
        // more CTAs = fewer rows per warp on the serial path of an iteration (config 1: 4.07 us per iteration with 8 CTAs, 3.46 with 16),
        // as long as a CTA's rows still cover the stencil reach (k_pcg_cl2's window condition NNOD <= rows)
        while (c < PCG_CL_MAX && (S->n + 2 * c - 1) / (2 * c) >= S->nnod) c *= 2;
        if (const char *e = getenv("CATHY_PCG_CL")) { int v = atoi(e); if (v >= 1 && v <= PCG_CL_MAX && !(v & (v - 1)) && (long long)v * 1024 >= S->n) c = v; }
        const int rows = (int)((((size_t)S->n + c - 1) / c + 31) / 32 * 32);
        const size_t smem = ((size_t)NDIAG * (rows + S->nnod) + (size_t)5 * rows) * sizeof(double);
        // k_pcg_cl2 (one barrier per iteration, nothing but shared memory inside the iteration): window vectors for R + 2 H rows
        const size_t smem2 = ((size_t)NDIAG * (rows + S->nnod) + (size_t)6 * (rows + 2 * (size_t)S->nnod) + (size_t)3 * rows) * sizeof(double);
        cudaFuncAttributes at, at2;
        CK(cudaFuncGetAttributes(&at, (const void *)k_pcg_cl));
        CK(cudaFuncGetAttributes(&at2, (const void *)k_pcg_cl2));
        int optin = 0;
        CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, p.device));
        const bool v2 = !(getenv("CATHY_PCG_CL2") && atoi(getenv("CATHY_PCG_CL2")) == 0) && S->nnod <= rows && smem2 + at2.sharedSizeBytes <= (size_t)optin;
        if ((long long)c * 1024 >= S->n && rows <= 1024 && (v2 || smem + at.sharedSizeBytes <= (size_t)optin)) {
            const void *fk = v2 ? (const void *)k_pcg_cl2 : (const void *)k_pcg_cl;
            S->pcl_c = c; S->pcl_rows = rows; S->pcl_smem = v2 ? smem2 : smem; S->pcl_v2 = v2;
            if (const char *e = getenv("CATHY_PCG_CL_BLOCK")) { int b = atoi(e); if (b >= 32 && b <= 1024 && b % 32 == 0) S->pcl_block = b; }
            CK(cudaFuncSetAttribute(fk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)optin - (v2 ? at2 : at).sharedSizeBytes)));
            if (c > 8) CK(cudaFuncSetAttribute(fk, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        }
    }
    if (!S->dd && (S->pcg_algo == 3 || S->pcg_algo == 4)) {
        // k_pcg_res*: r, p, B (and x if there is room) of a CTA's rows stay in its shared memory for the whole solve
        const void *fres = S->pcg_algo == 4 ? pcg_res2_fn(S) : (const void *)k_pcg_res<1024>;
        cudaFuncAttributes at;
        CK(cudaFuncGetAttributes(&at, fres));
        int optin = 0;
        CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, p.device));
        const size_t avail = (size_t)optin > at.sharedSizeBytes ? (size_t)optin - at.sharedSizeBytes : 0;
        const int rows = (int)((((size_t)S->n + S->grid_pcg - 1) / S->grid_pcg + 31) / 32 * 32);
        if ((size_t)3 * rows * sizeof(double) <= avail && rows <= 32 * 1024) {
            S->res_rows = rows;
            S->res_x = (size_t)4 * rows * sizeof(double) <= avail ? 1 : 0;
            if (const char *e = getenv("CATHY_PCG_RES_X")) S->res_x = S->res_x && atoi(e);
            // the diagonals (64 B/row) stay in the 126 MB L2 up to ~1 M rows (measured: prefetch costs 4 % at 848 k rows); beyond that they
            // stream from HBM and the prefetch pays (+13 % at 1.32 M rows)
            S->res_prefetch = S->pcg_prefetch && (size_t)S->n * 64 > ((size_t)64 << 20);
            if (const char *e = getenv("CATHY_PCG_RES_PREFETCH")) S->res_prefetch = atoi(e);
            CK(cudaFuncSetAttribute(fres, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)avail));
            if (S->pcg_cluster > 8) CK(cudaFuncSetAttribute(fres, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        }
        if (S->pcg_cluster > 0 && S->res_rows == 0) FAIL(-2, "CATHY_PCG_CLUSTER=%d: %d rows per CTA do not fit in shared memory", S->pcg_cluster, rows);
    }
    S->halo = ((size_t)S->nnod + 1 + 31) / 32 * 32;
    int rc = build_static(S);
    if (rc) return rc;
    const int N = S->n, NN = S->nnod;
    int a = 0;
    a |= S->A.alloc((size_t)NDIAG * S->ld, S->halo);
    DBuf<double> *vn[] = {&S->diag_true, &S->diag_bc, &S->grav, &S->m2, &S->pnew, &S->pold, &S->ptimep, &S->ptnew, &S->pdiff, &S->sw, &S->ckrw,
                          &S->ckrwp, &S->et1, &S->et2, &S->swnew, &S->swtimep, &S->rhs, &S->xt5, &S->qtranie, &S->wr, &S->wz, &S->wp0, &S->wp1, &S->wbv};
    for (auto *b : vn) a |= b->alloc(N, S->halo);   // halo: stencil gathers need no bounds checks
    if (p.kslope != 0) a |= S->ptold.alloc(N);
    if (p.nlrelx == 2) {
        a |= S->relx_part.alloc(S->grid_n); a |= S->d_omega.alloc(2);
        const double one2[2] = {1.0, 1.0};
        if (!a) cudaMemcpy(S->d_omega.p, one2, sizeof one2, cudaMemcpyHostToDevice);
    }
    a |= S->dis.alloc(N, S->halo); a |= S->wq0.alloc(N, S->halo); a |= S->wq1.alloc(N, S->halo);
    a |= S->krt.alloc(S->nt); a |= S->e1t.alloc(S->nt);
    a |= S->partial.alloc(10 * (size_t)std::max(S->grid_pcg, 1));
    {   // Streaming PCG in the column-major permutation k' = s L + l.  In the layer-major numbering the stencil reaches NNOD rows up and
        // down, and once NNOD rows of matrix + vectors (~100 B each) exceed the L2 (config 5: 1 M rows = 100 MB) the z / p gathers and
        // the lower-triangle reads of the neighbouring layers miss and come from HBM a second and third time.  Column-major, the
        // half bandwidth is (NC1 + 1) L rows (config 5: 31 k rows = 3 MB): every operand is read from HBM once per phase.
        // Default: on whenever the resident kernels do not apply and the mesh is large, and for every partitioned handle;
        // CATHY_PCG_CM=0/1 overrides.  CATHY_PCG_TMA=0 keeps k_pcg (direct loads) on the permuted arrays instead of k_pcg_tma.
        const char *e = getenv("CATHY_PCG_CM"), *et = getenv("CATHY_PCG_TMA");
        const bool want = e ? atoi(e) != 0 : (S->dd || (size_t)N * 100 > ((size_t)48 << 20));
        const bool streaming = S->dd || (!((S->pcg_algo == 3 || S->pcg_algo == 4) && S->res_rows > 0) && S->pcg_algo != 2);
        const bool tma = !(et && atoi(et) == 0);
        if (!S->newton && streaming && want && S->nstr >= 2 && (!S->dd || tma)) {
            const int L = S->nstr + 1, nc1 = S->ncol + 1;
            const int off[NDIAG] = {0, 1, L - 1, L, nc1 * L - 1, nc1 * L, (nc1 + 1) * L - 1, (nc1 + 1) * L};
            for (int d = 0; d < NDIAG; ++d) S->cm_off[d] = off[d];
            S->cm_halo = ((size_t)off[NDIAG - 1] + TMA_T + 34 + 31) / 32 * 32;
            a |= S->cm_A.alloc((size_t)NDIAG * S->ld, S->cm_halo);
            DBuf<double> *cv[] = {&S->cm_diag, &S->cm_rhs, &S->cm_x, &S->cm_r, &S->cm_p0, &S->cm_p1, &S->cm_bv};
            for (auto *b : cv) a |= b->alloc(N, S->cm_halo);
            if (!S->dd) a |= S->cm_z.alloc(N, S->cm_halo);      // partitioned: z lives in the peer-mapped communication block (below)
            S->cm_on = !a;
            if (S->cm_on && tma) {
                const TmaLayout ly = tma_layout(off, L);
                S->tma_smem = (size_t)TMA_NS * ly.total * sizeof(double);
                int optin = 0;
                CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, p.device));
                cudaFuncAttributes at;
                CK(cudaFuncGetAttributes(&at, (const void *)k_pcg_tma<false>));
                if (S->tma_smem + at.sharedSizeBytes <= (size_t)optin) {
                    CK(cudaFuncSetAttribute((const void *)k_pcg_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S->tma_smem));
                    CK(cudaFuncSetAttribute((const void *)k_pcg_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S->tma_smem));
                    CK(cudaFuncSetAttribute((const void *)k_spmv_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S->tma_smem));
                    S->tma_on = true;
                }
            }
            if (S->dd && !S->tma_on) FAIL(-2, "row-block partition: the node layers do not fit the shared-memory tiles of k_pcg_tma (NSTR = %d); set CATHY_PCG_CM=0 for the layer-major kernel", S->nstr);
        }
    }
    if (S->newton) {
        // Ju and Jl share ONE allocation ([halo | 8 upper diagonals | halo][halo | 8 lower | halo]) so that one L2 access-policy
        // window covers the whole Jacobian (see solve_system_newton); Jl is a non-owning view
        a |= S->Ju.alloc((size_t)2 * NDIAG * S->ld + 2 * S->halo, S->halo);
        S->Jl.release();
        if (!a) { S->Jl.p = S->Ju.p + (size_t)NDIAG * S->ld + 2 * S->halo; S->Jl.n = (size_t)NDIAG * S->ld; S->Jl.pad = S->halo; }
        {   // L2 persistence for the Jacobian during the BiCGSTAB solve: OPT-IN (CATHY_L2_PERSIST=1: the device's maximum set-aside,
            // >1: that many MB).  Measured on B200 at config 3 (profiles/micro/r1i_l2_persist.log): the time per BiCGSTAB iteration does
            // not move (91.7 -> 91.6 us with the maximum, 89.4 us with 64 MB), i.e. the solve is not bound by re-reading the Jacobian
            // from HBM, while the set-aside slows every other kernel of the step (Newton workload 10.08 -> 11.85 ms/step).
            S->l2_window = S->l2_persist = 0;
            { const char *r = getenv("CATHY_L2_RESET"); S->l2_reset = !(r && atoi(r) == 0); }
            const char *e = getenv("CATHY_L2_PERSIST");
            const size_t jbytes = ((size_t)2 * NDIAG * S->ld + 4 * S->halo) * sizeof(double);
            int dev = 0, maxp = 0, maxw = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, dev);
            cudaDeviceGetAttribute(&maxw, cudaDevAttrMaxAccessPolicyWindowSize, dev);
            if (e && atoi(e) >= 1 && maxp > 0 && maxw > 0) {
                const size_t want = e && atoi(e) > 1 ? (size_t)atoi(e) << 20 : (size_t)maxp;
                const size_t persist = std::min(std::min((size_t)maxp, want), jbytes);
                if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist) == cudaSuccess) {
                    S->l2_window = std::min(jbytes, (size_t)maxw); S->l2_persist = persist; S->l2_maxwin = (size_t)maxw;
                } else cudaGetLastError();
            }
        }
        DBuf<double> *vv[] = {&S->dinv, &S->dckrw, &S->detai, &S->ws, &S->wsh, &S->wt, &S->widn, &S->wcp};
        for (auto *b : vv) a |= b->alloc(N, S->halo);
        a |= S->ts.alloc(4 * (size_t)S->nt); a |= S->s1.alloc(4 * (size_t)S->nt);
        {   // resident-vector BiCGSTAB in the column-major permutation (bicg_res.cuh); CATHY_BICG_ALGO=0 keeps k_bicgstab
            const char *e = getenv("CATHY_BICG_ALGO");
            const int g = S->pcg_shared_gpu ? S->grid_pcg : S->sms, L = S->nstr + 1, nc1 = S->ncol + 1;
            S->bres_rows = 0;
            if (!(e && atoi(e) == 0) && !S->dd && S->nstr >= 2 && g <= 160) {
                int cols = (NN + g - 1) / g;
                if (((long long)cols * L) & 1) ++cols;
                const long long rows = (long long)cols * L;
                const int off[NDIAG] = {0, 1, L - 1, L, nc1 * L - 1, nc1 * L, (nc1 + 1) * L - 1, (nc1 + 1) * L};
                const void *fn = bicg_res_fn(off);
                cudaFuncAttributes at;
                CK(cudaFuncGetAttributes(&at, fn));
                int optin = 0;
                CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, p.device));
                const size_t avail = (size_t)optin > at.sharedSizeBytes ? (size_t)optin - at.sharedSizeBytes : 0;
                // shared memory: 3 resident vectors (fp64) + 3 line-factor arrays (fp32, column stride L | 1)
                const size_t smem = (size_t)3 * rows * sizeof(double) + (size_t)3 * cols * (L | 1) * sizeof(float);
                if (smem <= avail && rows <= 32 * 1024 && rows * g >= N) {
                    S->bres_rows = (int)rows; S->bres_cols = cols; S->bres_smem = smem;
                    for (int d = 0; d < NDIAG; ++d) S->bres_off[d] = off[d];
                    S->bres_halo = ((size_t)off[NDIAG - 1] + 2 + 31) / 32 * 32;
                    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)avail));
                    // U' and L' share ONE allocation ([halo | 8 upper | halo][halo | 8 lower | halo]) so that one L2 access-policy window covers them
                    a |= S->bres_u.alloc((size_t)2 * NDIAG * S->ld + 2 * S->bres_halo, S->bres_halo);
                    S->bres_l.release();
                    if (!a) { S->bres_l.p = S->bres_u.p + (size_t)NDIAG * S->ld + 2 * S->bres_halo; S->bres_l.n = (size_t)NDIAG * S->ld; S->bres_l.pad = S->bres_halo; }
                    DBuf<double> *bv[] = {&S->bres_rhs, &S->bres_dinv, &S->bres_x, &S->bres_ph, &S->bres_sh, &S->bres_rt, &S->bres_p};
                    for (auto *b : bv) a |= b->alloc(N, S->bres_halo);
                    a |= S->bres_symf.alloc((size_t)g * ((rows + 2047) / 2048) * 32);
                }
            }
        }
    } a |= S->store_part.alloc(S->grid_n); a |= S->npart.alloc(S->grid_n); a |= S->spart.alloc(S->grid_n);
    a |= S->d_iter.alloc(1); a |= S->d_step.alloc(1); a |= S->ifatm.alloc(NN); a |= S->ifatmp.alloc(NN); a |= S->d_flags.alloc(4);
    DBuf<double> *vs[] = {&S->atmpot, &S->atmact, &S->atmold, &S->pondnod, &S->ovflnod, &S->ovflp};
    for (auto *b : vs) a |= b->alloc(NN);
    a |= S->scal3.alloc(4);
    if (S->bc_any) {
        a |= S->contp_flag.alloc(N); a |= S->contq_flag.alloc(N); a |= S->contp_val.alloc(N); a |= S->qneu.alloc(N);
        a |= S->qlist.alloc(N); a |= S->qpnew.alloc(N); a |= S->qpold.alloc(N); a |= S->contp_list.alloc(N); a |= S->bcsum.alloc(4);
    }
    if (S->sf_n > 0) {
        std::vector<int> nodes((size_t)S->sf_n);
        std::vector<unsigned char> seen((size_t)N, 0);
        for (int i = 0; i < p.nsf; ++i)
            for (int j = p.sf_ptr[i]; j < p.sf_ptr[i + 1]; ++j) {
                int k = p.sf_node[j] - 1;
                if (k < 0 || k >= N) FAIL(-4, "seepage face %d: node %d out of range", i + 1, p.sf_node[j]);
                if (seen[k]) FAIL(-4, "seepage face %d: node %d is listed twice", i + 1, p.sf_node[j]);
                seen[k] = 1;
                if (j > p.sf_ptr[i] && S->hz[nodes[j - 1]] < S->hz[k])
                    FAIL(-4, "input error : elevation values not in descending order on seepage face %d", i + 1);
                nodes[j] = k;
            }
        a |= S->sf_node.upload(nodes); a |= S->sf_ex.alloc(S->sf_n); a |= S->sf_exp.alloc(S->sf_n); a |= S->sf_exit.alloc(S->sf_n);
        a |= S->sf_q.alloc(S->sf_n); a |= S->sf_qp.alloc(S->sf_n); a |= S->d_sf.alloc(1);
    }
    if (a) FAIL(-101, "device allocation failed (N=%d): %s", N, cudaGetErrorString(cudaGetLastError()));
    if (S->dd) {
        std::vector<unsigned char> own((size_t)N, 0);
        for (int l = 0; l <= S->nstr; ++l)
            for (int r = S->own_a; r < S->own_b; ++r)
                for (int j = 0; j < S->nc1; ++j) {
                    unsigned char f = 1;
                    if (r < S->own_a + DD_W) f |= 2;
                    if (r >= S->own_b - DD_W) f |= 4;
                    own[(size_t)l * NN + (size_t)r * S->nc1 + j] = f;
                }
        if (S->own.upload(own)) FAIL(-101, "owned-row mask upload failed");
        DDComm *c = new DDComm();
        S->comm = c;
        const long long hcap = (long long)DD_W * (S->nstr + 1) * S->nc1;
        c->bytes = sizeof(DDBox) + (size_t)4 * hcap * sizeof(double);
        if (S->cm_on) c->bytes += ((size_t)N + 2 * S->cm_halo) * sizeof(double);      // z of k_pcg_tma: the neighbours store their boundary rows into its ghost rows
        CK(cudaMalloc(&c->base, c->bytes));
        CK(cudaMemset(c->base, 0, c->bytes));
        if (S->cm_on) {
            S->cm_z.release();
            S->cm_z.p = (double *)((char *)c->base + sizeof(DDBox) + (size_t)4 * hcap * sizeof(double)) + S->cm_halo;      // non-owning view
            S->cm_z.n = (size_t)N; S->cm_z.pad = S->cm_halo;
            const int rowlen = S->nc1 * (S->nstr + 1);
            const int geom[4] = {S->own_a * rowlen, S->own_b * rowlen, N, DD_W * rowlen};
            CK(cudaMemcpy(&((DDBox *)c->base)->geom[0], geom, sizeof geom, cudaMemcpyHostToDevice));
        }
        CK(cudaMalloc((void **)&c->seq, 2 * sizeof(unsigned int))); CK(cudaMemset(c->seq, 0, 2 * sizeof(unsigned int)));
        CK(cudaMalloc((void **)&c->err, sizeof(int))); CK(cudaMemset(c->err, 0, sizeof(int)));
        CK(cudaMalloc((void **)&c->recv_counter, sizeof(unsigned int))); CK(cudaMemset(c->recv_counter, 0, sizeof(unsigned int)));
        CK(cudaIpcGetMemHandle(&c->handle, c->base));
        DDCtx &x = c->ctx;
        x.world = S->dd_world; x.rank = S->dd_rank;
        x.north = S->dd_rank > 0 ? S->dd_rank - 1 : -1; x.south = S->dd_rank + 1 < S->dd_world ? S->dd_rank + 1 : -1;
        x.me = (DDBox *)c->base; x.inbox_me = (double *)((char *)c->base + sizeof(DDBox));
        for (int r = 0; r < DD_MAXW; ++r) { x.peer[r] = nullptr; x.inbox_peer[r] = nullptr; }
        x.peer[x.rank] = x.me; x.inbox_peer[x.rank] = x.inbox_me;
        x.hcap = hcap; x.nc1 = S->nc1; x.nlay = S->nstr + 1; x.nnod = NN; x.own_a = S->own_a; x.own_b = S->own_b;
        x.seq = c->seq; x.err = c->err;
        CK(cudaDeviceSynchronize());
    }
    CK(cudaMallocHost((void **)&S->h_iter, sizeof(IterOut)));
    CK(cudaMallocHost((void **)&S->h_rb, sizeof(*S->h_rb)));
    CK(cudaMallocHost((void **)&S->h_dt, 2 * sizeof(double)));
    if (S->d_dt.alloc(2)) FAIL(-101, "device allocation failed");
    // graph replay of the Picard iteration: default for the small meshes the cluster solvers take (see picard_iteration)
    S->graph_mode = (S->pcl_c > 0 && !S->newton && !S->dd && p.nlrelx != 2) ? 1 : 0;      // RELXOM takes the iteration number as an argument
    // ... and for the resident PCG kernels when several handles share the GPU (ensemble members, CATHY_PCG_GRID): their launches from
    // concurrent host threads contend for the driver, one graph launch per iteration instead of ~12 calls relieves that
    const bool res_ok = !S->newton && !S->dd && p.nlrelx != 2 && S->pcg_algo == 4 && S->res_rows > 0 && S->pcl_c == 0 && S->pcg_cluster == 0 && !S->tma_on && !S->cm_on;
    if (res_ok && S->pcg_shared_gpu) S->graph_mode = 1;
    if (const char *e = getenv("CATHY_GRAPH")) { const int v = atoi(e); S->graph_mode = v == 0 ? 0 : (S->graph_mode || (v == 1 && res_ok)) ? 1 : 0; }
    CK(cudaMallocHost((void **)&S->h_step, sizeof(StepOut)));
    {
        size_t cnt = (size_t)p.natm * (p.hspatm ? 1 : NN);
        std::vector<double> tab(p.atm_val, p.atm_val + cnt);
        if (S->atmtab.upload(tab)) FAIL(-101, "atmbc table upload failed");
    }
    if (S->surf) { rc = build_surface(S); if (rc) return rc; }
    // ---- initial conditions (SRC/datin.f:380-403, SRC/icvhe.f, icvhwt.f, icvdwt.f) on the host, then upload
    std::vector<double> pt(N, 0.0), pond(NN, 0.0);
    if (p.indp == 0 || p.indp == 1) for (int k = 0; k < N; ++k) pt[k] = p.ic_psi[k];
    if (p.ipond != 0 && prob->ic_pond) for (int k = 0; k < NN; ++k) { pond[k] = prob->ic_pond[k]; if (pond[k] > 0.0) pt[k] = pond[k]; }
    const double *Z = S->hz.data();
    double mult = p.ipond == 0 ? 0.0 : 1.0;
    for (int i = 0; i < NN && p.indp >= 2; ++i)
        for (int k = 0; k <= S->nstr; ++k) {
            size_t kk = (size_t)k * NN + i;
            if (p.indp == 2) pt[kk] = (Z[i] + mult * pond[i]) - Z[kk];
            else if (p.indp == 3) pt[kk] = p.wtposition - Z[kk] + Z[(size_t)S->nstr * NN + i];
            else pt[kk] = Z[i] - Z[kk] - p.wtposition;
        }
    CK(cudaMemcpy(S->ptimep.p, pt.data(), (size_t)N * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(S->pnew.p, pt.data(), (size_t)N * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(S->pondnod.p, pond.data(), (size_t)NN * sizeof(double), cudaMemcpyHostToDevice));
    // SRC/init1.f
    S->deltat = p.deltat; S->dtmin = p.dtmin; S->dtmax = p.dtmax; S->tmax = p.tmax; S->tetaf = p.tetaf;
    if (S->deltat > S->dtmax) S->deltat = S->dtmax;
    if (S->deltat <= S->dtmin) { S->deltat = S->dtmin; S->dtgmin = 0; } else S->dtgmin = 1;
    S->timep = 0.0; S->time = S->deltat;
    S->ponding = p.ipond != 0; S->pondp = S->ponding;
    return 0;
}

// ATMONE (SRC/atmone.f) + MBINIT + CHVELO/STORCAL (SRC/cathy_main.f:2658-2660); also used by cathy_set_psi
static int init_atm_and_storage(CathySim *S)
{
    const CathyProblem &p = S->p;
    const int NN = S->nnod, N = S->n;
    S->htiatm = 0; S->atmtim[0] = S->atmtim[1] = S->atmtim[2] = 0.0; S->atmrec[0] = S->atmrec[1] = S->atmrec[2] = -1; S->atm_next = 0;
    if (S->bc_any) {   // BCONE x2 (SRC/inital.f: NATM,NSF DIRICHLET / NEUMANN)
        int wd = -1, wn = -1;
        bc_one(S->dir, S->time, wd); bc_one(S->neu, S->time, wn);
        S->dir.active = S->neu.active = -2;
        int rcb = bc_upload(S, wd, wn);
        if (rcb) return rcb;
        CK(cudaMemsetAsync(S->qpold.p, 0, (size_t)N * sizeof(double), S->st));
        CK(cudaMemsetAsync(S->qpnew.p, 0, (size_t)N * sizeof(double), S->st));
    }
    if (S->sf_n > 0) {   // SFINIT (SRC/inital.f:215-230): SFQP = 0, exit points from the initial heads
        CK(cudaMemsetAsync(S->sf_q.p, 0, (size_t)S->sf_n * sizeof(double), S->st));
        CK(cudaMemsetAsync(S->sf_qp.p, 0, (size_t)S->sf_n * sizeof(double), S->st));
        LAUNCH(S, k_sf_init, nblk(S->sf_n, S->grid_n), RED_BLOCK, S->sf_n, S->sf_node.p, S->sf_ex.p, S->sf_exp.p, S->sf_exit.p, S->ptimep.p, S->pnew.p);
        S->sfchek = p.isfcvg == 1; S->ksfzer = 1; S->sfflw = S->sfflwp = S->vsfflw = 0.0;
    }
    CK(cudaMemsetAsync(S->atmpot.p, 0, (size_t)NN * sizeof(double), S->st));
    CK(cudaMemsetAsync(S->atmact.p, 0, (size_t)NN * sizeof(double), S->st));
    CK(cudaMemsetAsync(S->atmold.p, 0, (size_t)NN * sizeof(double), S->st));
    if (p.atm_none) {
        S->htiatm = 1;
        std::vector<int> m1(NN, -1);
        CK(cudaMemcpy(S->ifatm.p, m1.data(), (size_t)NN * sizeof(int), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(S->ifatmp.p, m1.data(), (size_t)NN * sizeof(int), cudaMemcpyHostToDevice));
    } else {
        CK(cudaMemsetAsync(S->ifatm.p, 0, (size_t)NN * sizeof(int), S->st));
        CK(cudaMemsetAsync(S->ifatmp.p, 0, (size_t)NN * sizeof(int), S->st));
        if (p.natm == 0) S->htiatm = 1;
        else {
            S->atmtim[2] = p.atm_time[0]; S->atmrec[2] = 0; S->atm_next = 1;
            if (S->atmtim[2] <= 0.0) {   // ATMOLD = first record * area: reuse the interpolation kernel with slot 3 only, into ATMOLD
                LAUNCH(S, k_atm_interp, nblk(NN, S->grid_n), RED_BLOCK, NN, S->atmtab.p, p.hspatm == 0 ? 1 : 0, -1, 0, 1, 0.0, 0.0, 0.0, p.ieto,
                       p.scf, S->arenod.p, S->ifatm.p, 0, S->atmold.p, S->atmact.p);
            }
            atm_shift_read(S, S->time);
            atm_interp_launch(S, 1, 2, S->time, 0);
        }
        if (S->have_dir || S->have_neu)
            LAUNCH(S, k_mark_nonatm, nblk(NN, S->grid_n), RED_BLOCK, NN, S->have_dir ? S->contp_flag.p : nullptr,
                   S->have_neu ? S->contq_flag.p : nullptr, S->ifatm.p, S->ifatmp.p);
        if (S->sf_n > 0) LAUNCH(S, k_sf_mark_nonatm, nblk(S->sf_n, S->grid_n), RED_BLOCK, S->sf_n, S->sf_node.p, NN, S->ifatm.p, S->ifatmp.p);
        LAUNCH(S, k_atmone, nblk(NN, S->grid_n), RED_BLOCK, NN, p.pmin, p.pondh_min, p.scf, S->atmpot.p, S->atmold.p, S->atmact.p, S->pnew.p,
               S->ptimep.p, S->ifatm.p, S->ifatmp.p);
    }
    weight_and_copy(S);
    LAUNCH(S, k_mbinit, 1, RED_BLOCK, NN, S->ifatmp.p, S->atmold.p, S->scal3.p, S->dd ? S->own.p : (const unsigned char *)nullptr);
    double h3[3];
    chvelo_launch(S, S->ptimep.p);
    int rc = step_final_sync(S, S->scal3.p);      // partitioned: the three MBINIT sums are combined across ranks with the step scalars
    if (rc) return rc;
    CK(cudaMemcpyAsync(h3, S->scal3.p, 3 * sizeof(double), cudaMemcpyDeviceToHost, S->st));
    CK(cudaStreamSynchronize(S->st));
    S->aactp = h3[0]; S->aninp = h3[1]; S->anoutp = h3[2];
    S->adinp = S->adoutp = S->ndinp = S->ndoutp = S->nninp = S->nnoutp = 0.0;
    if (S->have_neu) {   // MBINIT's NNINP/NNOUTP, after the initial NEUMANN call for free drainage (SRC/cathy_main.f:2691-2708)
        if (S->free_drain) { CK(cudaMemcpyAsync(S->ckrwp.p, S->ckrw.p, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, S->st)); neumann_device(S, S->ckrw.p); }
        LAUNCH(S, k_flux_sums, 1, RED_BLOCK, S->neu.anbc(), S->qlist.p, S->bcsum.p + 2);
        double hb[2];
        CK(cudaMemcpyAsync(hb, S->bcsum.p + 2, 2 * sizeof(double), cudaMemcpyDeviceToHost, S->st));
        CK(cudaStreamSynchronize(S->st));
        S->nninp = hb[0]; S->nnoutp = hb[1];
    }
    S->store0 = S->store1 = S->store2 = S->h_step->store1;
    S->timep_dirty = 1;
    (void)N;
    return 0;
}

int32_t cathy_create(const CathyProblem *prob, CathySim **out)
{
    g_err[0] = 0;
    *out = nullptr;
    if (!prob || prob->abi_version != CATHY_ABI_VERSION) FAIL(-1, "ABI version mismatch");
    if (prob->iopt != 1 && prob->iopt != 2) FAIL(-2, "IOPT=%d: must be 1 (Picard) or 2 (Newton)", prob->iopt);
    if (prob->iopt == 2 && prob->tetaf != 1.0 && prob->tetaf <= 0.0) FAIL(-2, "TETAF must be positive");
    if (prob->kslope != 0 && !(prob->kslope >= 1 && prob->kslope <= 4 && prob->ivghu == 0 && prob->iopt == 1 && prob->dd_world <= 1))
        FAIL(-2, "KSLOPE=%d: chord and localized slopes (1-4) are implemented for van Genuchten curves (IVGHU=0) under Picard on one GPU", prob->kslope);
    if (prob->kslope == 4 && !(prob->pser > prob->psel)) FAIL(-2, "KSLOPE=4 needs PSEL < PSER (parm line PKRL PKRR PSEL PSER)");
    if (!(prob->ivghu >= 0 && prob->ivghu <= 4))
        FAIL(-2, "IVGHU=%d: van Genuchten (0), extended van Genuchten (1), Huyakorn (2, 3) and Brooks-Corey (4) curves are implemented; look-up tables (-1) are not", prob->ivghu);
    if (prob->lump == 0) FAIL(-2, "LUMP=0 (consistent mass matrix) is not implemented on the device");
    if (prob->nlrelx < 0 || prob->nlrelx > 2) FAIL(-2, "NLRELX=%d: must be 0 (none), 1 (constant OMEGA) or 2 (RELXOM)", prob->nlrelx);
    if (prob->nlrelx == 2 && prob->dd_world > 1) FAIL(-2, "NLRELX=2 is not available on a row-block partitioned mesh");
    if (prob->isimgr != 1 && prob->isimgr != 2) FAIL(-2, "ISIMGR=%d not supported", prob->isimgr);
    if (prob->deltat >= 1.0e15) FAIL(-2, "steady-state runs (DELTAT>=1e15) are not implemented");
    if (prob->ituns > CATHY_MAXIT) FAIL(-2, "ITUNS larger than %d", CATHY_MAXIT);
    CathySim *S = new CathySim();
    int rc = create_impl(prob, S);
    if (rc == 0 && !S->dd) rc = init_atm_and_storage(S);   // partitioned handles finish their set-up in cathy_dd_connect (needs the peers)
    if (rc == 0) rc = launch_check(S);
    if (rc) { std::string keep = g_err; cathy_destroy(S); snprintf(g_err, sizeof g_err, "%s", keep.c_str()); return rc; }
    // the caller's arrays are not referenced after this point, except the copied atm_time
    *out = S;
    return 0;
}

int32_t cathy_get_dims(const CathySim *S, int64_t dims[5])
{
    dims[0] = S->nnod; dims[1] = S->n; dims[2] = S->nt; dims[3] = S->nterm; dims[4] = 2 * S->nterm - S->n;
    return 0;
}
int32_t cathy_get_mesh(const CathySim *S, double *x, double *y, double *z, int32_t *tetra)
{
    if (x) memcpy(x, S->hx.data(), (size_t)S->n * sizeof(double));
    if (y) memcpy(y, S->hy.data(), (size_t)S->n * sizeof(double));
    if (z) memcpy(z, S->hz.data(), (size_t)S->n * sizeof(double));
    if (tetra)
        for (int j = 0; j < S->nstr; ++j)
            for (int i = 0; i < S->ntri; ++i) {
                int pr[3][4];
                gen_tets_of_prism(&S->htri[4 * (size_t)i], j * S->nnod, (j + 1) * S->nnod, pr);
                size_t e0 = 3 * ((size_t)j * S->ntri + i);
                for (int w = 0; w < 3; ++w) {
                    for (int q = 0; q < 4; ++q) tetra[5 * (e0 + w) + q] = pr[w][q] + 1;
                    tetra[5 * (e0 + w) + 4] = S->htri[4 * (size_t)i + 3];
                }
            }
    return 0;
}
double cathy_initial_storage(const CathySim *S) { return S->store0; }

int32_t cathy_get_state(CathySim *S, double *psi, double *sw, double *ckrw, double *qtranie, double *pond, double *atmact, double *atmpot,
                        double *ovfl, int32_t *ifatm)
{
    CK(cudaSetDevice(S->p.device));
    CK(cudaStreamSynchronize(S->st));
    size_t bn = (size_t)S->n * sizeof(double), bs = (size_t)S->nnod * sizeof(double);
    // queued on the handle's stream and awaited once: with page-locked destination buffers the copies run back to back at PCIe speed
    if (psi) CK(cudaMemcpyAsync(psi, S->pnew.p, bn, cudaMemcpyDeviceToHost, S->st));
    if (sw) CK(cudaMemcpyAsync(sw, S->sw.p, bn, cudaMemcpyDeviceToHost, S->st));
    if (ckrw) CK(cudaMemcpyAsync(ckrw, S->ckrw.p, bn, cudaMemcpyDeviceToHost, S->st));
    if (qtranie) CK(cudaMemcpyAsync(qtranie, S->qtranie.p, bn, cudaMemcpyDeviceToHost, S->st));
    if (pond) CK(cudaMemcpyAsync(pond, S->pondnod.p, bs, cudaMemcpyDeviceToHost, S->st));
    if (atmact) CK(cudaMemcpyAsync(atmact, S->atmact.p, bs, cudaMemcpyDeviceToHost, S->st));
    if (atmpot) CK(cudaMemcpyAsync(atmpot, S->atmpot.p, bs, cudaMemcpyDeviceToHost, S->st));
    if (ovfl) CK(cudaMemcpyAsync(ovfl, S->ovflnod.p, bs, cudaMemcpyDeviceToHost, S->st));
    if (ifatm) CK(cudaMemcpyAsync(ifatm, S->ifatm.p, (size_t)S->nnod * sizeof(int), cudaMemcpyDeviceToHost, S->st));
    CK(cudaStreamSynchronize(S->st));
    return 0;
}

// Pipelined read-back: the state is snapshotted device-to-device on the compute stream (tens of microseconds), then a second
// stream drains the snapshot to the caller's (page-locked) buffers while the next cathy_step computes.
int32_t cathy_get_state_async(CathySim *S, double *psi, double *sw, double *ckrw, double *qtranie, double *pond, double *atmact, double *atmpot,
                              double *ovfl, int32_t *ifatm)
{
    CK(cudaSetDevice(S->p.device));
    const size_t n = (size_t)S->n, nn = (size_t)S->nnod;
    if (!S->st_copy) {
        CK(cudaStreamCreateWithFlags(&S->st_copy, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&S->ev_snap, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&S->ev_drained, cudaEventDisableTiming));
        if (S->snap.alloc(4 * n + 4 * nn) || S->snap_i.alloc(nn)) FAIL(-101, "cathy_get_state_async: snapshot allocation failed");
    } else
        CK(cudaStreamWaitEvent(S->st, S->ev_drained, 0));       // the previous snapshot must have left the staging buffers
    struct Item { double *host; const double *dev; size_t cnt; };
    const Item items[8] = {{psi, S->pnew.p, n}, {sw, S->sw.p, n}, {ckrw, S->ckrw.p, n}, {qtranie, S->qtranie.p, n},
                           {pond, S->pondnod.p, nn}, {atmact, S->atmact.p, nn}, {atmpot, S->atmpot.p, nn}, {ovfl, S->ovflnod.p, nn}};
    size_t off = 0;
    SnapArgs sa;
    for (int q = 0; q < 8; ++q) {
        sa.src[q] = items[q].dev; sa.dst[q] = items[q].host ? S->snap.p + off : nullptr;
        off += items[q].cnt;
    }
    sa.isrc = S->ifatm.p; sa.idst = ifatm ? S->snap_i.p : nullptr; sa.n = S->n; sa.nn = S->nnod;
    LAUNCH(S, k_snapshot, nblk(S->n, S->grid_n), RED_BLOCK, sa);
    CK(cudaEventRecord(S->ev_snap, S->st));
    CK(cudaStreamWaitEvent(S->st_copy, S->ev_snap, 0));
    // host buffers carved out of ONE block in the staging order (capi.state_buffers(pinned=True) does that): one copy instead of eight
    bool contiguous = true;
    for (int q = 0; q + 1 < 8 && contiguous; ++q) contiguous = items[q].host && items[q + 1].host && items[q].host + items[q].cnt == items[q + 1].host;
    if (contiguous) CK(cudaMemcpyAsync(items[0].host, S->snap.p, (4 * n + 4 * nn) * sizeof(double), cudaMemcpyDeviceToHost, S->st_copy));
    else {
        off = 0;
        for (const Item &it : items) {
            if (it.host) CK(cudaMemcpyAsync(it.host, S->snap.p + off, it.cnt * sizeof(double), cudaMemcpyDeviceToHost, S->st_copy));
            off += it.cnt;
        }
    }
    if (ifatm) CK(cudaMemcpyAsync(ifatm, S->snap_i.p, nn * sizeof(int), cudaMemcpyDeviceToHost, S->st_copy));
    CK(cudaEventRecord(S->ev_drained, S->st_copy));
    return 0;
}
int32_t cathy_state_wait(CathySim *S)
{
    if (!S->st_copy) return 0;
    CK(cudaSetDevice(S->p.device));
    CK(cudaStreamSynchronize(S->st_copy));
    return 0;
}

int32_t cathy_get_velocity(CathySim *S, double *uu, double *vv, double *ww, double *unod, double *vnod, double *wnod)
{
    CK(cudaSetDevice(S->p.device));
    const int n = S->n, nt = S->nt;
    DBuf<double> dx, dy, px, py, pz, du, dv, dw, dun, dvn, dwn;
    DBuf<int> tz;
    std::vector<double> hx(S->hx.begin(), S->hx.end()), hy(S->hy.begin(), S->hy.end());
    std::vector<int> trizone(S->ntri);
    for (int t = 0; t < S->ntri; ++t) trizone[t] = S->htri[4 * (size_t)t + 3] - 1;
    const size_t nsz = (size_t)S->nstr * S->p.nzone;
    if (S->h_perm.size() != 3 * nsz) FAIL(-1, "cathy_get_velocity: conductivity tables are not available");
    std::vector<double> kx(S->h_perm.begin(), S->h_perm.begin() + nsz), ky(S->h_perm.begin() + nsz, S->h_perm.begin() + 2 * nsz),
        kz(S->h_perm.begin() + 2 * nsz, S->h_perm.end());
    int rc = 0;
    rc |= dx.upload(hx); rc |= dy.upload(hy); rc |= tz.upload(trizone); rc |= px.upload(kx); rc |= py.upload(ky); rc |= pz.upload(kz);
    rc |= du.alloc(nt); rc |= dv.alloc(nt); rc |= dw.alloc(nt); rc |= dun.alloc(n); rc |= dvn.alloc(n); rc |= dwn.alloc(n);
    if (rc) FAIL(-101, "cathy_get_velocity: device allocation failed");
    LAUNCH(S, k_vel3d, nblk(nt, 4 * S->grid_n), RED_BLOCK, nt, S->ntri, S->p.nzone, S->tet.p, tz.p, px.p, py.p, pz.p, dx.p, dy.p, S->z.p, S->pnew.p, S->ckrw.p,
           du.p, dv.p, dw.p);
    LAUNCH(S, k_vnod3d, nblk(n, S->grid_n), RED_BLOCK, n, S->plan, du.p, dv.p, dw.p, dun.p, dvn.p, dwn.p);
    CK(cudaStreamSynchronize(S->st));
    if (uu) CK(cudaMemcpy(uu, du.p, (size_t)nt * sizeof(double), cudaMemcpyDeviceToHost));
    if (vv) CK(cudaMemcpy(vv, dv.p, (size_t)nt * sizeof(double), cudaMemcpyDeviceToHost));
    if (ww) CK(cudaMemcpy(ww, dw.p, (size_t)nt * sizeof(double), cudaMemcpyDeviceToHost));
    if (unod) CK(cudaMemcpy(unod, dun.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
    if (vnod) CK(cudaMemcpy(vnod, dvn.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
    if (wnod) CK(cudaMemcpy(wnod, dwn.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
    DBuf<double> *all[] = {&dx, &dy, &px, &py, &pz, &du, &dv, &dw, &dun, &dvn, &dwn};
    for (auto *b : all) b->release();
    tz.release();
    return 0;
}

int32_t cathy_get_recharge(CathySim *S, double *recnod, double *recflow)
{
    CK(cudaSetDevice(S->p.device));
    const int n = S->n, nn = S->nnod;
    std::vector<double> w(n), rec(nn);
    int rc = cathy_get_velocity(S, nullptr, nullptr, nullptr, nullptr, nullptr, w.data());
    if (rc) return rc;
    DBuf<double> dw, dr;
    if (dw.upload(w) || dr.alloc(nn)) FAIL(-101, "cathy_get_recharge: device allocation failed");
    LAUNCH(S, k_recharge, nblk(nn, S->grid_n), RED_BLOCK, nn, S->nstr, S->pnew.p, dw.p, S->arenod.p, dr.p);
    CK(cudaStreamSynchronize(S->st));
    CK(cudaMemcpy(rec.data(), dr.p, (size_t)nn * sizeof(double), cudaMemcpyDeviceToHost));
    dw.release(); dr.release();
    double flow = 0.0;
    for (int s = 0; s < nn; ++s) flow = flow + rec[s];      // RECFLOW: the reference's sequential sum over the surface nodes
    if (recnod) memcpy(recnod, rec.data(), (size_t)nn * sizeof(double));
    if (recflow) *recflow = flow;
    return 0;
}
int32_t cathy_get_wtdepth(CathySim *S, const int32_t *nodvp, int32_t numvp, double *wt)
{
    CK(cudaSetDevice(S->p.device));
    if (numvp <= 0) return 0;
    for (int i = 0; i < numvp; ++i) if (nodvp[i] < 1 || nodvp[i] > S->nnod) FAIL(-1, "cathy_get_wtdepth: NODVP(%d) = %d is not a surface node", i + 1, nodvp[i]);
    DBuf<int> dn; DBuf<double> dwt;
    std::vector<int> hn(nodvp, nodvp + numvp);
    if (dn.upload(hn) || dwt.alloc(numvp)) FAIL(-101, "cathy_get_wtdepth: device allocation failed");
    LAUNCH(S, k_wtdepth, (numvp + 127) / 128, 128, numvp, dn.p, S->nnod, S->nstr, S->z.p, S->pnew.p, dwt.p);
    CK(cudaStreamSynchronize(S->st));
    CK(cudaMemcpy(wt, dwt.p, (size_t)numvp * sizeof(double), cudaMemcpyDeviceToHost));
    dn.release(); dwt.release();
    return 0;
}
int32_t cathy_set_psi(CathySim *S, const double *psi)
{
    if (S->nstep != 1 || S->kbackt != 0 || S->itrtot != 0) FAIL(-1, "cathy_set_psi is only valid before the first step");
    CK(cudaSetDevice(S->p.device));
    CK(cudaMemcpy(S->ptimep.p, psi, (size_t)S->n * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(S->pnew.p, psi, (size_t)S->n * sizeof(double), cudaMemcpyHostToDevice));
    return init_atm_and_storage(S);
}

int32_t cathy_upload_atm_record(CathySim *S, int32_t rec, const double *vals)
{
    if (rec < 0 || rec >= S->p.natm) FAIL(-1, "atm record %d out of range", rec);
    CK(cudaSetDevice(S->p.device));
    size_t w = S->p.hspatm ? 1 : (size_t)S->nnod;
    CK(cudaMemcpyAsync(S->atmtab.p + (size_t)rec * w, vals, w * sizeof(double), cudaMemcpyHostToDevice, S->st));
    return 0;
}

// time loop body, SRC/cathy_main.f:2882-3829
int32_t cathy_step(CathySim *S, CathyStepReport *rep)
{
    const CathyProblem &p = S->p;
    if (S->finished) FAIL(-1, "simulation already finished");
    if (S->dd && !S->comm->connected) FAIL(-1, "row-block partition: call cathy_dd_connect on every rank before stepping");
    CK(cudaSetDevice(p.device));
    memset(rep, 0, sizeof *rep);
    int64_t l0 = S->launches, pi0 = S->pcg_iters, ps0 = S->pcg_solves;
    double pm0 = S->pcg_ms;
    const int NN = S->nnod, N = S->n;
    size_t bn = (size_t)N * sizeof(double), bs = (size_t)NN * sizeof(double);
    CK(cudaEventRecord(S->ev0, S->st));
    {
        int rcb = bc_next_both(S, false);
        if (rcb) return rcb;
        if (S->have_neu) neumann_device(S, S->ckrw.p);
    }
    atmnxt(S);
    CK(cudaMemsetAsync(S->d_flags.p + 1, 0, sizeof(int), S->st));
    LAUNCH(S, k_etran, nblk(NN, S->grid_n), RED_BLOCK, NN, S->nstr, S->z.p, S->pnew.p, S->atmpot.p, S->veg.p, S->vegpar.p, p.scf, S->qtranie.p, S->d_flags.p + 1);
    if (!S->surf) LAUNCH(S, k_switch_old, nblk(NN, S->grid_n), RED_BLOCK, NN, p.pmin, S->atmpot.p, S->ifatm.p, S->atmact.p, S->pnew.p);
    else LAUNCH(S, k_adrstn, nblk(NN, S->grid_n), RED_BLOCK, NN, p.pmin, S->atmpot.p, S->ifatm.p, S->atmact.p, S->pnew.p);
    weight_and_copy(S);
    int nsurf = 0, status = 0;
    S->attempts.clear();
    for (;;) {
        nsurf = 0;
        if (S->surf && S->ponding) {
            CK(cudaMemcpyAsync(S->d_akmax.p + 2, S->d_akmax.p, sizeof(double), cudaMemcpyDeviceToDevice, S->st));   // AK_MAX_SAV
            copy_cells(S, S->q_in_kk_sav, S->q_in_kk); copy_cells(S, S->q_out_kk_1_sav, S->q_out_kk_1);
            copy_cells(S, S->q_out_kk_2_sav, S->q_out_kk_2); copy_cells(S, S->volume_kk_sav, S->volume_kk);
            nsurf = surf_flowtra(S);
            if (nsurf < 0) return nsurf;
            S->nsurft += nsurf;
        }
        CK(cudaMemcpyAsync(S->pold.p, S->pnew.p, bs, cudaMemcpyDeviceToDevice, S->st));   // VCOPYR(NNOD,POLD,PNEW), SRC/cathy_main.f:3099
        int rc = flow3d(S, &status);
        if (rc) return rc;
        if (status != 1) break;
        if (S->surf) {
            CK(cudaMemcpyAsync(S->d_akmax.p + 1, S->d_akmax.p + 2, sizeof(double), cudaMemcpyDeviceToDevice, S->st));   // AK_MAX_P = AK_MAX_SAV
            copy_cells(S, S->q_in_kk_p, S->q_in_kk_sav); copy_cells(S, S->q_out_kk_1_p, S->q_out_kk_1_sav);
            copy_cells(S, S->q_out_kk_2_p, S->q_out_kk_2_sav); copy_cells(S, S->volume_kk_p, S->volume_kk_sav);
            zero_cells(S, S->q_in_kkp1); zero_cells(S, S->q_out_kkp1_1); zero_cells(S, S->q_out_kkp1_2); zero_cells(S, S->volume_kkp1);
        }
        {   // the failed attempt, as output/iter lists it
            CathySim::Attempt a;
            a.deltat = S->deltat; a.time = S->time; a.n = std::min(S->iter, (int)CATHY_MAXIT);
            memcpy(a.rec, S->itrec, sizeof(CathyIterRecord) * a.n);
            S->attempts.push_back(a);
        }
        bkstep(S);
        if (S->have_neu) neumann_device(S, S->ckrwp.p);
    }
    if (S->surf) LAUNCH(S, k_pond_zero, nblk(NN, S->grid_n), RED_BLOCK, NN, S->pnew.p, S->pondnod.p);
    chvelo_launch(S, S->pnew.p);
    if (S->surf) {
        CK(cudaMemcpyAsync(&S->d_step.p->q_out1, S->q_out_kkp1_1.p + S->outlet_cell, sizeof(double), cudaMemcpyDeviceToDevice, S->st));
        CK(cudaMemcpyAsync(&S->d_step.p->q_out2, S->q_out_kkp1_2.p + S->outlet_cell, sizeof(double), cudaMemcpyDeviceToDevice, S->st));
        CK(cudaMemcpyAsync(&S->d_step.p->ak_max, S->d_akmax.p, sizeof(double), cudaMemcpyDeviceToDevice, S->st));
    }
    // end-of-step copies (SRC/cathy_main.f:3762-3790) are queued before the single synchronisation of the step
    {
        int nbs = nblk(NN, S->grid_n);
        LAUNCH(S, k_step_partial, nbs, RED_BLOCK, NN, S->nstr, p.pmin, p.pondh_min, S->ifatm.p, S->atmpot.p, S->atmact.p, S->pnew.p, S->spart.p, S->dd ? S->own.p : (const unsigned char *)nullptr);
        LAUNCH(S, k_step_final, 1, RED_BLOCK, nbs, S->spart.p, S->grid_n, S->store_part.p, S->d_step.p);
        if (S->dd) LAUNCH(S, k_dd_combine_step, 1, 32, S->comm->ctx, S->d_step.p, (double *)nullptr);
    }
    CK(cudaMemcpyAsync(S->h_step, S->d_step.p, sizeof(StepOut), cudaMemcpyDeviceToHost, S->st));
    CK(cudaMemcpyAsync(S->ifatmp.p, S->ifatm.p, (size_t)NN * sizeof(int), cudaMemcpyDeviceToDevice, S->st));
    CK(cudaMemcpyAsync(S->atmold.p, S->atmact.p, bs, cudaMemcpyDeviceToDevice, S->st));
    CK(cudaMemcpyAsync(S->ptimep.p, S->pnew.p, bn, cudaMemcpyDeviceToDevice, S->st));
    if (S->have_dir) CK(cudaMemcpyAsync(S->qpold.p, S->qpnew.p, (size_t)S->dir.anbc() * sizeof(double), cudaMemcpyDeviceToDevice, S->st));
    if (S->free_drain) CK(cudaMemcpyAsync(S->ckrwp.p, S->ckrw.p, bn, cudaMemcpyDeviceToDevice, S->st));
    if (S->sf_n > 0) {   // SFEXP = SFEXIT = SFEX, SFQP = SFQ (SRC/cathy_main.f:3294-3300)
        CK(cudaMemcpyAsync(S->sf_exp.p, S->sf_ex.p, (size_t)S->sf_n * sizeof(int), cudaMemcpyDeviceToDevice, S->st));
        CK(cudaMemcpyAsync(S->sf_exit.p, S->sf_ex.p, (size_t)S->sf_n * sizeof(int), cudaMemcpyDeviceToDevice, S->st));
        CK(cudaMemcpyAsync(S->sf_qp.p, S->sf_q.p, (size_t)S->sf_n * sizeof(double), cudaMemcpyDeviceToDevice, S->st));
    }
    S->timep_dirty = 1;
    if (S->surf) {
        S->pondp = S->ponding;
        CK(cudaMemcpyAsync(S->ovflp.p, S->ovflnod.p, bs, cudaMemcpyDeviceToDevice, S->st));
        copy_cells(S, S->q_in_kk, S->q_in_kkp1); zero_cells(S, S->q_in_kkp1);
        copy_cells(S, S->q_out_kk_1, S->q_out_kkp1_1); zero_cells(S, S->q_out_kkp1_1);
        copy_cells(S, S->q_out_kk_2, S->q_out_kkp1_2); zero_cells(S, S->q_out_kkp1_2);
        copy_cells(S, S->volume_kk, S->volume_kkp1); zero_cells(S, S->volume_kkp1);
    }
    int h_err = 0;
    CK(cudaMemcpyAsync(&h_err, S->d_flags.p + 1, sizeof(int), cudaMemcpyDeviceToHost, S->st));
    CK(cudaEventRecord(S->ev1, S->st));
    CK(cudaStreamSynchronize(S->st));
    if (h_err) FAIL(-5, "ETRAN: ZROOT reaches the bottom layer (decrease ZROOT)");
    { int lc = launch_check(S); if (lc) return lc; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, S->ev0, S->ev1);
    const StepOut &so = *S->h_step;
    S->store1 = so.store1; S->store2 = S->store2 + S->dstore;
    for (int q = 0; q < 9; ++q) S->hgflag[q] += so.hgflag[q];
    rep->nstep = S->nstep; rep->deltat = S->deltat; rep->time = S->time; rep->iter = S->iter; rep->nitert = S->nitert;
    rep->kbackt = S->kbackt; rep->nsurf = nsurf; rep->nsurft = S->nsurft; rep->noback = status == 2; rep->ponding = S->ponding;
    rep->store1 = S->store1; rep->store2 = S->store2; rep->dstore = S->dstore; rep->vin = S->vin; rep->vout = S->vout;
    rep->erras = S->erras; rep->errel = S->errel; rep->adin = S->adin; rep->adout = S->adout; rep->anin = S->anin; rep->anout = S->anout;
    rep->ndin = S->ndin; rep->ndout = S->ndout; rep->nnin = S->nnin; rep->nnout = S->nnout;
    rep->vndin = S->vndin; rep->vndout = S->vndout; rep->vnnin = S->vnnin; rep->vnnout = S->vnnout;
    rep->sfflw = S->sfflw; rep->vsfflw = S->vsfflw;
    rep->apot = so.apot; rep->aact = so.aact; rep->ovflow = so.ovflow; rep->reflow = so.reflow;
    rep->aact_prev = S->aactp; S->aactp = so.aact;      // AACTP = AACT, SRC/cathy_main.f:3695
    rep->areatot = S->areatot;
    const double NNg = S->dd ? (double)S->gnnod : (double)NN;
    rep->fhort = (double)so.nhort / NNg; rep->fdunn = (double)so.ndunn / NNg; rep->fpond = (double)so.npond / NNg; rep->fsat = (double)so.nsat / NNg;
    rep->n_iter_rec = std::min(S->iter, CATHY_MAXIT);
    memcpy(rep->it, S->itrec, sizeof(CathyIterRecord) * rep->n_iter_rec);
    rep->klsfai_total = S->klsfai; rep->kback_total = S->kback;
    if (S->surf) { rep->q_outlet_1 = so.q_out1; rep->q_outlet_2 = so.q_out2; rep->ak_max = so.ak_max; }
    S->adinp = S->adin; S->adoutp = S->adout; S->aninp = S->anin; S->anoutp = S->anout;
    S->ndinp = S->ndin; S->ndoutp = S->ndout; S->nninp = S->nnin; S->nnoutp = S->nnout;
    S->sfflwp = S->sfflw;
    if (status == 2) S->finished = 1;
    else if (std::fabs(S->time - S->tmax) <= 0.001 * S->deltat) S->finished = 1;
    else {   // TIMUPD + TIMNXT (SRC/timupd.f, SRC/timnxt.f)
        S->timep = S->time;
        if (S->iter < p.ituns1) { S->deltat = S->deltat * p.dtmagm + p.dtmaga; if (S->deltat > S->dtmax) S->deltat = S->dtmax; }
        if (S->iter >= p.ituns2) { S->deltat = S->deltat * p.dtredm - p.dtreds; if (S->deltat < S->dtmin) S->deltat = S->dtmin; }
        if ((S->time + S->deltat) >= S->tmax) { S->deltat = S->tmax - S->time; S->time = S->tmax; }
        else { if ((S->time + 2 * S->deltat) > S->tmax) S->deltat = (S->tmax - S->time) / 2; S->time = S->time + S->deltat; }
        S->dtgmin = !(S->deltat <= S->dtmin);
        S->nstep++; S->iter = 1; S->nitert = 0; S->kbackt = 0; S->nsurft = 0;
        if (!(S->time <= S->tmax)) S->finished = 1;
    }
    rep->finished = S->finished; rep->next_deltat = S->deltat; rep->next_time = S->time;
    rep->itrtot = S->itrtot;
    for (int q = 0; q < 9; ++q) rep->hgflag[q] = S->hgflag[q];
    rep->gpu_ms = ms; rep->launches = S->launches - l0;
    rep->pcg_ms = S->pcg_ms - pm0; rep->pcg_iters = S->pcg_iters - pi0; rep->pcg_solves = S->pcg_solves - ps0;
    return 0;
}


// ---- in-process ensemble support (data assimilation restarts) ----------------------------------
int32_t cathy_pack_state(CathySim *S, int32_t which, double *dX, int64_t ld, int64_t col)
{
    if (which != 0 && which != 1) FAIL(-1, "cathy_pack_state: which must be 0 (psi) or 1 (sw)");
    CK(cudaSetDevice(S->p.device));
    LAUNCH(S, k_pack_col, nblk(S->n, S->grid_n), RED_BLOCK, S->n, which == 0 ? S->pnew.p : S->sw.p, dX, (long long)ld, (long long)col);
    CK(cudaStreamSynchronize(S->st));   // the matrix is consumed on the caller's stream
    return 0;
}
int32_t cathy_unpack_psi(CathySim *S, const double *dX, int64_t ld, int64_t col)
{
    CK(cudaSetDevice(S->p.device));
    LAUNCH(S, k_unpack_col, nblk(S->n, S->grid_n), RED_BLOCK, S->n, dX, (long long)ld, (long long)col, S->pnew.p, S->ptimep.p);
    CK(cudaStreamSynchronize(S->st));
    return 0;
}
// Start a new run from the CURRENT pressure heads (device resident) at time 0 with a new TMAX: what pyCATHY does
// between assimilation windows by rewriting input/ic + input/parm and relaunching the processor
// (pyCATHY/DA/cathy_DA.py:1863-1875 update_ENS_files, pyCATHY/cathy_tools.py:593-740 run_processor).
int32_t cathy_restart(CathySim *S, double tmax, double deltat)
{
    S->graph_drop();      // frozen launch arguments may refer to what this call replaces
    CK(cudaSetDevice(S->p.device));
    const CathyProblem &p = S->p;
    if (tmax > 0.0) S->p.tmax = tmax;
    if (deltat > 0.0) S->p.deltat = deltat;
    CK(cudaMemcpyAsync(S->ptimep.p, S->pnew.p, (size_t)S->n * sizeof(double), cudaMemcpyDeviceToDevice, S->st));
    // SRC/init1.f
    S->deltat = p.deltat; S->dtmin = p.dtmin; S->dtmax = p.dtmax; S->tmax = p.tmax; S->tetaf = p.tetaf;
    if (S->deltat > S->dtmax) S->deltat = S->dtmax;
    if (S->deltat <= S->dtmin) { S->deltat = S->dtmin; S->dtgmin = 0; } else S->dtgmin = 1;
    S->timep = 0.0; S->time = S->deltat;
    S->nstep = 1; S->iter = 1; S->nitert = 0; S->itlin = 0; S->itrtot = 0; S->kbackt = 0; S->kback = 0; S->klsfai = 0; S->nsurft = 0;
    S->finished = 0; S->lsfail = 0;
    S->ponding = p.ipond != 0; S->pondp = S->ponding;
    for (int q = 0; q < 9; ++q) S->hgflag[q] = 0;
    CK(cudaMemsetAsync(S->pondnod.p, 0, (size_t)S->nnod * sizeof(double), S->st));
    CK(cudaMemsetAsync(S->ovflnod.p, 0, (size_t)S->nnod * sizeof(double), S->st));
    CK(cudaMemsetAsync(S->ovflp.p, 0, (size_t)S->nnod * sizeof(double), S->st));
    CK(cudaMemsetAsync(S->qtranie.p, 0, (size_t)S->n * sizeof(double), S->st));
    if (S->surf) {
        DBuf<double> *bufs[] = {&S->sw_sn, &S->q_in_kk, &S->q_in_kkp1, &S->q_out_kk_1, &S->q_out_kk_2, &S->q_out_kkp1_1, &S->q_out_kkp1_2,
                                &S->volume_kk, &S->volume_kkp1, &S->h_water, &S->q_in_kk_sav, &S->q_out_kk_1_sav, &S->q_out_kk_2_sav,
                                &S->volume_kk_sav, &S->q_in_kk_p, &S->q_out_kk_1_p, &S->q_out_kk_2_p, &S->volume_kk_p};
        for (auto *b : bufs) zero_cells(S, *b);
        CK(cudaMemsetAsync(S->d_akmax.p, 0, 3 * sizeof(double), S->st));
    }
    return init_atm_and_storage(S);
}
// Replace the soil tables ([nstr][nzone] each, SRC/datin.f:510-514) -- the parameter update of the DA analysis
// (pyCATHY/cathy_tools.py:3043-3130 update_soil).  Nodal constants and the assembly coefficients are rebuilt.
int32_t cathy_set_soil(CathySim *S, const double *permx, const double *permy, const double *permz, const double *elstor,
                       const double *poros, const double *vgn, const double *vgrmc, const double *vgpsat)
{
    S->graph_drop();      // frozen launch arguments may refer to what this call replaces
    CK(cudaSetDevice(S->p.device));
    CK(cudaStreamSynchronize(S->st));
    CathyProblem keep = S->p;
    S->p.permx = permx; S->p.permy = permy; S->p.permz = permz; S->p.elstor = elstor; S->p.poros = poros;
    S->p.vgn = vgn; S->p.vgrmc = vgrmc; S->p.vgpsat = vgpsat;
    S->p.dem = S->h_dem.data(); S->p.zone = S->h_zone.data(); S->p.root_map = S->h_root.data(); S->p.zratio = S->h_zratio.data();
    S->p.pcana = S->h_veg.data(); S->p.pcref = S->h_veg.data() + keep.nveg; S->p.pcwlt = S->h_veg.data() + 2 * keep.nveg;
    S->p.zroot = S->h_veg.data() + 3 * keep.nveg; S->p.pz = S->h_veg.data() + 4 * keep.nveg; S->p.omgc = S->h_veg.data() + 5 * keep.nveg;
    int rc = build_static(S);
    S->timep_dirty = 1;
    return rc;
}
// Replace the atmospheric forcing table (times and rates) -- the per-window input/atmbc rewrite of the DA loop
// (pyCATHY/DA/cathy_DA.py update_ENS_files -> update_atmbc).  Takes effect at the next cathy_restart.
int32_t cathy_set_atm_table(CathySim *S, int32_t natm, const double *times, const double *vals)
{
    if (natm < 0) FAIL(-1, "cathy_set_atm_table: natm < 0");
    CK(cudaSetDevice(S->p.device));
    CK(cudaStreamSynchronize(S->st));
    delete[] S->p.atm_time;
    double *t = new double[std::max(natm, 1)];
    for (int i = 0; i < natm; ++i) t[i] = times[i];
    S->p.atm_time = t; S->p.natm = natm;
    size_t cnt = (size_t)natm * (S->p.hspatm ? 1 : S->nnod);
    std::vector<double> tab(vals, vals + cnt);
    if (S->atmtab.upload(tab)) FAIL(-101, "atmbc table upload failed");
    return 0;
}

// ---- row-block partition: peer-memory wiring -------------------------------------------------
// k_pcg_tma: where the neighbours' z arrays are and where my boundary rows go in them (their geometry sits in their DDBox)
static int dd_wire_tma(CathySim *S)
{
    if (!S->cm_on) return 0;
    const DDCtx &x = S->comm->ctx;
    const long long hcap = x.hcap;
    auto zof = [&](DDBox *box) { return (double *)((char *)box + sizeof(DDBox) + (size_t)4 * hcap * sizeof(double)) + S->cm_halo; };
    int g[4];
    S->tma_zpeer_n = S->tma_zpeer_s = nullptr;
    if (x.north >= 0) {
        CK(cudaMemcpy(g, &x.peer[x.north]->geom[0], sizeof g, cudaMemcpyDeviceToHost));
        S->tma_zpeer_n = zof(x.peer[x.north]); S->tma_ndst0 = g[1];                 // its south ghost rows start at its hi
    }
    if (x.south >= 0) {
        CK(cudaMemcpy(g, &x.peer[x.south]->geom[0], sizeof g, cudaMemcpyDeviceToHost));
        S->tma_zpeer_s = zof(x.peer[x.south]); S->tma_sdst0 = (long long)g[0] - g[3];   // its north ghost rows end at its lo
    }
    return 0;
}
int32_t cathy_dd_export(CathySim *S, void *handle64)
{
    if (!S->dd) FAIL(-1, "cathy_dd_export: the handle is not partitioned (dd_world = 1)");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    memcpy(handle64, &S->comm->handle, 64);
    return 0;
}
int32_t cathy_dd_connect(CathySim *S, const void *handles)
{
    if (!S->dd) FAIL(-1, "cathy_dd_connect: the handle is not partitioned (dd_world = 1)");
    CK(cudaSetDevice(S->p.device));
    DDComm *c = S->comm;
    for (int r = 0; r < S->dd_world; ++r) {
        if (r == S->dd_rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + (size_t)64 * r, 64);
        void *ptr = nullptr;
        CK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer_base[r] = ptr; c->opened[r] = true;
        c->ctx.peer[r] = (DDBox *)ptr;
        c->ctx.inbox_peer[r] = (double *)((char *)ptr + sizeof(DDBox));
    }
    c->connected = true;
    { int rct = dd_wire_tma(S); if (rct) return rct; }
    return init_atm_and_storage(S);      // collective: ATMONE / MBINIT / initial storage sums are combined across the ranks
}
// Same-process variant (all ranks are handles of ONE process, e.g. one host thread per GPU): direct pointers, no IPC.
int32_t cathy_dd_connect_local(CathySim *S, CathySim *const *all)
{
    if (!S->dd) FAIL(-1, "cathy_dd_connect_local: the handle is not partitioned (dd_world = 1)");
    CK(cudaSetDevice(S->p.device));
    DDComm *c = S->comm;
    for (int r = 0; r < S->dd_world; ++r) {
        if (r == S->dd_rank) continue;
        CathySim *o = all[r];
        if (!o || !o->dd || o->dd_rank != r || o->dd_world != S->dd_world) FAIL(-1, "cathy_dd_connect_local: handle %d is not rank %d of this partition", r, r);
        if (o->p.device != S->p.device) {
            cudaError_t e = cudaDeviceEnablePeerAccess(o->p.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) FAIL(-100, "peer access %d -> %d: %s", S->p.device, o->p.device, cudaGetErrorString(e));
            cudaGetLastError();
        }
        c->ctx.peer[r] = (DDBox *)o->comm->base;
        c->ctx.inbox_peer[r] = (double *)((char *)o->comm->base + sizeof(DDBox));
    }
    c->connected = true;
    return dd_wire_tma(S);
}
// second half of the local connection: the collective part of the set-up, to be called concurrently (one thread per handle)
int32_t cathy_dd_start(CathySim *S)
{
    if (!S->dd || !S->comm->connected) FAIL(-1, "cathy_dd_start: connect first");
    CK(cudaSetDevice(S->p.device));
    return init_atm_and_storage(S);
}
int32_t cathy_solver_info(const CathySim *S, int64_t info[4])
{
    const bool res = !S->dd && (S->pcg_algo == 3 || S->pcg_algo == 4) && S->res_rows > 0;
    info[0] = S->newton ? (S->bres_rows > 0 ? 11 : 10) : S->pcl_c > 0 ? (S->pcl_v2 ? 8 : 7) : res ? S->pcg_algo : (!S->dd && S->pcg_algo == 2) ? 2 : S->tma_on ? 6 : S->cm_on ? 5 : 1;
    info[1] = res ? S->res_rows : (S->newton ? S->bres_rows : 0); info[2] = res ? S->res_x : 0; info[3] = S->grid_pcg;
    if (!S->newton && S->pcl_c > 0) { info[1] = S->pcl_rows; info[2] = 1; info[3] = S->pcl_c; }
    return 0;
}
int32_t cathy_solver_limits(const CathySim *S, double lim[5])
{
    lim[0] = (double)S->itmax_dev; lim[1] = S->tol_dev; lim[2] = S->itmxcg_scale; lim[3] = S->tolcg_scale;
    lim[4] = S->newton ? ((S->bicg_line || S->bres_rows > 0) ? 2.0 : 1.0) : 1.0;      // preconditioner: 1 = diagonal (point Jacobi), 2 = vertical line (tridiagonal per DEM column)
    return 0;
}
int32_t cathy_plan_info(const CathySim *S, int64_t info[2])
{
    info[0] = S->geom.rel ? 1 : 0;      // 1: assembly derives the tet indices (k_assemble_a), 0: stored index lists (k_assemble)
    info[1] = S->geom.rel ? S->geom.wrel : 0;
    return 0;
}
int32_t cathy_dd_info(const CathySim *S, int64_t info[8])
{
    info[0] = S->grow0; info[1] = S->nrow + 1; info[2] = S->grow0 + S->own_a; info[3] = S->grow0 + S->own_b;
    info[4] = S->nnod; info[5] = S->n;
    info[6] = S->dd ? S->gnnod : S->nnod; info[7] = S->dd ? (int64_t)S->gnnod * (S->nstr + 1) : S->n;
    if (!S->dd) { info[2] = 0; info[3] = S->nrow + 1; }
    return 0;
}

// ---- kernel-level entry points ------------------------------------------------------------
int32_t cathy_debug_assemble(CathySim *S, double deltat, int32_t *topol, int32_t *ja, double *coef1, double *rhs)
{
    CK(cudaSetDevice(S->p.device));
    S->timep_dirty = 1;
    if (S->newton) {   // Jacobian in full CSR, rows ascending (SRC/strnew.f layout); Dirichlet diagonals carry the reference's penalty
        int rcn = assemble_system_newton(S, deltat);
        if (rcn) return rcn;
        CK(cudaStreamSynchronize(S->st));
        const int n = S->n;
        std::vector<double> hU, hL, hdi;
        if (coef1) {
            hU.resize((size_t)NDIAG * S->ld); hL.resize((size_t)NDIAG * S->ld); hdi.resize(n);
            CK(cudaMemcpy(hU.data(), S->Ju.p, hU.size() * sizeof(double), cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(hL.data(), S->Jl.p, hL.size() * sizeof(double), cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(hdi.data(), S->dinv.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
        }
        int64_t m = 0;
        for (int k = 0; k < n; ++k) {
            if (topol) topol[k] = (int32_t)(m + 1);
            for (int d = NDIAG - 1; d >= 1; --d) {
                int c = k - S->off[d];
                if (c < 0 || !S->hexist[(size_t)d * n + c]) continue;
                if (ja) ja[m] = c + 1;
                if (coef1) coef1[m] = hL[(size_t)d * S->ld + c];
                ++m;
            }
            for (int d = 0; d < NDIAG; ++d) {
                if (d > 0 && !S->hexist[(size_t)d * n + k]) continue;
                if (ja) ja[m] = k + S->off[d] + 1;
                if (coef1) coef1[m] = (d == 0 && hdi[k] == 0.0) ? 1.0e-9 * RMAX_ : hU[(size_t)d * S->ld + k];
                ++m;
            }
        }
        if (topol) topol[n] = (int32_t)(m + 1);
        if (rhs) CK(cudaMemcpy(rhs, S->rhs.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
        return 0;
    }
    int rc = assemble_system(S, deltat);
    if (rc) return rc;
    CK(cudaStreamSynchronize(S->st));
    const int n = S->n;
    std::vector<double> hA, hd;
    if (coef1) {
        hA.resize((size_t)NDIAG * S->ld); hd.resize(n);
        CK(cudaMemcpy(hA.data(), S->A.p, hA.size() * sizeof(double), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hd.data(), S->diag_bc.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
    }
    int64_t m = 0;
    for (int k = 0; k < n; ++k) {
        if (topol) topol[k] = (int32_t)(m + 1);
        for (int d = 0; d < NDIAG; ++d) {
            if (d > 0 && !S->hexist[(size_t)d * n + k]) continue;
            if (ja) ja[m] = k + S->off[d] + 1;
            if (coef1) coef1[m] = d == 0 ? hd[k] : hA[(size_t)d * S->ld + k];
            ++m;
        }
    }
    if (topol) topol[n] = (int32_t)(m + 1);
    if (rhs) CK(cudaMemcpy(rhs, S->rhs.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}
int32_t cathy_debug_spmv(CathySim *S, const double *x, double *y, int32_t reps, double *ms)
{
    CK(cudaSetDevice(S->p.device));
    const int n = S->n;
    CK(cudaMemcpy(S->wp0.p, x, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
    Diag A = make_diag(S, S->A.p);
    reps = std::max(reps, 1);
    if (S->cm_on && !S->dd) {
        // large meshes: the product runs in the numbering the solver uses (column-major permutation, see create_impl): matrix, diagonal
        // and x are permuted once, the timed launches stream the permuted arrays, y comes back in the reference numbering
        const int NN = S->nnod, L = S->nstr + 1;
        Diag P;
        for (int d = 0; d < NDIAG; ++d) { P.d[d] = S->cm_A.p + (size_t)d * S->ld; P.off[d] = S->cm_off[d]; }
        static const int newd[NDIAG] = {0, 3, 5, 7, 6, 4, 2, 1};
        PermArgs pa;
        int q = 0;
        for (int d = 1; d < NDIAG; ++d) {
            const bool swp = d >= 4 && d <= 6;
            pa.src[q] = A.d[d]; pa.dst[q] = P.d[newd[d]]; pa.shift[q] = swp ? S->cm_off[newd[d]] : 0; ++q;
        }
        pa.src[q] = S->diag_bc.p; pa.dst[q] = S->cm_diag.p; pa.shift[q] = 0; ++q;
        pa.src[q] = S->wp0.p; pa.dst[q] = S->cm_p0.p; pa.shift[q] = 0; ++q;
        pa.nnod = NN; pa.nl = L; pa.n = n;
        const size_t tile = (size_t)L * 33 * sizeof(double);
        k_permute_cols<<<dim3((NN + 31) / 32, q), 256, tile, S->st>>>(pa);
        CK(cudaGetLastError());
        if (S->tma_on && !getenv("CATHY_SPMV_PLAIN")) {      // the product with the solver's own staging (k_spmv_tma, pcg_tma.cuh)
            TmaArgs ta;
            memset(&ta, 0, sizeof ta);
            ta.n = n; ta.lo = 0; ta.hi = n; ta.A = P; ta.dg = S->cm_diag.p; ta.nl = L;
            const double *xin = S->cm_p0.p;
            double *yout = S->cm_bv.p;
            void *targs[] = {&ta, &xin, &yout};
            CK(cudaLaunchKernel((const void *)k_spmv_tma, dim3(S->sms), dim3(TMA_BLOCK), targs, S->tma_smem, S->st));   // warm-up
            CK(cudaEventRecord(S->ev0, S->st));
            for (int r = 0; r < reps; ++r) { CK(cudaLaunchKernel((const void *)k_spmv_tma, dim3(S->sms), dim3(TMA_BLOCK), targs, S->tma_smem, S->st)); S->launches++; }
            CK(cudaEventRecord(S->ev1, S->st));
        } else {
        LAUNCH(S, k_spmv, nblk(n, S->grid_n), RED_BLOCK, n, P, S->cm_diag.p, S->cm_p0.p, S->cm_bv.p);   // warm-up
        CK(cudaEventRecord(S->ev0, S->st));
        for (int r = 0; r < reps; ++r) LAUNCH(S, k_spmv, nblk(n, S->grid_n), RED_BLOCK, n, P, S->cm_diag.p, S->cm_p0.p, S->cm_bv.p);
        CK(cudaEventRecord(S->ev1, S->st));
        }
        k_unpermute_cols<<<(NN + 31) / 32, 256, tile, S->st>>>(NN, L, S->cm_bv.p, S->wbv.p);
        CK(cudaGetLastError());
    } else {
        LAUNCH(S, k_spmv, nblk(n, S->grid_n), RED_BLOCK, n, A, S->diag_bc.p, S->wp0.p, S->wbv.p);   // warm-up
        CK(cudaEventRecord(S->ev0, S->st));
        for (int r = 0; r < reps; ++r) LAUNCH(S, k_spmv, nblk(n, S->grid_n), RED_BLOCK, n, A, S->diag_bc.p, S->wp0.p, S->wbv.p);
        CK(cudaEventRecord(S->ev1, S->st));
    }
    CK(cudaStreamSynchronize(S->st));
    float t = 0.f;
    cudaEventElapsedTime(&t, S->ev0, S->ev1);
    if (ms) *ms = t / reps;
    CK(cudaMemcpy(y, S->wbv.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}
int32_t cathy_debug_solve(CathySim *S, double *sol, int32_t *niter, double *err, double *ms)
{
    CK(cudaSetDevice(S->p.device));
    CK(cudaEventRecord(S->ev0, S->st));
    int rc = S->newton ? solve_system_newton(S) : solve_system(S);
    if (rc) return rc;
    CK(cudaEventRecord(S->ev1, S->st));
    CK(cudaMemcpyAsync(S->h_iter, S->d_iter.p, sizeof(IterOut), cudaMemcpyDeviceToHost, S->st));
    CK(cudaStreamSynchronize(S->st));
    float t = 0.f;
    cudaEventElapsedTime(&t, S->ev0, S->ev1);
    if (ms) *ms = t;
    S->barrier_epoch = (unsigned int)S->h_iter->pad;
    if (niter) *niter = S->h_iter->pcg_niter;
    if (err) *err = S->h_iter->pcg_err;
    if (sol) CK(cudaMemcpy(sol, S->pdiff.p, (size_t)S->n * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

}  // extern "C"
