// flow_kernels.cuh -- after the solve: head update (SHLPIC), back-calculated fluxes (BKPIC), seepage faces (seepage.cuh), norms + storage change + boundary flux sums (NORMS, STORMB, FLUXMB), atmospheric switching (SWITCH, SWITCH_OLD, ADRSTN, PONDUPD), atmospheric forcing and root uptake (ATMNXT, ETRAN).
// Part of the single translation unit cathy_b200.cu (included in dependency order; shares its structs and helpers).
#pragma once

// ------------------------------------------------------------------------------------------
// after the solve: PNEW += PDIFF and SHLPIC's Dirichlet reset (SRC/picard.f:185-198, SRC/shlpic.f:30-56)
// ------------------------------------------------------------------------------------------
// one launch instead of nine device-to-device copies: the state arrays that cathy_get_state returns, packed into the staging buffer
struct SnapArgs { const double *src[8]; double *dst[8]; const int *isrc; int *idst; int n, nn; };
__global__ void k_snapshot(SnapArgs a)
{
    const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (int k = t0; k < a.n; k += stride) {
#pragma unroll
        for (int q = 0; q < 4; ++q) if (a.dst[q]) a.dst[q][k] = a.src[q][k];
    }
    for (int k = t0; k < a.nn; k += stride) {
#pragma unroll
        for (int q = 4; q < 8; ++q) if (a.dst[q]) a.dst[q][k] = a.src[q][k];
        if (a.idst) a.idst[k] = a.isrc[k];
    }
}
__global__ void k_update(int n, int nnod, const double *__restrict__ pdiff, const double *__restrict__ pold,
                         const int *__restrict__ ifatm, const unsigned char *__restrict__ contp_flag,
                         const double *__restrict__ contp_val, double *__restrict__ pnew)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        double v = pnew[k] + pdiff[k];
        if (contp_flag && contp_flag[k]) v = contp_val[k];
        if (k < nnod) { int f = ifatm[k]; if (f == 1 || f == 2) v = pold[k]; }
        pnew[k] = v;
    }
}

// back-calculated fluxes at atmospheric Dirichlet nodes (BKPIC, SRC/bkpic.f:27-53): only the rows
// that are read afterwards are formed, i.e. one 15-point row product per Dirichlet node.
// row product with the ORIGINAL matrix when its off-diagonals are stored symmetrically scaled (dis != nullptr, see k_pcg2)
__device__ __forceinline__ double dia_row_orig(const Diag &A, const double *__restrict__ diag0, const double *__restrict__ dis,
                                               const double *__restrict__ x, int k, int n)
{
    if (!dis) return dia_row(A, diag0, x, k, n);
    double acc = 0.0;
#pragma unroll
    for (int d = 1; d < NDIAG; ++d) { int j = k + A.off[d]; double dj = dis[j]; if (dj != 0.0) acc += A.d[d][k] * (x[j] / dj); }
#pragma unroll
    for (int d = 1; d < NDIAG; ++d) { int j = k - A.off[d]; double dj = dis[j]; if (dj != 0.0) acc += A.d[d][j] * (x[j] / dj); }
    return diag0[k] * x[k] + acc / dis[k];
}
__global__ void k_bkflux(int n, int nnod, Diag A, const double *__restrict__ diag_true, const double *__restrict__ pdiff,
                         const double *__restrict__ xt5, const int *__restrict__ ifatm, double tetaf,
                         const double *__restrict__ atmold, double *__restrict__ atmact, const double *__restrict__ dis)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nnod; k += gridDim.x * blockDim.x) {
        int f = ifatm[k];
        if (f == 1 || f == 2) {
            double scr = dia_row_orig(A, diag_true, dis, pdiff, k, n) - xt5[k];
            atmact[k] = (scr - (1.0 - tetaf) * atmold[k]) * (1.0 / tetaf);
        }
    }
}
// same for the prescribed-head nodes: QPNEW (SRC/bkpic.f:38-41), indexed by list position like the reference
__global__ void k_bkflux_list(int n, int m, const int *__restrict__ list, Diag A, const double *__restrict__ diag_true,
                              const double *__restrict__ pdiff, const double *__restrict__ xt5, double tetaf,
                              const double *__restrict__ qpold, double *__restrict__ qpnew, const double *__restrict__ dis)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        int k = list[i];
        double scr = dia_row_orig(A, diag_true, dis, pdiff, k, n) - xt5[k];
        qpnew[i] = (scr - (1.0 - tetaf) * qpold[i]) * (1.0 / tetaf);
    }
}
// surface nodes carrying a non-atmospheric BC leave the atmospheric state machine (SRC/atmone.f label 400, SRC/atmnxt.f label 800)
__global__ void k_mark_nonatm(int nnod, const unsigned char *__restrict__ contp_flag, const unsigned char *__restrict__ contq_flag,
                              int *__restrict__ ifatm, int *__restrict__ ifatmp)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nnod; i += gridDim.x * blockDim.x)
        if ((contp_flag && contp_flag[i]) || (contq_flag && contq_flag[i])) { ifatm[i] = -1; if (ifatmp) ifatmp[i] = -1; }
}
// signed sums of a flux list (NDIN/NDOUT, NNIN/NNOUT of SRC/fluxmb.f:29-48), one block, fixed order
__global__ void k_flux_sums(int m, const double *__restrict__ q, double *__restrict__ out2)
{
    __shared__ double sh[32];
    double a = 0.0, b = 0.0;
    for (int k = threadIdx.x; k < m; k += blockDim.x) { double v = q[k]; if (v > 0.0) a += v; else b += v; }
    double t0 = block_sum<RED_BLOCK>(a, sh), t1 = block_sum<RED_BLOCK>(b, sh);
    if (threadIdx.x == 0) { out2[0] = t0; out2[1] = t1; }
}
// free drainage writes into the list-ordered Q array as well as the dense one
__global__ void k_free_drain_list(int nnod, int nstr, const double *__restrict__ arenod, const double *__restrict__ ckrw,
                                  const double *__restrict__ kznod, double *__restrict__ qlist, double *__restrict__ qdense)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nnod; i += gridDim.x * blockDim.x) {
        size_t nn = (size_t)nstr * nnod + i;
        double q = -1.0 * arenod[i] * ckrw[nn] * kznod[nn];
        qlist[i] = q;
        qdense[nn] = q;
    }
}

#include "seepage.cuh"

// norms (NORMS, SRC/norms.f:18-38) + storage change (STORMB, SRC/stormb.f) + boundary flux sums
// (FLUXMB, SRC/fluxmb.f:29-88): block partials in fixed order
struct NormPartial { double pl2, fl2, dstore, pinf, finf, adin, adout, anin, anout; int ik; int pad; };
// pdiff != nullptr: SHLPIC's update PNEW += PDIFF (k_update) is done here, in the same pass (Picard; Newton needs the new heads in
// k_sw_pair first and keeps the separate launch)
__global__ void k_norms(int n, int nnod, double *pnew, const double *__restrict__ pold,
                        const double *__restrict__ rhs, const double *__restrict__ ptimep,
                        const double *__restrict__ swnew, const double *__restrict__ swtimep,
                        const double *__restrict__ volnod, const double *__restrict__ snodi,
                        const double *__restrict__ pnodi, const int *__restrict__ ifatm, const double *__restrict__ atmact,
                        NormPartial *__restrict__ part, const unsigned char *__restrict__ own, double omega,
                        const double *__restrict__ pdiff, const unsigned char *__restrict__ contp_flag, const double *__restrict__ contp_val,
                        const double *__restrict__ omega_dev)
{
    __shared__ double sh[32];
    __shared__ double shv[RED_BLOCK / 32];
    __shared__ int shi[RED_BLOCK / 32];
    if (omega_dev) omega = *omega_dev;      // NLRELX = 2: the relaxation parameter of this iteration was formed on the device (k_relxom_final)
    double pl2 = 0, fl2 = 0, ds = 0, pinf = 0, finf = 0, adin = 0, adout = 0, anin = 0, anout = 0;
    int ik = 0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        double pn;
        if (pdiff) {
            pn = pnew[k] + pdiff[k];
            if (contp_flag && contp_flag[k]) pn = contp_val[k];
            if (k < nnod) { int f = ifatm[k]; if (f == 1 || f == 2) pn = pold[k]; }
            pnew[k] = pn;
        } else
            pn = pnew[k];
        if (own && !(own[k] & 1)) continue;      // row-block partition: ghost rows belong to another rank
        // NLRELX = 1: the norms see the relaxed heads (SRC/relax.f runs between MASBAL and NORMS), the storage change below does not
        const double pr = omega == 1.0 ? pn : (1.0 - omega) * pold[k] + omega * pn;
        double d = pr - pold[k], da = fabs(d), f = rhs[k];
        pl2 += d * d;
        fl2 += f * f;
        if (da > pinf || (da == pinf && k >= ik)) { pinf = da; ik = k; }
        finf = fmax(finf, fabs(f));
        ds += volnod[k] * (snodi[k] * (swnew[k] + swtimep[k]) * 0.5 * (pn - ptimep[k]) + pnodi[k] * (swnew[k] - swtimep[k]));
        if (k < nnod) {
            int fa = ifatm[k];
            if (fa != -1) {
                double a = atmact[k];
                if (fa == 1 || fa == 2) { if (a > 0.0) adin += a; else adout += a; }
                else { if (a > 0.0) anin += a; else anout += a; }
            }
        }
    }
    double t1 = block_sum<RED_BLOCK>(pl2, sh), t2 = block_sum<RED_BLOCK>(fl2, sh), t3 = block_sum<RED_BLOCK>(ds, sh);
    double t4 = block_sum<RED_BLOCK>(adin, sh), t5 = block_sum<RED_BLOCK>(adout, sh), t6 = block_sum<RED_BLOCK>(anin, sh), t7 = block_sum<RED_BLOCK>(anout, sh);
    // max reductions (ties -> larger index, i.e. the LAST node like the sequential >= test)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_down_sync(0xffffffffu, pinf, o);
        int oi = __shfl_down_sync(0xffffffffu, ik, o);
        double of = __shfl_down_sync(0xffffffffu, finf, o);
        if (ov > pinf || (ov == pinf && oi > ik)) { pinf = ov; ik = oi; }
        finf = fmax(finf, of);
    }
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) { shv[w] = pinf; shi[w] = ik; sh[w] = finf; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < RED_BLOCK / 32; ++q) {
            if (shv[q] > pinf || (shv[q] == pinf && shi[q] > ik)) { pinf = shv[q]; ik = shi[q]; }
            finf = fmax(finf, sh[q]);
        }
        NormPartial p;
        p.pl2 = t1; p.fl2 = t2; p.dstore = t3; p.pinf = pinf; p.finf = finf; p.ik = ik; p.pad = 0;
        p.adin = t4; p.adout = t5; p.anin = t6; p.anout = t7;
        part[blockIdx.x] = p;
    }
}
// final fixed-order reduction of the block partials, one block
__global__ void k_norms_final(int nb, const NormPartial *__restrict__ part, const double *__restrict__ pnew,
                              const double *__restrict__ pold, IterOut *__restrict__ out)
{
    __shared__ double sh[32];
    __shared__ double shv[RED_BLOCK / 32];
    __shared__ int shi[RED_BLOCK / 32];
    double pl2 = 0, fl2 = 0, ds = 0, pinf = 0, finf = 0, adin = 0, adout = 0, anin = 0, anout = 0;
    int ik = 0;
    for (int b = threadIdx.x; b < nb; b += blockDim.x) {
        NormPartial p = part[b];
        pl2 += p.pl2; fl2 += p.fl2; ds += p.dstore; adin += p.adin; adout += p.adout; anin += p.anin; anout += p.anout;
        if (p.pinf > pinf || (p.pinf == pinf && p.ik > ik)) { pinf = p.pinf; ik = p.ik; }
        finf = fmax(finf, p.finf);
    }
    double t1 = block_sum<RED_BLOCK>(pl2, sh), t2 = block_sum<RED_BLOCK>(fl2, sh), t3 = block_sum<RED_BLOCK>(ds, sh);
    double t4 = block_sum<RED_BLOCK>(adin, sh), t5 = block_sum<RED_BLOCK>(adout, sh), t6 = block_sum<RED_BLOCK>(anin, sh), t7 = block_sum<RED_BLOCK>(anout, sh);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_down_sync(0xffffffffu, pinf, o);
        int oi = __shfl_down_sync(0xffffffffu, ik, o);
        double of = __shfl_down_sync(0xffffffffu, finf, o);
        if (ov > pinf || (ov == pinf && oi > ik)) { pinf = ov; ik = oi; }
        finf = fmax(finf, of);
    }
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) { shv[w] = pinf; shi[w] = ik; sh[w] = finf; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < RED_BLOCK / 32; ++q) {
            if (shv[q] > pinf || (shv[q] == pinf && shi[q] > ik)) { pinf = shv[q]; ik = shi[q]; }
            finf = fmax(finf, sh[q]);
        }
        out->pl2 = sqrt(t1); out->fl2 = sqrt(t2); out->dstore = t3; out->pinf = pinf; out->finf = finf;
        out->ikmax = ik; out->pnew_ik = pnew[ik]; out->pold_ik = pold[ik];
        out->adin = t4; out->adout = t5; out->anin = t6; out->anout = t7; out->ndin = 0.0; out->ndout = 0.0;
    }
}

// ------------------------------------------------------------------------------------------
// atmospheric boundary condition state machine per surface node
// ------------------------------------------------------------------------------------------
// SWITCH (SRC/switch.f), condensed branch for branch; sets the PONDING flag through *ponding
__global__ void k_switch(int nnod, double deltat, double pmin, double ph, const double *__restrict__ arenod,
                         const double *__restrict__ pondnod, const double *__restrict__ atmpot,
                         const double *__restrict__ qtranie, int *__restrict__ ifatm, double *__restrict__ atmact,
                         double *__restrict__ pnew, double *__restrict__ ovfl, int *__restrict__ ponding, const double *__restrict__ dtp)
{
    if (dtp) deltat = dtp[0];   // graph replay, see k_rhs_lhs
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nnod; i += gridDim.x * blockDim.x) {
        int f = ifatm[i];
        if (f == -1) { ovfl[i] = 0.0; continue; }
        double pot = atmpot[i], act = atmact[i];
        double atmdif = pot - (act - qtranie[i]);
        if (fabs(atmdif) < 1.0e-14) atmdif = 0.0;
        double pl = pondnod[i] + (atmdif * deltat / arenod[i]);
        double drain = -pondnod[i] * arenod[i] / deltat;
        if (f == 2 || f == 1) {
            bool rain = pot >= 0.0, infl = act >= 0.0;
            if (f == 1 && !rain) {
                if (infl) { if (pnew[i] <= pmin) continue; }
                else if (pnew[i] <= pmin) {
                    if (act < pot) { ifatm[i] = 0; atmact[i] = pot; pnew[i] = pmin; ovfl[i] = 0.0; }
                    continue;
                }
            }
            if (pl >= ph) { *ponding = 1; ifatm[i] = 2; pnew[i] = pl; ovfl[i] = atmdif; continue; }
            if (pl >= 0.0) { ifatm[i] = 1; ovfl[i] = atmdif; if (f == 2 && !rain) pnew[i] = 0.0; continue; }
            if (rain && !infl) { ifatm[i] = 1; ovfl[i] = atmdif; continue; }
            ifatm[i] = 0; atmact[i] = pot;
            if (f == 2 && !rain && pl > pmin) pnew[i] = 0.0;
            ovfl[i] = drain;
            continue;
        }
        if (f == 0) {
            double pn = pnew[i];
            if (pn >= ph) { *ponding = 1; ifatm[i] = 2; ovfl[i] = (pn - pondnod[i]) * arenod[i] / deltat; }
            else if (pn >= 0.0) { ifatm[i] = 1; ovfl[i] = (pn - pondnod[i]) * arenod[i] / deltat; }
            else if (pn > pmin) { ifatm[i] = 0; ovfl[i] = drain; }
            else { ifatm[i] = 1; pnew[i] = pmin; ovfl[i] = drain; }
        }
    }
}
// SWITCH_OLD (SRC/switch_old.f), subsurface-only runs
__global__ void k_switch_old(int nnod, double pmin, const double *__restrict__ atmpot, int *__restrict__ ifatm,
                             double *__restrict__ atmact, double *__restrict__ pnew)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nnod; i += gridDim.x * blockDim.x) {
        int f = ifatm[i];
        if (f == -1) continue;
        double pot = atmpot[i], act = atmact[i], pn = pnew[i];
        if (f == 1 && pn >= 0.0 && (pot < 0.0 || act > pot)) { ifatm[i] = 0; atmact[i] = pot; continue; }
        if (f == 1 && pn <= pmin && (pot > 0.0 || act < pot)) { ifatm[i] = 0; atmact[i] = pot; continue; }
        if (f == 0 && pn >= 0.0 && pot >= 0.0) { ifatm[i] = 1; pnew[i] = 0.0; continue; }
        if (f == 0 && pn <= pmin && pot < 0.0) { ifatm[i] = 1; pnew[i] = pmin; continue; }
    }
}
// ADRSTN (SRC/adrstn.f)
__global__ void k_adrstn(int nnod, double pmin, const double *__restrict__ atmpot, int *__restrict__ ifatm,
                         double *__restrict__ atmact, double *__restrict__ pnew)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nnod; i += gridDim.x * blockDim.x)
        if (pnew[i] <= pmin && (atmpot[i] > 0.0 || atmact[i] < atmpot[i])) { ifatm[i] = 0; atmact[i] = atmpot[i]; pnew[i] = pmin; }
}
// PONDUPD (SRC/pondupd.f) -- *ponding must be zeroed before the launch
__global__ void k_pondupd(int nnod, double ph, double dtr, const double *__restrict__ pondnod,
                          const double *__restrict__ arenod, const double *__restrict__ atmpot,
                          const int *__restrict__ ifatm, double *__restrict__ atmact, double *__restrict__ pnew,
                          int *__restrict__ ponding)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nnod; i += gridDim.x * blockDim.x) {
        int f = ifatm[i];
        if (f == -1) continue;
        if (pondnod[i] >= ph) {
            if (f == 1 || f == 2) { pnew[i] = pondnod[i]; *ponding = 1; }
            else if (f == 0) { *ponding = 1; atmact[i] = atmpot[i] + pondnod[i] * arenod[i] * dtr; }
        }
    }
}
// ATMNXT / ATMBAK interpolation (SRC/atmnxt.f:46-75, SRC/atmbak.f): values of two table slots
__global__ void k_atm_interp(int nnod, const double *__restrict__ tab, int stride, int rec_a, int rec_b, int use_b_only,
                             double ta, double tb, double time, int ieto, double scf, const double *__restrict__ arenod,
                             const int *__restrict__ ifatm, int set_act, double *__restrict__ atmpot,
                             double *__restrict__ atmact)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nnod; i += gridDim.x * blockDim.x) {
        // stride = 1: one value per surface node and record; stride = 0: homogeneous, one value per record
        double va = rec_a >= 0 ? (stride ? tab[(size_t)rec_a * nnod + i] : tab[rec_a]) : 0.0;
        double vb = rec_b >= 0 ? (stride ? tab[(size_t)rec_b * nnod + i] : tab[rec_b]) : 0.0;
        double pot;
        if (use_b_only) pot = vb * arenod[i];
        else {
            double slope = (vb - va) / (tb - ta);
            if (ieto != 0) slope = 0.0;
            pot = (va + slope * (time - ta)) * arenod[i];
        }
        atmpot[i] = pot;
        if (set_act && ifatm[i] == 0) atmact[i] = pot >= 0.0 ? pot : (1.0 - scf) * pot;
    }
}
// ETRAN (SRC/etran.f): Feddes root water uptake, one thread per surface column
__global__ void k_etran(int nnod, int nstr, const double *__restrict__ z, const double *__restrict__ psi,
                        const double *__restrict__ atmpot, const int *__restrict__ veg, const double *__restrict__ vegpar /* [nveg][6] */,
                        double scf, double *__restrict__ qtranie, int *__restrict__ errflag)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nnod; i += gridDim.x * blockDim.x) {
        const double *vp = vegpar + 6 * veg[i];
        double pcana = vp[0], pcref = vp[1], pcwlt = vp[2], zroot = vp[3], pz = vp[4], omgc = vp[5];
        double etp = atmpot[i] < 0.0 ? -1.0 * scf * atmpot[i] : 0.0;
        double zsurf = z[i], depth = 0.0, btran = 0.0, omg = 0.0;
        int j = 1;
        for (int l = 0; l <= nstr; ++l) qtranie[(size_t)l * nnod + i] = 0.0;
        while (depth <= zroot) {
            size_t k = (size_t)(j - 1) * nnod + i;
            if (j > nstr) { *errflag = 1; break; }
            double s1 = pcana, s2 = pcana + 1.0e-3;
            double dz = j == 1 ? (zsurf - z[k + nnod]) / 2.0 : (z[k - nnod] - z[k + nnod]) / 2.0;
            double sh = psi[k];
            double gx1 = fmin(1.0, fmax(0.0, (sh - pcwlt) / (pcref - pcwlt)));
            double gx2 = fmin(1.0, fmax(0.0, 1.0 - (sh - s1) / (s2 - s1)));
            double gx = fmin(gx1, gx2);
            double beta = (1 - depth / zroot) * exp(-1.0 * pz * depth / zroot);
            qtranie[k] = fmax(0.0, beta * dz * gx);   // BTRANI for now
            btran = btran + beta * dz;
            omg = omg + gx * beta * dz;
            ++j;
            depth = zsurf - z[(size_t)(j - 1) * nnod + i];
        }
        btran = fmax(0.0, btran);
        omg = omg / btran;
        double den = fmax(omg, omgc);
        for (int l = 0; l <= nstr; ++l) {
            size_t k = (size_t)l * nnod + i;
            qtranie[k] = etp * qtranie[k] / btran / den;
        }
    }
}
