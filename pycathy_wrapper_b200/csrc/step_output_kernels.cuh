// step_output_kernels.cuh -- end-of-step bookkeeping and output-time kernels: hydrograph terms and saturation fractions (HGRAPH, SAT_FRAC), relaxation, WEIGHT, ATMONE, MBINIT, Darcy velocities (VEL3D, VNOD3D), RECHARGE, WTDEPTH, ensemble state packing.
// Part of the single translation unit cathy_b200.cu (included in dependency order; shares its structs and helpers).
#pragma once

// end-of-step surface bookkeeping: PONDNOD=0 where PNEW<=0 (SRC/cathy_main.f:3181-3184)
__global__ void k_pond_zero(int nnod, const double *__restrict__ pnew, double *__restrict__ pondnod)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nnod; i += gridDim.x * blockDim.x)
        if (pnew[i] <= 0.0) pondnod[i] = 0.0;
}

// HGRAPH + SAT_FRAC (SRC/hgraph.f, SRC/sat_frac.f): block partials over the surface nodes ...
struct StepPartial { double apot, aact, refl, ovf; int c[13]; int pad; };
__global__ void k_step_partial(int nnod, int nstr, double pmin, double ph, const int *__restrict__ ifatm,
                               const double *__restrict__ atmpot, const double *__restrict__ atmact,
                               const double *__restrict__ pnew, StepPartial *__restrict__ part, const unsigned char *__restrict__ own)
{
    __shared__ double sh[32];
    __shared__ int shi[13];
    if (threadIdx.x < 13) shi[threadIdx.x] = 0;
    __syncthreads();
    double apot = 0, aact = 0, refl = 0, ovf = 0;
    int hg[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, nh = 0, nd = 0, np = 0, ns = 0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nnod; k += gridDim.x * blockDim.x) {
        if (own && !(own[k] & 1)) continue;
        double pot = atmpot[k], act = atmact[k], pn = pnew[k];
        int f = ifatm[k];
        apot += pot; aact += act;
        if (f == 2) { if (act < 0.0) refl = refl - act; ovf = ovf - act + pot; }
        else if (f == 1) {
            if (pn >= 0.0) {
                if (act < 0.0) {
                    if (pot >= 0.0) { refl = refl - act; ovf = ovf - act + pot; }
                    else { hg[4]++; if (act <= pot) { refl = refl - act + pot; ovf = ovf - act + pot; } }
                } else {
                    if (pot >= 0.0) { if (act <= pot) ovf = ovf - act + pot; else hg[0]++; }
                    else hg[5]++;
                }
            } else if (pn <= pmin) {
                if (act < 0.0) { if (pot >= 0.0) { ovf = ovf + pot; hg[1]++; } }
                else {
                    hg[3]++;
                    if (pot >= 0.0) { if (act <= pot) { ovf = ovf - act + pot; hg[8]++; } else hg[2]++; }
                    else hg[7]++;
                }
            }
        }
        if (pn >= 0.0) {
            ns++;
            if (pn >= ph) np++;
            int hd = 0;
            for (int l = 1; l <= nstr; ++l) if (pnew[(size_t)l * nnod + k] < 0.0) hd = 1;
            if (hd) nh++; else nd++;
        }
    }
    double t1 = block_sum<RED_BLOCK>(apot, sh), t2 = block_sum<RED_BLOCK>(aact, sh);
    double t3 = block_sum<RED_BLOCK>(refl, sh), t4 = block_sum<RED_BLOCK>(ovf, sh);
    for (int q = 0; q < 9; ++q) if (hg[q]) atomicAdd(&shi[q], hg[q]);
    if (nh) atomicAdd(&shi[9], nh);
    if (nd) atomicAdd(&shi[10], nd);
    if (np) atomicAdd(&shi[11], np);
    if (ns) atomicAdd(&shi[12], ns);
    __syncthreads();
    if (threadIdx.x == 0) {
        StepPartial p;
        p.apot = t1; p.aact = t2; p.refl = t3; p.ovf = t4; p.pad = 0;
        for (int q = 0; q < 13; ++q) p.c[q] = shi[q];
        part[blockIdx.x] = p;
    }
}
// ... and their fixed-order reduction together with STORE1 (SRC/storcal.f), one block
__global__ void k_step_final(int nbs, const StepPartial *__restrict__ spart, int nbpart, const double *__restrict__ store_part,
                             StepOut *__restrict__ out)
{
    __shared__ double sh[32];
    __shared__ int shi[13];
    if (threadIdx.x < 13) shi[threadIdx.x] = 0;
    __syncthreads();
    double st = 0.0, apot = 0, aact = 0, refl = 0, ovf = 0;
    int c[13] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int b = threadIdx.x; b < nbpart; b += blockDim.x) st += store_part[b];
    for (int b = threadIdx.x; b < nbs; b += blockDim.x) {
        StepPartial p = spart[b];
        apot += p.apot; aact += p.aact; refl += p.refl; ovf += p.ovf;
        for (int q = 0; q < 13; ++q) c[q] += p.c[q];
    }
    double t0 = block_sum<RED_BLOCK>(st, sh), t1 = block_sum<RED_BLOCK>(apot, sh), t2 = block_sum<RED_BLOCK>(aact, sh);
    double t3 = block_sum<RED_BLOCK>(refl, sh), t4 = block_sum<RED_BLOCK>(ovf, sh);
    for (int q = 0; q < 13; ++q) if (c[q]) atomicAdd(&shi[q], c[q]);
    __syncthreads();
    if (threadIdx.x == 0) {
        out->store1 = t0; out->apot = t1; out->aact = t2; out->reflow = t3; out->ovflow = t4;
        for (int q = 0; q < 9; ++q) out->hgflag[q] = shi[q];
        out->nhort = shi[9]; out->ndunn = shi[10]; out->npond = shi[11]; out->nsat = shi[12];
    }
}
// RELAX with a constant factor (SRC/relax.f, NLRELX = 1): PNEW = (1 - OMEGA) POLD + OMEGA PNEW, after the mass balance and
// before the convergence norms (SRC/flow3d.f:165-190)
// RELXOM (SRC/relxom.f:20-39, NLRELX = 2): the signed head change of largest magnitude (ties -> the LAST node, the sequential >= test),
// block partials in fixed order, then OMEGA from Huyakorn's adaptation of Cooley's scheme with the previous iteration's signed maximum
// PIKMXV(ITER-1) = PNEW(IKMAX) - POLD(IKMAX) of NORMS, still in the IterOut record on the device
struct RelxPartial { double amax, diff; int ik, pad; };
__global__ void k_relxom_partial(int n, const double *__restrict__ pnew, const double *__restrict__ pold, RelxPartial *__restrict__ part)
{
    __shared__ double sha[RED_BLOCK / 32], shd[RED_BLOCK / 32];
    __shared__ int shi[RED_BLOCK / 32];
    double am = 0.0, df = 0.0;
    int ik = -1;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const double d = pnew[k] - pold[k], da = fabs(d);
        if (da > am || (da == am && k >= ik)) { am = da; df = d; ik = k; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double oa = __shfl_down_sync(0xffffffffu, am, o), od = __shfl_down_sync(0xffffffffu, df, o);
        const int oi = __shfl_down_sync(0xffffffffu, ik, o);
        if (oa > am || (oa == am && oi > ik)) { am = oa; df = od; ik = oi; }
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { sha[w] = am; shd[w] = df; shi[w] = ik; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < RED_BLOCK / 32; ++q)
            if (sha[q] > am || (sha[q] == am && shi[q] > ik)) { am = sha[q]; df = shd[q]; ik = shi[q]; }
        RelxPartial p; p.amax = am; p.diff = df; p.ik = ik; p.pad = 0;
        part[blockIdx.x] = p;
    }
}
__global__ void k_relxom_final(int nb, const RelxPartial *__restrict__ part, int iter, const IterOut *__restrict__ prev, double *__restrict__ om)
{   // om[0] = OMEGA, om[1] = OMEGAP; one thread
    double omega = 1.0;
    if (iter > 1) {
        double am = 0.0, difmx = 0.0;
        int ik = -1;
        for (int q = 0; q < nb; ++q)
            if (part[q].amax > am || (part[q].amax == am && part[q].ik > ik)) { am = part[q].amax; difmx = part[q].diff; ik = part[q].ik; }
        const double difmxp = prev->pnew_ik - prev->pold_ik, zeta = difmx / (om[1] * difmxp);
        omega = zeta >= -1.0 ? (3.0 + zeta) / (3.0 + fabs(zeta)) : 0.5 / fabs(zeta);
    }
    om[0] = omega; om[1] = omega;
}
__global__ void k_relax(int n, double omega, const double *__restrict__ pold, double *__restrict__ pnew, const double *__restrict__ omega_dev)
{
    if (omega_dev) omega = *omega_dev;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) pnew[k] = (1.0 - omega) * pold[k] + omega * pnew[k];
}
__global__ void k_weight(int n, double tetaf, const double *__restrict__ pnew, const double *__restrict__ ptimep, double *__restrict__ ptnew)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) ptnew[k] = tetaf * pnew[k] + (1.0 - tetaf) * ptimep[k];
}
// ATMONE's classification of surface nodes (SRC/atmone.f label 500 onwards)
__global__ void k_atmone(int nnod, double pmin, double ph, double scf, const double *__restrict__ atmpot, double *__restrict__ atmold,
                         double *__restrict__ atmact, double *__restrict__ pnew, double *__restrict__ ptimep, int *__restrict__ ifatm,
                         int *__restrict__ ifatmp)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nnod; i += gridDim.x * blockDim.x) {
        int f = ifatm[i], fp = ifatmp[i];
        if (f != -1) {
            if (pnew[i] >= ph) { f = 2; fp = 2; }
            else {
                if (pnew[i] >= 0.0 && atmpot[i] > 0.0) f = 1;
                if (ptimep[i] >= 0.0 && atmold[i] > 0.0) fp = 1;
                if (pnew[i] <= pmin && atmpot[i] < 0.0) { pnew[i] = pmin; f = 1; }
                if (ptimep[i] <= pmin && atmold[i] < 0.0) { ptimep[i] = pmin; fp = 1; }
            }
        }
        ifatm[i] = f; ifatmp[i] = fp;
        if (f == 0) atmact[i] = atmpot[i] >= 0.0 ? atmpot[i] : (1.0 - scf) * atmpot[i];
        else atmact[i] = 0.0;
        if (fp == 1 || fp == 2) atmold[i] = 0.0;
    }
}
__global__ void k_mbinit(int nnod, const int *__restrict__ ifatmp, const double *__restrict__ atmold, double *__restrict__ out3,
                         const unsigned char *__restrict__ own)
{   // MBINIT sums (SRC/mbinit.f): AACTP, ANINP, ANOUTP -- one block
    __shared__ double sh[32];
    double a = 0, b = 0, c = 0;
    for (int k = threadIdx.x; k < nnod; k += blockDim.x)
        if (ifatmp[k] == 0 && (!own || (own[k] & 1))) { a += atmold[k]; if (atmold[k] > 0.0) b += atmold[k]; else c += atmold[k]; }
    double t0 = block_sum<RED_BLOCK>(a, sh), t1 = block_sum<RED_BLOCK>(b, sh), t2 = block_sum<RED_BLOCK>(c, sh);
    if (threadIdx.x == 0) { out3[0] = t0; out3[1] = t1; out3[2] = t2; }
}


// VEL3D (SRC/vel3d.f): Darcy velocity per element from the nodal heads, basis-function coefficients recomputed from the node
// coordinates (SRC/basis6.f / volbas.f formulas, same operation order as build_static) instead of being stored per element
__global__ void k_vel3d(int nt, int ntri, int nzone, const int4 *__restrict__ tet, const int *__restrict__ trizone,
                        const double *__restrict__ permx, const double *__restrict__ permy, const double *__restrict__ permz,
                        const double *__restrict__ X, const double *__restrict__ Y, const double *__restrict__ Z,
                        const double *__restrict__ psi, const double *__restrict__ ckrw, double *__restrict__ uu, double *__restrict__ vv,
                        double *__restrict__ ww)
{
    const double amen[5] = {-1.0, 1.0, -1.0, 1.0, -1.0};
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nt; e += gridDim.x * blockDim.x) {
        int4 t4 = tet[e];
        const int T[4] = {t4.x, t4.y, t4.z, t4.w};
        double x[4], y[4], z[4], p[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { x[q] = X[T[q]]; y[q] = Y[T[q]]; z[q] = Z[T[q]]; p[q] = psi[T[q]]; }
        double vol = 0.0, bb = 0.0, cc = 0.0, dd = 0.0;
#pragma unroll
        for (int nn = 0; nn < 4; ++nn) {
            const int o3[3] = {(nn + 1) & 3, (nn + 2) & 3, (nn + 3) & 3};
            double a2 = 0.0, a3 = 0.0;
#pragma unroll
            for (int ii = 0; ii < 3; ++ii) { int I = o3[ii], J = o3[(ii + 1) % 3], M = o3[(ii + 2) % 3]; a3 = y[I] * z[J] + a3; a2 = y[I] * z[M] + a2; }
            double b = amen[nn] * (a3 - a2) / 6.0;
            vol = vol + x[nn] * amen[nn] * (a3 - a2) / 6.0;
            a2 = a3 = 0.0;
#pragma unroll
            for (int ii = 0; ii < 3; ++ii) { int I = o3[ii], J = o3[(ii + 1) % 3], M = o3[(ii + 2) % 3]; a3 = x[I] * z[J] + a3; a2 = x[I] * z[M] + a2; }
            double c = amen[nn + 1] * (a3 - a2) / 6.0;
            a2 = a3 = 0.0;
#pragma unroll
            for (int ii = 0; ii < 3; ++ii) { int I = o3[ii], J = o3[(ii + 1) % 3], M = o3[(ii + 2) % 3]; a3 = x[I] * y[J] + a3; a2 = x[I] * y[M] + a2; }
            double d = amen[nn] * (a3 - a2) / 6.0;
            bb = bb + p[nn] * b; cc = cc + p[nn] * c; dd = dd + p[nn] * d;
        }
        const int ivol = vol < 0.0 ? -1 : 1;
        const double volur = 1.0 / fabs(vol);
        const double kre = (((ckrw[T[0]] + ckrw[T[1]]) + ckrw[T[2]]) + ckrw[T[3]]) * 0.25;
        const int lay = e / (3 * ntri), tri = (e - lay * 3 * ntri) / 3, idx = lay * nzone + trizone[tri];
        const double xyz = -kre * volur * ivol;
        uu[e] = bb * xyz * permx[idx];
        vv[e] = cc * xyz * permy[idx];
        ww[e] = (dd * xyz - kre) * permz[idx];
    }
}
// VNOD3D (SRC/vnod3d.f): nodal velocity = mean over the elements of the node, summed in element order (the node family of the
// assembly plan lists them in that order; padding entries carry coef2 = 0)
__global__ void k_vnod3d(int n, EllPlan P, const double *__restrict__ uu, const double *__restrict__ vv, const double *__restrict__ ww,
                         double *__restrict__ unod, double *__restrict__ vnod, double *__restrict__ wnod)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const EllFamily f = P.node;
        double a = 0.0, b = 0.0, c = 0.0;
        int cnt = 0;
        for (int q = 0; q < f.w; ++q) {
            size_t i = (size_t)q * P.ld + k;
            if (f.coef2[i] != 0.0) { int t = f.tet[i]; a = a + uu[t]; b = b + vv[t]; c = c + ww[t]; ++cnt; }
        }
        unod[k] = a / cnt; vnod[k] = b / cnt; wnod[k] = c / cnt;
    }
}


// RECHARGE (SRC/recharge.f): per surface column, the vertical nodal velocity at the node just above the water table
__global__ void k_recharge(int nnod, int nstr, const double *__restrict__ psi, const double *__restrict__ wnod, const double *__restrict__ arenod,
                           double *__restrict__ recnod)
{
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nnod; s += gridDim.x * blockDim.x) {
        const size_t i = (size_t)nnod * nstr + s;
        double r = 0.0;
        bool done = false;
        for (int j = 1; j <= nstr && !done; ++j)
            if (psi[i - (size_t)(j - 1) * nnod] > 0.0 && psi[i - (size_t)j * nnod] <= 0.0 && wnod[i - (size_t)j * nnod] <= 0.0) {
                r = -1.0 * wnod[i - (size_t)j * nnod] * arenod[s];
                done = true;
            }
        if (!done && psi[s] >= 0.0 && wnod[s] <= 0.0) r = -1.0 * wnod[s] * arenod[s];
        recnod[s] = r;
    }
}
// WTDEPTH (SRC/wtdepth.f), one thread per requested surface node
__global__ void k_wtdepth(int numvp, const int *__restrict__ nodvp, int nnod, int nstr, const double *__restrict__ Z, const double *__restrict__ P,
                          double *__restrict__ wt)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numvp) return;
    const int nd = nodvp[i] - 1;
    int flag = 0;
    double v = Z[nd];
    for (int j = nstr; j >= 1; --j) {
        const size_t i1 = nd + (size_t)j * nnod, i2 = nd + (size_t)(j - 1) * nnod;
        if (P[i1] >= 0.0 && P[i2] < 0.0 && flag == 0) { double rc = (Z[i1] - Z[i2]) / (P[i1] - P[i2]); v = Z[i1] - rc * P[i1]; flag = 1; }
        else if (P[i1] >= 0.0 && P[i2] < 0.0 && flag == 1) flag = 2;
        else if (j == 1 && P[i2] >= 0.0 && flag == 0) { flag = 3; v = Z[nd] + P[i2]; }
        else if (j == 1 && flag == 0) { flag = 4; v = Z[nd + (size_t)nstr * nnod]; }
    }
    wt[i] = v;
}

// one member's state <-> column `col` of a row-major ensemble matrix [n][ld]
__global__ void k_pack_col(int n, const double *__restrict__ v, double *__restrict__ X, long long ld, long long col)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) X[(long long)k * ld + col] = v[k];
}
__global__ void k_unpack_col(int n, const double *__restrict__ X, long long ld, long long col, double *__restrict__ a, double *__restrict__ b)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) { double v = X[(long long)k * ld + col]; a[k] = v; b[k] = v; }
}
