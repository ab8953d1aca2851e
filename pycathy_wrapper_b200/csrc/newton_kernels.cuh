// newton_kernels.cuh -- Newton scheme (IOPT = 2): moisture-curve derivatives, element factors, Jacobian assembly, RHS, back-calculated fluxes, the Jacobi-preconditioned persistent BiCGSTAB k_bicgstab.
// Part of the single translation unit cathy_b200.cu (included in dependency order; shares its structs and helpers).
#pragma once

// ==========================================================================================
// Newton scheme (IOPT = 2): SRC/newton.f.  The Jacobian J = TETAF*A + M/dt + C3 is nonsymmetric with the same
// 15-point stencil: upper part (incl. diagonal) in 8 diagonals Ju[d][k] = J(k, k+off_d), lower part in 7 diagonals
// Jl[d][k] = J(k+off_d, k) -- same coalesced, index-free layout as the Picard matrix.
// ==========================================================================================
// SRC/fvgdkr.f, SRC/fvgdds.f
__device__ __forceinline__ double fvgdkr(double psi, double psat, double n, double m, double n1, double m52, double mm1)
{
    if (psi < -1.0e-14) {
        double beta = pow(fabs(psi / psat), n);
        double b1 = beta + 1.0;
        double v1 = pow(fabs(b1), m) - pow(fabs(beta), m);
        double v2 = psat / psi;
        double v3 = n1 * beta * v2 * pow(fabs(1.0 / b1), m52) / psat;
        double v4 = v2 * ((2.5 / b1) * beta - 2.0) - 0.5 * pow(fabs(b1), mm1);
        return v3 * v1 * v4;
    }
    return 0.0;
}
__device__ __forceinline__ double fvgdds(double psi, double psat, double n, double m, double n1)
{
    if (psi < -1.0e-14) {
        double beta = pow(fabs(psi / psat), n);
        double b1 = beta + 1.0, b1r = 1.0 / b1;
        return n1 * (beta / psi) * (1.0 / psi) * ((1.0 + n * (beta - 1.0)) / pow(fabs(b1), m)) * b1r * b1r;
    }
    return 0.0;
}
// NEWUNS -> CHNEW0 (SRC/newuns.f, SRC/chnew0.f, IVGHU = 0)
__global__ void k_curves_newton(int n, Soil s, const double *__restrict__ ptnew, double *__restrict__ sw, double *__restrict__ ckrw,
                                double *__restrict__ etai, double *__restrict__ dckrw, double *__restrict__ detai)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double n_ = s.vgn[i], m = s.vgm[i], psat = s.vgpsat[i], pnot = s.vgpnot[i], n1 = s.vgn1[i];
        double psi = ptnew[i];
        double se = fvgse(psi, psat, n_, m);
        double dswdp = pnot * fvgdse(psi, psat, n_, n1, s.vgnr[i], s.vgpsn[i]);
        double w = pnot * se + s.rr[i];
        sw[i] = w;
        etai[i] = w * s.snodi[i] + s.pnodi[i] * dswdp;
        detai[i] = dswdp * s.snodi[i] + s.pnodi[i] * pnot * fvgdds(psi, psat, n_, m, n1);
        ckrw[i] = fvgkr(psi, se, m, s.vgmr[i]);
        dckrw[i] = fvgdkr(psi, psat, n_, m, n1, s.vgm52[i], s.vgmm1[i]);
    }
}
// CHNEW0 for IVGHU = 1..4 (SRC/chnew0.f:39-89): the curve of the node plus the derivatives the Jacobian needs, d(kr)/d(psi) and
// d(eta)/d(psi) -- SRC/fxvddm.f, fxvdkr.f (extended van Genuchten), fhudds.f, fhudk2.f, fhudk3.f (Huyakorn), fbcdds.f, fbcdkr.f (Brooks-Corey)
__global__ void k_curves_newton_alt(int n, CurveModel c, Soil s, const double *__restrict__ ptnew, double *__restrict__ sw, double *__restrict__ ckrw,
                                    double *__restrict__ etai, double *__restrict__ dckrw, double *__restrict__ detai)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double psi = ptnew[i], sn = s.snodi[i], po = s.pnodi[i];
        double w, kr, eta, dkr = 0.0, deta = 0.0;
        if (c.ivghu == 1) {
            xvg_node(s, i, psi, true, true, w, kr, eta);
            const double nn = s.vgn[i], m = s.vgm[i], psat = s.vgpsat[i], n1 = s.vgn1[i];
            if (psi < -1.0e-14) {
                const double beta = pow(fabs(psi / psat), nn), b1 = beta + 1.0, b1r = 1.0 / b1;
                const double v1 = pow(b1, m) - pow(beta, m), v2 = psat / psi;
                const double v3 = n1 * beta * v2 * pow(b1r, s.vgm52[i]) / psat;
                const double v4 = v2 * ((2.5 / b1) * beta - 2.0) - 0.5 * pow(b1, s.vgmm1[i]);
                dkr = v3 * v1 * v4;
                if (psi < s.vgpnot[i]) deta = n1 * (po - s.rr[i]) * (beta / psi) * (1.0 / psi) * ((1.0 + nn * (beta - 1.0)) / pow(b1, m)) * b1r * b1r;
            }
        } else {
            double dsw;
            curve_alt(c, psi, po, w, kr, dsw, true);
            eta = w * sn + po * dsw;
            double d2 = 0.0;     // d2(sw)/d(psi)2
            if (c.ivghu == 4) {
                if (psi < c.bcpsat) {
                    const double porm = (po - c.bcrmc) / po, q = c.bcpsat / psi;
                    d2 = porm * ((c.bcbeta * c.bcb1 / (c.bcpsat * c.bcpsat)) * pow(q, c.bcbeta + 2.0));
                    dkr = (c.bc23b / fabs(c.bcpsat)) * pow(q, 3.0 + (3.0 * c.bcbeta));
                }
            } else if (psi < c.hupsia) {
                const double pap = c.hupsia - psi, papr = 1.0 / pap, lambda = c.hualb * pow(pap, c.hubeta), lamr = 1.0 / (1.0 + lambda);
                const double se = pow(lamr, c.hugama), dsedp = (c.hugb * lambda / pap) * pow(lamr, c.hugam1);
                d2 = c.huswr1 * (c.hugb * lambda * papr * papr * ((1.0 - c.hubeta) + (1.0 + c.hugb) * lambda) * pow(lamr, c.hugama + 2.0));
                dkr = c.ivghu == 2 ? c.hun * pow(se, c.hun - 1.0) * dsedp : ((2.0 * c.hua) * se + c.hub2a) * dsedp * kr * log(10.0);
            }
            deta = dsw * sn + po * d2;
        }
        sw[i] = w; ckrw[i] = kr; etai[i] = eta; dckrw[i] = dkr; detai[i] = deta;
    }
}
__global__ void k_sw_pair_alt(int n, CurveModel c, Soil s, const double *__restrict__ pnew, const double *__restrict__ ptimep, int do_timep,
                              double *__restrict__ swnew, double *__restrict__ swtimep)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double w, d1, d2;
        if (c.ivghu == 1) xvg_node(s, i, pnew[i], false, false, w, d1, d2); else curve_alt(c, pnew[i], s.pnodi[i], w, d1, d2, false);
        swnew[i] = w;
        if (do_timep) {
            if (c.ivghu == 1) xvg_node(s, i, ptimep[i], false, false, w, d1, d2); else curve_alt(c, ptimep[i], s.pnodi[i], w, d1, d2, false);
            swtimep[i] = w;
        }
    }
}
// SWNEW = Sw(PNEW), SWTIMEP = Sw(PTIMEP) for the storage change of the mass balance (the reference's Newton path leaves
// them unset -- its mbeconv prints NaN there)
__global__ void k_sw_pair(int n, Soil s, const double *__restrict__ pnew, const double *__restrict__ ptimep, int do_timep,
                          double *__restrict__ swnew, double *__restrict__ swtimep)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double n_ = s.vgn[i], m = s.vgm[i], psat = s.vgpsat[i], pnot = s.vgpnot[i], rr = s.rr[i];
        swnew[i] = pnot * fvgse(pnew[i], psat, n_, m) + rr;
        if (do_timep) swtimep[i] = pnot * fvgse(ptimep[i], psat, n_, m) + rr;
    }
}
// Element pass of ASSNEW (SRC/assnew.f:29-66): element means of kr and eta, and per local node k the two factors of the
// derivative terms, TSUMTD = TETAF*(K0_e psi)_k + TETAF*Kz*IVOL*d_k and SUM1TV = LMASS(k,k)*(psi_k - psi0_k)*TETAF*V/dt (LUMP = 1).
__global__ void k_tet_newton(int nt, const int4 *__restrict__ tet, const double *__restrict__ ckrw, const double *__restrict__ etai,
                             const double *__restrict__ ptnew, const double *__restrict__ pnew, const double *__restrict__ ptimep,
                             const double *__restrict__ k0, const double *__restrict__ gz, const double *__restrict__ vol, double tetaf,
                             double rdt, double *__restrict__ krt, double *__restrict__ etat, double *__restrict__ ts, double *__restrict__ s1)
{
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nt; e += gridDim.x * blockDim.x) {
        int4 t = tet[e];
        const int nd[4] = {t.x, t.y, t.z, t.w};
        krt[e] = (((ckrw[t.x] + ckrw[t.y]) + ckrw[t.z]) + ckrw[t.w]) * 0.25;
        etat[e] = (((etai[t.x] + etai[t.y]) + etai[t.z]) + etai[t.w]) * 0.25;
        double K[4][4], psi[4];
        int pr = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int l = k; l < 4; ++l, ++pr) { double v = k0[(size_t)pr * nt + e]; K[k][l] = v; K[l][k] = v; }
#pragma unroll
        for (int k = 0; k < 4; ++k) psi[k] = ptnew[nd[k]];
        const double tvd = tetaf * vol[e] * rdt;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double sum = 0.0;
#pragma unroll
            for (int m = 0; m < 4; ++m) sum = sum + K[k][m] * psi[m];
            ts[(size_t)k * nt + e] = tetaf * sum + tetaf * gz[(size_t)k * nt + e];
            s1[(size_t)k * nt + e] = (0.25 * (pnew[nd[k]] - ptimep[nd[k]])) * tvd;
        }
    }
}
// Gather pass of ASSNEW: stiffness A (symmetric, 8 upper diagonals) and the derivative part C3 of the Jacobian, upper and lower.
// DERIVED: tet indices as base(k) + per-class offset (tables of k_assemble_a, verified against every stored entry at cathy_create)
// instead of the stored lists: 4 bytes less per contribution, same contributions in the same order.
template <bool DERIVED>
__global__ void __launch_bounds__(RED_BLOCK) k_assemble_newton(int n, int nt, EllPlan P, PlanGeom G, const unsigned char *__restrict__ loc,
                                                               const double *__restrict__ krt, const double *__restrict__ etat,
                                                               const double *__restrict__ ts, const double *__restrict__ s1,
                                                               const double *__restrict__ dckrw, const double *__restrict__ detai, Diag A,
                                                               Diag C3u, Diag C3l, double *__restrict__ grav, double *__restrict__ m2)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        int base = 0;
        const int *__restrict__ rl = nullptr;
        if (DERIVED) {
            const int l = k / G.nnod, sidx = k - l * G.nnod, i = sidx / G.nc1, j = sidx - i * G.nc1;
            const int cls = ((l == 0 ? 0 : l == G.nstr ? 2 : 1) * 3 + (i == 0 ? 0 : i == G.nrow ? 2 : 1)) * 3 + (j == 0 ? 0 : j == G.ncol ? 2 : 1);
            base = G.ntri3 * l + 6 * (i * G.ncol + j);
            rl = G.rel + (size_t)cls * NDIAG * G.wrel;
        }
#pragma unroll
        for (int d = 0; d < NDIAG; ++d) {
            const EllFamily f = P.diag[d];
            const unsigned char *lc = loc + (f.tet - P.diag[0].tet);
            double acc = 0.0, gu = 0.0, hu = 0.0, gl = 0.0, hl = 0.0;
            for (int c = 0; c < f.w; ++c) {
                size_t q = (size_t)c * P.ld + k;
                int t = DERIVED ? min(max(base + __ldg(rl + d * G.wrel + c), 0), G.nt - 1) : f.tet[q];
                unsigned l = lc[q];
                acc += krt[t] * f.coef[q];
                if (l & 16u) {
                    size_t ia = (size_t)(l & 3u) * nt + t, ib = (size_t)((l >> 2) & 3u) * nt + t;
                    gu += ts[ia]; hu += s1[ia];
                    gl += ts[ib]; hl += s1[ib];
                }
            }
            const int col = k + A.off[d];
            A.d[d][k] = acc;
            C3u.d[d][k] = dckrw[col] * gu + detai[col] * hu;     // J(k, k+off): derivative w.r.t. the COLUMN node's head
            if (d > 0) C3l.d[d][k] = dckrw[k] * gl + detai[k] * hl;   // J(k+off, k)
        }
        const EllFamily f = P.node;
        double g = 0.0, m = 0.0;
        for (int c = 0; c < f.w; ++c) {
            size_t q = (size_t)c * P.ld + k;
            int t = DERIVED ? min(max(base + __ldg(rl + c), 0), G.nt - 1) : f.tet[q];     // DERIVED implies node.pad: the node family lists the tets of diag[0]
            g += krt[t] * f.coef[q];
            m += etat[t] * f.coef2[q];
        }
        grav[k] = g;
        m2[k] = m;
    }
}
// RHSNEW + CFMATN + RHSGRV + BCNEW (SRC/rhsnew.f, cfmatn.f, rhsgrv.f, bcnew.f): RHS, Jacobian in place of C3, Dirichlet mask
__global__ void k_rhs_lhs_newton(int n, int nnod, Diag A, Diag Ju, Diag Jl, double tetaf, double rdt, const double *__restrict__ ptnew,
                                 const double *__restrict__ pnew, const double *__restrict__ ptimep, const double *__restrict__ m2,
                                 const double *__restrict__ grav, const int *__restrict__ ifatm, const unsigned char *__restrict__ contp_flag,
                                 const double *__restrict__ qneu, const double *__restrict__ atmact, const double *__restrict__ atmold,
                                 double *__restrict__ rhs, double *__restrict__ xt5, double *__restrict__ diag_true, double *__restrict__ dinv)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        double ax = dia_row(A, A.d[0], ptnew, k, n);
        double b = -ax - m2[k] * rdt * (pnew[k] - ptimep[k]) - grav[k];
        xt5[k] = b;
        double dg = tetaf * A.d[0][k] + m2[k] * rdt + Ju.d[0][k];
        Ju.d[0][k] = dg;
#pragma unroll
        for (int d = 1; d < NDIAG; ++d) {
            double a = tetaf * A.d[d][k];
            Ju.d[d][k] = a + Ju.d[d][k];
            Jl.d[d][k] = a + Jl.d[d][k];
        }
        diag_true[k] = dg;
        bool dir = is_dirichlet(k, nnod, ifatm, contp_flag);
        if (dir) b = 0.0;
        if (qneu) b += qneu[k];
        if (k < nnod && ifatm[k] == 0) b = b + (tetaf * atmact[k] + (1.0 - tetaf) * atmold[k]);
        rhs[k] = b;
        dinv[k] = dir ? 0.0 : 1.0 / dg;      // Dirichlet rows: increment pinned to 0 (the reference's 1.7e91 penalty gives |x| ~ 1e-91)
    }
}
// nonsymmetric DIA row product
__device__ __forceinline__ double dia_row_n(const Diag &U, const Diag &L, const double *x, int k)
{
    double acc = U.d[0][k] * x[k];
#pragma unroll
    for (int d = 1; d < NDIAG; ++d) acc += U.d[d][k] * x[k + U.off[d]];
#pragma unroll
    for (int d = 1; d < NDIAG; ++d) acc += L.d[d][k - U.off[d]] * x[k - U.off[d]];
    return acc;
}
// BKNEW (SRC/bknew.f) at atmospheric Dirichlet nodes / prescribed-head nodes
__global__ void k_bkflux_n(int nnod, Diag U, Diag L, const double *__restrict__ pdiff, const double *__restrict__ xt5,
                           const int *__restrict__ ifatm, double tetaf, const double *__restrict__ atmold, double *__restrict__ atmact)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nnod; k += gridDim.x * blockDim.x) {
        int f = ifatm[k];
        if (f == 1 || f == 2) {
            double scr = dia_row_n(U, L, pdiff, k) - xt5[k];
            atmact[k] = (scr - (1.0 - tetaf) * atmold[k]) * (1.0 / tetaf);
        }
    }
}
__global__ void k_bkflux_list_n(int m, const int *__restrict__ list, Diag U, Diag L, const double *__restrict__ pdiff,
                                const double *__restrict__ xt5, double tetaf, const double *__restrict__ qpold, double *__restrict__ qpnew)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        int k = list[i];
        double scr = dia_row_n(U, L, pdiff, k) - xt5[k];
        qpnew[i] = (scr - (1.0 - tetaf) * qpold[i]) * (1.0 / tetaf);
    }
}

// NSYSLV (SRC/solscal-extended.f:3063-3240) as ONE persistent cooperative kernel: right-preconditioned BiCGSTAB
// (the recurrence of GCSTAS, :1010-1128, with M = diag(J) in place of the sequential ILU(0) factors), four grid barriers
// per iteration.  Dirichlet rows carry dinv = 0: every Krylov vector stays exactly zero there, which is the limit of the
// reference's penalty rows.  Stopping test as in GCSTAS: ||r||_2 / ||b_free||_2 <= TOLCG.
struct BicgArgs {
    int n, itmax;
    double tol;
    Diag U, L;
    const double *dinv, *rhs;
    double *x, *r, *rt, *p, *ph, *v, *s, *sh, *t;
    double *partial;         // [2][5][gridDim.x]
    unsigned int *counter;
    unsigned int epoch0;
    IterOut *out;
    int prefetch;
    int zigzag;              // 1: boustrophedon sweeps (Jacobian larger than the L2), see k_bicgstab
    int line;                // 1: vertical-line (one tridiagonal system per DEM column) preconditioner, 0: point Jacobi
    int nnod, nl;            // surface nodes (= columns) and node layers (rows of a column: s, s + nnod, ...)
    double *idn, *cp;        // Thomas factors of the column systems: 1 / pivot and the eliminated super-diagonal
};
template <int BLOCK, int NS>
__device__ __forceinline__ void grid_reduce_n(unsigned int *counter, unsigned int &epoch, unsigned int &flip, const double (&in)[NS],
                                              double *partial_base, double (*sh)[NS], double (&out)[NS])
{
    const int nb = gridDim.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    // two partial buffers used alternately: a block may start the next reduction while a slower one is still reading the
    // partials of this one
    double *partial = partial_base + (size_t)(flip & 1u) * NS * nb;
    ++flip;
    double v[NS];
#pragma unroll
    for (int q = 0; q < NS; ++q) v[q] = warp_sum(in[q]);
    if (lane == 0)
#pragma unroll
        for (int q = 0; q < NS; ++q) sh[w][q] = v[q];
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int q = 0; q < NS; ++q) {
            double t = lane < BLOCK / 32 ? sh[lane][q] : 0.0;
            t = warp_sum(t);
            if (lane == 0) partial[q * nb + blockIdx.x] = t;
        }
    }
    grid_barrier(counter, epoch);
    if (w < NS) {
        double s0 = 0.0;
        const volatile double *pp = partial + w * nb;
        for (int i = lane; i < nb; i += 32) s0 += pp[i];
        double t = warp_sum(s0);
        if (lane == 0) sh[0][w] = t;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NS; ++q) out[q] = sh[0][q];
    __syncthreads();
}
// the 15 matrix streams of the next grid-stride row, pulled into L2 while this row's FMA chain runs (as in k_pcg: the Jacobian
// of a large mesh streams from HBM twice per iteration)
__device__ __forceinline__ void bicg_prefetch_row(const Diag &U, const Diag &L, int kn)
{
#pragma unroll
    for (int d = 0; d < NDIAG; ++d) l2_prefetch(&U.d[d][kn]);
#pragma unroll
    for (int d = 1; d < NDIAG; ++d) l2_prefetch(&L.d[d][kn - U.off[d]]);
}
// Vertical-line preconditioner.  The layers of the DEM mesh are thin against the cell size (config 3: 0.15 m against 0.5 m), so on
// saturated (elliptic) systems the coupling between the nodes of one DEM column dominates: M = the block diagonal of J with one
// nonsymmetric tridiagonal block per column (sub-/super-diagonal = the +-NNOD diagonals).  Opt-in (CATHY_BICG_LINE=1): measured on
// B200 at config 3 it saves 36 % of the BiCGSTAB iterations of the saturated storm (103 -> 66 per solve) but each iteration costs
// 46 % more (two latency-bound column sweeps of 21 us), and it saves nothing on unsaturated systems.
// One thread per column: Thomas factorisation once per solve, two dependent sweeps over the nl layers per application; adjacent
// threads own adjacent columns, so every access is coalesced.  Dirichlet rows (dinv = 0) are identity rows with a zero right-hand
// side: their factor entries are 0, which also removes them from the neighbouring rows' recurrences.
__device__ __forceinline__ void line_factor(const BicgArgs &a, int t0, int stride)
{
    const double *lo = a.L.d[NDIAG - 1], *up = a.U.d[NDIAG - 1], *dg = a.U.d[0];
    for (int sidx = t0; sidx < a.nnod; sidx += stride) {
        double cprev = 0.0;
        for (int l = 0, k = sidx; l < a.nl; ++l, k += a.nnod) {
            double idn = 0.0, c = 0.0;
            if (a.dinv[k] != 0.0) {
                const double piv = dg[k] - (l ? lo[k - a.nnod] * cprev : 0.0);
                idn = 1.0 / piv;
                c = l + 1 < a.nl ? up[k] * idn : 0.0;
            }
            a.idn[k] = idn; a.cp[k] = c;
            cprev = c;
        }
    }
}
__device__ __forceinline__ void line_solve(const BicgArgs &a, const double *in, double *out, int t0, int stride)
{
    const double *lo = a.L.d[NDIAG - 1];
    const int nnod = a.nnod, nl = a.nl;
    for (int sidx = t0; sidx < nnod; sidx += stride) {
        double y = in[sidx] * a.idn[sidx];
        out[sidx] = y;
#pragma unroll 4
        for (int l = 1; l < nl; ++l) {
            const int k = sidx + l * nnod;
            y = (in[k] - lo[k - nnod] * y) * a.idn[k];
            out[k] = y;
        }
#pragma unroll 4
        for (int l = nl - 2; l >= 0; --l) {
            const int k = sidx + l * nnod;
            y = out[k] - a.cp[k] * y;
            out[k] = y;
        }
    }
}
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, 1) k_bicgstab(BicgArgs a)
{
    const bool PF = a.prefetch != 0;
    const bool LINE = a.line != 0;
    const bool ZZ = a.zigzag != 0;
    __shared__ double sh[BLOCK / 32][5];
    unsigned int epoch = a.epoch0, flip = 0;
    const int n = a.n, stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int klast = t0 < n ? t0 + ((n - 1 - t0) / stride) * stride : -1;     // this thread's last row
    const double *__restrict__ di = a.dinv;
    double in[5] = {0, 0, 0, 0, 0}, out[5];
    // x0 = M^-1 b, xlung = ||b_free||^2
    for (int k = t0; k < n; k += stride) {
        double b = a.rhs[k], d = di[k];
        a.x[k] = b * d;
        if (d != 0.0) in[0] += b * b;
    }
    if (LINE) line_factor(a, t0, stride);
    grid_reduce_n<BLOCK, 5>(a.counter, epoch, flip, in, a.partial, sh, out);
    const double xlung = out[0];
    // r0 = b - J x0 (zero on Dirichlet rows), rt = r0, p = r0, ph = M^-1 p; rho = (rt, r0)
    in[0] = 0.0;
    for (int k = t0; k < n; k += stride) {
        double d = di[k];
        double r = d != 0.0 ? a.rhs[k] - dia_row_n(a.U, a.L, a.x, k) : 0.0;
        a.r[k] = r; a.rt[k] = r; a.p[k] = r; a.ph[k] = r * d; a.v[k] = 0.0;
        in[0] += r * r;
    }
    grid_reduce_n<BLOCK, 5>(a.counter, epoch, flip, in, a.partial, sh, out);
    double rho = out[0], err = xlung > 0.0 ? sqrt(out[0] / xlung) : sqrt(out[0] / n);
    int niter = 0;
    if (rho == 0.0 || err <= a.tol) { if (t0 == 0) { a.out->pcg_niter = 1; a.out->pcg_err = err; a.out->pad = (int)epoch; } return; }
    if (LINE) { line_solve(a, a.p, a.ph, t0, stride); grid_barrier(a.counter, epoch); }
    for (;;) {
        ++niter;
        // ---- v = J ph, sigma = (rt, v)
        in[0] = in[1] = in[2] = in[3] = in[4] = 0.0;
        for (int k = t0; k < n; k += stride) {
            if (PF && k + stride < n) bicg_prefetch_row(a.U, a.L, k + stride);
            double v = di[k] != 0.0 ? dia_row_n(a.U, a.L, a.ph, k) : 0.0;
            a.v[k] = v;
            in[0] += a.rt[k] * v;
        }
        grid_reduce_n<BLOCK, 5>(a.counter, epoch, flip, in, a.partial, sh, out);
        const double alpha = rho / out[0];
        // ---- s = r - alpha v, sh = M^-1 s
        // Boustrophedon sweeps (a.zigzag): the Jacobian of a large mesh (config 3: 197 MB) does not fit the 126 MB L2, so two
        // product sweeps in the same direction re-read ALL of it from HBM.  This pass and the second product run from the last row
        // back to the first: they start on the rows the first product touched last, which are still in the L2.
        if (ZZ) for (int k = klast; k >= 0; k -= stride) {
            double s = a.r[k] - alpha * a.v[k];
            a.s[k] = s;
            if (!LINE) a.sh[k] = s * di[k];
        }
        else
        for (int k = t0; k < n; k += stride) {
            double s = a.r[k] - alpha * a.v[k];
            a.s[k] = s;
            if (!LINE) a.sh[k] = s * di[k];
        }
        grid_barrier(a.counter, epoch);
        if (LINE) { line_solve(a, a.s, a.sh, t0, stride); grid_barrier(a.counter, epoch); }
        // ---- t = J sh; (t,s), (t,t), (rt,s), (rt,t)
        in[0] = in[1] = in[2] = in[3] = in[4] = 0.0;
        if (ZZ) for (int k = klast; k >= 0; k -= stride) {
            if (PF && k - stride >= 0) bicg_prefetch_row(a.U, a.L, k - stride);
            double t = di[k] != 0.0 ? dia_row_n(a.U, a.L, a.sh, k) : 0.0, s = a.s[k], rt = a.rt[k];
            a.t[k] = t;
            in[0] += t * s; in[1] += t * t; in[2] += rt * s; in[3] += rt * t;
        }
        else
        for (int k = t0; k < n; k += stride) {
            if (PF && k + stride < n) bicg_prefetch_row(a.U, a.L, k + stride);
            double t = di[k] != 0.0 ? dia_row_n(a.U, a.L, a.sh, k) : 0.0, s = a.s[k], rt = a.rt[k];
            a.t[k] = t;
            in[0] += t * s; in[1] += t * t; in[2] += rt * s; in[3] += rt * t;
        }
        grid_reduce_n<BLOCK, 5>(a.counter, epoch, flip, in, a.partial, sh, out);
        const double omega = out[1] > 0.0 ? out[0] / out[1] : 0.0;
        const double rho_new = out[2] - omega * out[3];
        const bool breakdown = omega == 0.0 || rho_new == 0.0;
        const double beta = breakdown ? 0.0 : (rho_new / rho) * (alpha / omega);
        // ---- x += alpha ph + omega sh; r = s - omega t; next p = r + beta (p - omega v), ph = M^-1 p; ||r||^2 summed directly
        // (the algebraic form (s,s) - 2 omega (t,s) + omega^2 (t,t) cancels catastrophically on ill-conditioned systems)
        in[0] = 0.0;
        for (int k = t0; k < n; k += stride) {
            a.x[k] = a.x[k] + alpha * a.ph[k] + omega * a.sh[k];
            double r = a.s[k] - omega * a.t[k];
            a.r[k] = r;
            in[0] += r * r;
            double p = r + beta * (a.p[k] - omega * a.v[k]);
            a.p[k] = p;
            if (!LINE) a.ph[k] = p * di[k];
        }
        grid_reduce_n<BLOCK, 5>(a.counter, epoch, flip, in, a.partial, sh, out);
        err = xlung > 0.0 ? sqrt(out[0] / xlung) : sqrt(out[0] / n);
        if (!(err > a.tol) || niter >= a.itmax || breakdown) break;
        if (LINE) { line_solve(a, a.p, a.ph, t0, stride); grid_barrier(a.counter, epoch); }
        rho = rho_new;
    }
    // a breakdown (NaN / zero inner products) without convergence is reported as "ITMXCG reached" so that FLOW3D back-steps
    if (t0 == 0) { a.out->pcg_niter = (err > a.tol || !(err == err)) ? max(niter, a.itmax) : niter; a.out->pcg_err = err; a.out->pad = (int)epoch; }
}
