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

// ------------------------------------------------------------------------------------------
// device-side parameter blocks
// ------------------------------------------------------------------------------------------
struct Diag {           // the 8 upper diagonals of a symmetric matrix
    double *d[NDIAG];   // d[0] = main diagonal
    int off[NDIAG];
};

struct Soil {           // nodal van Genuchten constants (SRC/tpnodi.f, SRC/chparm.f:22-35)
    const double *vgn, *vgm, *vgpsat, *vgpnot, *rr /* VGRMC/PNODI */, *snodi, *pnodi, *vgn1, *vgnr, *vgpsn, *vgmr, *vgm52, *vgmm1;
};

// scalars that cross to the host once per nonlinear iteration
struct IterOut {
    double pl2, pinf, fl2, finf, pnew_ik, pold_ik, dstore;
    double adin, adout, anin, anout, ndin, ndout;
    double pcg_err;
    int ikmax, pcg_niter, ponding, pad;
};
struct StepOut {        // once per accepted step
    double store1, apot, aact, ovflow, reflow, q_out1, q_out2, ak_max;
    int nhort, ndunn, npond, nsat, nsurf, hgflag[9], pad;
};

// ------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
// fixed-order block sum: every thread gets nothing, thread 0 gets the total
template <int NT_>
__device__ __forceinline__ double block_sum(double v, double *sh /* [32] */)
{
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double t = 0.0;
    if (w == 0) {
        t = lane < (NT_ >> 5) ? sh[lane] : 0.0;
        t = warp_sum(t);
    }
    return t;
}

// van Genuchten functions, SRC/fvgse.f:9-24, SRC/fvgkr.f, SRC/fvgdse.f (threshold psi < -1e-14)
__device__ __forceinline__ double fvgse(double psi, double psat, double n, double m)
{
    if (psi < -1.0e-14) {
        double beta = pow(fabs(psi / psat), n);
        return pow(fabs(1.0 / (beta + 1.0)), m);
    }
    return 1.0;
}
__device__ __forceinline__ double fvgkr(double psi, double se, double m, double mr)
{
    if (psi < -1.0e-14) {
        double omega = pow(fabs(se), mr);
        double v1 = 1.0 - pow(fabs(1.0 - omega), m);
        return sqrt(se) * v1 * v1;
    }
    return 1.0;
}
__device__ __forceinline__ double fvgdse(double psi, double psat, double n, double n1, double nr, double psn)
{
    if (psi < -1.0e-14) {
        double beta = pow(fabs(psi / psat), n);
        double b1 = beta + 1.0, b1r = 1.0 / b1;
        return n1 * (pow(fabs(psi), n1) / psn) * pow(fabs(b1), nr) * b1r * b1r;
    }
    return 0.0;
}

// The three van Genuchten functions of one node with 3 instead of 6 pow() calls (k_curves is bound by the instruction issue of
// the fp64 pow, ncu: issue 68 %, DRAM 16 %).  With b1 = 1 + beta, beta = |psi/psat|^n and se = b1^-m, m = 1 - 1/n:
//   FVGKR's  omega = se^(1/m)      = 1/b1, and 1 - omega = beta/b1 (no cancellation near saturation);
//   FVGDSE's |psi|^(n-1) / |psat|^n = beta/|psi|  and  b1^(1/n) = b1^(1-m) = b1 se.
// The values agree with fvgse / fvgkr / fvgdse to a few ulp (the parity gates are 1e-6); se itself is computed as in fvgse.
__device__ __forceinline__ void vg_node(double psi, double psat, double n, double m, double n1, bool need_d, double &se, double &kr, double &dse)
{
    if (psi < -1.0e-14) {
        const double beta = pow(fabs(psi / psat), n), b1 = beta + 1.0, b1r = 1.0 / b1;
        se = pow(fabs(b1r), m);
        const double v1 = 1.0 - pow(beta * b1r, m);
        kr = sqrt(se) * v1 * v1;
        dse = need_d ? n1 * (beta / fabs(psi)) * (b1 * se) * b1r * b1r : 0.0;
    } else { se = 1.0; kr = 1.0; dse = 0.0; }
}

// Huyakorn (IVGHU = 2, 3) and Brooks-Corey (IVGHU = 4) models: global parameters, constants of SRC/chparm.f:79-106
struct CurveModel {
    int ivghu;
    double hupsia, hubeta, hugama, huswr, huswr1, hualb, hugam1, hugb, hun, hua, hub2a, huab;
    double bcpsat, bcbeta, bcrmc, bcb1, bcbps, bc23b;
};
// SRC/fhuse.f, fhudse.f, fhukr2.f, fhukr3.f, fbcse.f, fbcdse.f, fbckr.f: saturation sw, kr and d(sw)/d(psi) of one node
__device__ __forceinline__ void curve_alt(const CurveModel &c, double psi, double pnodi, double &sw, double &kr, double &dsw, bool need_d)
{
    if (c.ivghu == 4) {
        const double porm = (pnodi - c.bcrmc) / pnodi;
        if (psi < c.bcpsat) {
            const double q = fabs(c.bcpsat / psi);
            sw = porm * pow(q, c.bcbeta) + c.bcrmc / pnodi;
            kr = pow(q, c.bc23b);
            dsw = need_d ? porm * (c.bcbps * pow(q, c.bcb1)) : 0.0;
        } else { sw = porm * 1.0 + c.bcrmc / pnodi; kr = 1.0; dsw = need_d ? porm * 0.0 : 0.0; }
        return;
    }
    if (psi < c.hupsia) {
        const double pap = c.hupsia - psi, lambda = c.hualb * pow(pap, c.hubeta), lamr = 1.0 / (1.0 + lambda);
        const double se = pow(lamr, c.hugama);
        sw = c.huswr1 * se + c.huswr;
        kr = c.ivghu == 2 ? pow(se, c.hun) : pow(10.0, c.hua * se * se + c.hub2a * se + c.huab);
        dsw = need_d ? c.huswr1 * ((c.hugb * lambda / pap) * pow(lamr, c.hugam1)) : 0.0;
    } else { sw = c.huswr1 * 1.0 + c.huswr; kr = 1.0; dsw = need_d ? c.huswr1 * 0.0 : 0.0; }
}
// CHPIC0 for IVGHU = 2, 3, 4 (SRC/chpic0.f:51-99)
__global__ void k_curves_alt(int n, CurveModel c, const double *__restrict__ snodi, const double *__restrict__ pnodi, const double *__restrict__ ptnew,
                             const double *__restrict__ pnew, const double *__restrict__ ptimep, int do_timep, double *__restrict__ sw,
                             double *__restrict__ ckrw, double *__restrict__ et1, double *__restrict__ et2, double *__restrict__ swnew,
                             double *__restrict__ swtimep)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double po = pnodi[i], sn = snodi[i], psi = ptnew[i];
        double w, kr, dsw, dum1, dum2;
        curve_alt(c, psi, po, w, kr, dsw, true);
        const double etai = w * sn + po * dsw;
        sw[i] = w; ckrw[i] = kr;
        et1[i] = w * sn;
        et2[i] = (etai - w * sn) / po;
        const double pn = pnew[i];
        if (pn == psi) swnew[i] = w; else { curve_alt(c, pn, po, w, dum1, dum2, false); swnew[i] = w; }
        if (do_timep) { curve_alt(c, ptimep[i], po, w, dum1, dum2, false); swtimep[i] = w; }
    }
}
__global__ void k_chvelo_alt(int n, CurveModel c, const double *__restrict__ pnodi, const double *__restrict__ psiv, const double *__restrict__ volnod,
                             double *__restrict__ sw, double *__restrict__ ckrw, double *__restrict__ partial, const unsigned char *__restrict__ own)
{
    __shared__ double sh[32];
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double w, kr, d;
        curve_alt(c, psiv[i], pnodi[i], w, kr, d, false);
        sw[i] = w; ckrw[i] = kr;
        if (!own || (own[i] & 1)) acc += w * volnod[i] * pnodi[i];
    }
    double t = block_sum<RED_BLOCK>(acc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// Extended van Genuchten (IVGHU = 1): SRC/fxvmc.f, fxvkr.f, fxvdmc.f.  Above the head PNOT (where the slope of the van Genuchten
// curve has fallen to the specific storage) the moisture content continues linearly with slope SS.  With IVGHU = 1 Soil::vgpnot
// holds PNOT (bisection of SRC/chparm.f:36-78, done once on the host) and Soil::rr the residual moisture content VGRMC itself.
__device__ __forceinline__ void xvg_node(const Soil &s, int i, double psi, bool need_kr, bool need_d, double &sw, double &kr, double &dmc)
{
    const double n = s.vgn[i], m = s.vgm[i], psat = s.vgpsat[i], pnot = s.vgpnot[i], rmc = s.rr[i], ss = s.snodi[i], por = s.pnodi[i];
    const double tsr = por - rmc;
    kr = 1.0; dmc = ss;
    if (psi < pnot) {
        const double beta = pow(fabs(psi / psat), n), b1 = beta + 1.0, b1r = 1.0 / b1;
        sw = (rmc + (tsr / pow(b1, m))) / por;
        if (need_d) dmc = s.vgn1[i] * tsr * (pow(fabs(psi), s.vgn1[i]) / s.vgpsn[i]) * pow(b1, s.vgnr[i]) * b1r * b1r;
        if (need_kr) { const double v1 = pow(b1, m) - pow(beta, m); kr = pow(b1r, s.vgm52[i]) * v1 * v1; }
    } else {
        const double b01 = pow(fabs(pnot / psat), n) + 1.0;
        sw = (rmc + tsr * pow(b01, -m) + ss * (psi - pnot)) / por;
        if (need_kr && psi < -1.0e-14) {
            const double beta = pow(fabs(psi / psat), n), b1 = beta + 1.0, v1 = pow(b1, m) - pow(beta, m);
            kr = pow(1.0 / b1, s.vgm52[i]) * v1 * v1;
        }
    }
}
// CHPIC0 for IVGHU = 1 (SRC/chpic0.f:37-50)
__global__ void k_curves_xvg(int n, Soil s, const double *__restrict__ ptnew, const double *__restrict__ pnew, const double *__restrict__ ptimep,
                             int do_timep, double *__restrict__ sw, double *__restrict__ ckrw, double *__restrict__ et1, double *__restrict__ et2,
                             double *__restrict__ swnew, double *__restrict__ swtimep)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double psi = ptnew[i], sn = s.snodi[i], po = s.pnodi[i];
        double w, kr, etai, dum1, dum2;
        xvg_node(s, i, psi, true, true, w, kr, etai);
        sw[i] = w; ckrw[i] = kr;
        et1[i] = w * sn;
        et2[i] = (etai - w * sn) / po;
        const double pn = pnew[i];
        if (pn == psi) swnew[i] = w; else { xvg_node(s, i, pn, false, false, w, dum1, dum2); swnew[i] = w; }
        if (do_timep) { xvg_node(s, i, ptimep[i], false, false, w, dum1, dum2); swtimep[i] = w; }
    }
}
// CHVELO for IVGHU = 1 (SRC/chvelo.f:34-39) fused with STORCAL's sum term
__global__ void k_chvelo_xvg(int n, Soil s, const double *__restrict__ psiv, const double *__restrict__ volnod, double *__restrict__ sw,
                             double *__restrict__ ckrw, double *__restrict__ partial, const unsigned char *__restrict__ own)
{
    __shared__ double sh[32];
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double w, kr, d;
        xvg_node(s, i, psiv[i], true, false, w, kr, d);
        sw[i] = w; ckrw[i] = kr;
        if (!own || (own[i] & 1)) acc += w * volnod[i] * s.pnodi[i];
    }
    double t = block_sum<RED_BLOCK>(acc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// ------------------------------------------------------------------------------------------
// K1: moisture curves per node (PICUNS -> CHPIC0, SRC/picuns.f:22-48, SRC/chpic0.f:23-36)
// ------------------------------------------------------------------------------------------
__global__ void k_curves(int n, Soil s, const double *__restrict__ ptnew, const double *__restrict__ pnew,
                         const double *__restrict__ ptimep, int do_timep, double *__restrict__ sw,
                         double *__restrict__ ckrw, double *__restrict__ et1, double *__restrict__ et2,
                         double *__restrict__ swnew, double *__restrict__ swtimep)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double n_ = s.vgn[i], m = s.vgm[i], psat = s.vgpsat[i], pnot = s.vgpnot[i], rr = s.rr[i];
        double psi = ptnew[i];
        double se, kr, dse;
        vg_node(psi, psat, n_, m, s.vgn1[i], true, se, kr, dse);
        double w = pnot * se + rr;
        sw[i] = w;
        et1[i] = w * s.snodi[i];
        et2[i] = pnot * dse;
        ckrw[i] = kr;
        // PNEW can differ from PTNEW at ponded surface nodes even when TETAF = 1 (PONDUPD runs after WEIGHT)
        double pn = pnew[i];
        swnew[i] = pn == psi ? w : pnot * fvgse(pn, psat, n_, m) + rr;
        if (do_timep) swtimep[i] = pnot * fvgse(ptimep[i], psat, n_, m) + rr;
    }
}
// KSLOPE = 1, 2 (SRC/chpic1.f:26-50, SRC/chpic2.f:24-46; IVGHU = 0): dSe/dpsi as the chord slope between the current and the previous
// nonlinear iterate wherever they differ by TOLKSL or more, else analytical (1) / centred difference over 2 TOLKSL (2)
__global__ void k_curves_chord(int n, Soil s, int kslope, double tolksl, const double *__restrict__ ptnew, const double *__restrict__ ptold,
                               const double *__restrict__ pnew, const double *__restrict__ ptimep, int do_timep, double *__restrict__ sw,
                               double *__restrict__ ckrw, double *__restrict__ et1, double *__restrict__ et2,
                               double *__restrict__ swnew, double *__restrict__ swtimep)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double n_ = s.vgn[i], m = s.vgm[i], psat = s.vgpsat[i], pnot = s.vgpnot[i], rr = s.rr[i];
        const double psi = ptnew[i], pold = ptold[i], dp = psi - pold;
        const bool small = fabs(dp) < tolksl;
        double se, kr, dse;
        vg_node(psi, psat, n_, m, s.vgn1[i], small && kslope == 1, se, kr, dse);
        if (!small) dse = (se - fvgse(pold, psat, n_, m)) / dp;
        else if (kslope == 2) dse = (fvgse(psi + tolksl, psat, n_, m) - fvgse(psi - tolksl, psat, n_, m)) / (2.0 * tolksl);
        const double w = pnot * se + rr;
        sw[i] = w;
        et1[i] = w * s.snodi[i];
        et2[i] = pnot * dse;
        ckrw[i] = kr;
        const double pn = pnew[i];
        swnew[i] = pn == psi ? w : pnot * fvgse(pn, psat, n_, m) + rr;
        if (do_timep) swtimep[i] = pnot * fvgse(ptimep[i], psat, n_, m) + rr;
    }
}
// CHVELO (SRC/chvelo.f, IVGHU=0) fused with STORCAL's sum term (SRC/storcal.f)
__global__ void k_chvelo(int n, Soil s, const double *__restrict__ psiv, const double *__restrict__ volnod,
                         double *__restrict__ sw, double *__restrict__ ckrw, double *__restrict__ partial, const unsigned char *__restrict__ own)
{
    __shared__ double sh[32];
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double psi = psiv[i], m = s.vgm[i];
        double se, kr, dse;
        vg_node(psi, s.vgpsat[i], s.vgn[i], m, 0.0, false, se, kr, dse);
        double w = s.vgpnot[i] * se + s.rr[i];
        sw[i] = w;
        ckrw[i] = kr;
        if (!own || (own[i] & 1)) acc += w * volnod[i] * s.pnodi[i];
    }
    double t = block_sum<RED_BLOCK>(acc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// ------------------------------------------------------------------------------------------
// K2: node -> element averages (NODELT, SRC/nodelt.f:19-26) of kr and ET1
// ------------------------------------------------------------------------------------------
__global__ void k_tet_avg(int nt, const int4 *__restrict__ tet, const double *__restrict__ ckrw,
                          const double *__restrict__ et1, double *__restrict__ krt, double *__restrict__ e1t)
{
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nt; e += gridDim.x * blockDim.x) {
        int4 t = tet[e];
        krt[e] = (((ckrw[t.x] + ckrw[t.y]) + ckrw[t.z]) + ckrw[t.w]) * 0.25;
        e1t[e] = (((et1[t.x] + et1[t.y]) + et1[t.z]) + et1[t.w]) * 0.25;
    }
}

// ------------------------------------------------------------------------------------------
// K3: atomic-free assembly (ASSPIC, SRC/asspic.f:26-51; RHSGRV, SRC/rhsgrv.f:20-30).
// Every matrix slot owns a static list of (tet, coefficient) pairs sorted by tet, i.e. the
// reference's TETJA scatter turned into a gather; the sum runs in the reference's element order.
// ------------------------------------------------------------------------------------------
// The lists are stored ELL-style, transposed: entry c of row k of diagonal d sits at [c][k], so that
// consecutive threads (rows) read consecutive addresses; rows with fewer entries are padded with coef 0.
struct EllFamily { const int *tet; const double *coef; const double *coef2; int w; int pad; };   // node.pad = 1: node.tet == diag[0].tet entry for entry
struct EllPlan { EllFamily diag[NDIAG]; EllFamily node; long long ld; };
__global__ void __launch_bounds__(RED_BLOCK) k_assemble(int n, EllPlan P, const double *__restrict__ krt, const double *__restrict__ e1t,
                                                        Diag A, double *__restrict__ grav, double *__restrict__ m2)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const EllFamily f = P.node;
        double g = 0.0, m = 0.0;
        if (P.node.pad) {
            // the node family lists the tets around node k in the same order as the main-diagonal family (both are filled
            // tet by tet): one index stream and one gather of kr serve both
            const EllFamily f0 = P.diag[0];
            double acc = 0.0;
            for (int c = 0; c < f0.w; ++c) {
                size_t q = (size_t)c * P.ld + k;
                const int t = f0.tet[q];
                const double kr = krt[t];
                acc += kr * f0.coef[q];
                g += kr * f.coef[q];
                m += e1t[t] * f.coef2[q];
            }
            A.d[0][k] = acc;
        }
#pragma unroll
        for (int d = 0; d < NDIAG; ++d) {
            if (d == 0 && P.node.pad) continue;
            const EllFamily fd = P.diag[d];
            double acc = 0.0;
            for (int c = 0; c < fd.w; ++c) {
                size_t q = (size_t)c * P.ld + k;
                acc += krt[fd.tet[q]] * fd.coef[q];
            }
            A.d[d][k] = acc;
        }
        if (!P.node.pad) {
            for (int c = 0; c < f.w; ++c) {
                size_t q = (size_t)c * P.ld + k;
                int t = f.tet[q];
                g += krt[t] * f.coef[q];
                m += e1t[t] * f.coef2[q];
            }
        }
        grav[k] = g;
        m2[k] = m;
    }
}

// The same gather with the tet indices DERIVED instead of stored.  On the prism-split DEM mesh the tets around node (layer l, row i,
// column j) are base(k) + a fixed offset, base(k) = 3 NTRI l + 6 (i NCOL + j); the list of offsets depends only on which of the
// 27 boundary classes (top / inner / bottom layer x north / inner / south row x west / inner / east column) the node is in.  The
// host builds the 27 offset tables from the stored lists and checks EVERY entry of every row against them (any mismatch keeps
// the stored indices), so this kernel reads 8 instead of 12 bytes per contribution: -20 % of the DRAM traffic that bounds it.
struct PlanGeom { const int *rel; int wrel, nnod, nc1, ncol, nrow, nstr, ntri3, nt; };
__global__ void __launch_bounds__(RED_BLOCK) k_assemble_a(int n, EllPlan P, PlanGeom G, const double *__restrict__ krt, const double *__restrict__ e1t,
                                                          Diag A, double *__restrict__ grav, double *__restrict__ m2)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int l = k / G.nnod, sidx = k - l * G.nnod, i = sidx / G.nc1, j = sidx - i * G.nc1;
        const int cls = ((l == 0 ? 0 : l == G.nstr ? 2 : 1) * 3 + (i == 0 ? 0 : i == G.nrow ? 2 : 1)) * 3 + (j == 0 ? 0 : j == G.ncol ? 2 : 1);
        const int base = G.ntri3 * l + 6 * (i * G.ncol + j);
        const int *__restrict__ rl = G.rel + (size_t)cls * NDIAG * G.wrel;
        {
            const EllFamily f0 = P.diag[0], f = P.node;
            double acc = 0.0, g = 0.0, m = 0.0;
            for (int c = 0; c < f0.w; ++c) {
                const size_t q = (size_t)c * P.ld + k;
                const int t = min(max(base + __ldg(rl + c), 0), G.nt - 1);
                const double kr = krt[t];
                acc += kr * f0.coef[q];
                g += kr * f.coef[q];
                m += e1t[t] * f.coef2[q];
            }
            A.d[0][k] = acc;
            grav[k] = g;
            m2[k] = m;
        }
#pragma unroll
        for (int d = 1; d < NDIAG; ++d) {
            const EllFamily fd = P.diag[d];
            double acc = 0.0;
            for (int c = 0; c < fd.w; ++c) {
                const size_t q = (size_t)c * P.ld + k;
                const int t = min(max(base + __ldg(rl + d * G.wrel + c), 0), G.nt - 1);
                acc += krt[t] * fd.coef[q];
            }
            A.d[d][k] = acc;
        }
    }
}

// symmetric DIA row product: (A x)_k from the 8 upper diagonals.  Branch free: every gathered vector
// carries NNOD zero-filled halo elements on both sides and structurally absent entries are stored as 0.
__device__ __forceinline__ double dia_row(const Diag &A, const double *__restrict__ diag0, const double *__restrict__ x, int k, int n)
{
    (void)n;
    double acc = diag0[k] * x[k];
#pragma unroll
    for (int d = 1; d < NDIAG; ++d) acc += A.d[d][k] * x[k + A.off[d]];
#pragma unroll
    for (int d = 1; d < NDIAG; ++d) acc += A.d[d][k - A.off[d]] * x[k - A.off[d]];
    return acc;
}

__device__ __forceinline__ bool is_dirichlet(int k, int nnod, const int *__restrict__ ifatm, const unsigned char *__restrict__ contp_flag)
{
    if (contp_flag && contp_flag[k]) return true;
    if (k < nnod) { int f = ifatm[k]; return f == 1 || f == 2; }
    return false;
}

// ------------------------------------------------------------------------------------------
// K4: RHS + LHS diagonal + boundary conditions in one pass
// (RHSPIC SRC/rhspic.f:22-38, CFMATP SRC/cfmatp.f:21-26, RHSGRV, BCPIC SRC/bcpic.f:33-86)
// ------------------------------------------------------------------------------------------
__global__ void k_rhs_lhs(int n, int nnod, Diag A, double tetaf, double rdt, const double *__restrict__ ptnew,
                          const double *__restrict__ pnew, const double *__restrict__ ptimep,
                          const double *__restrict__ swnew, const double *__restrict__ swtimep,
                          const double *__restrict__ m2, const double *__restrict__ m4, const double *__restrict__ et2,
                          const double *__restrict__ grav, const int *__restrict__ ifatm,
                          const unsigned char *__restrict__ contp_flag, const double *__restrict__ qneu,
                          const double *__restrict__ atmact, const double *__restrict__ atmold,
                          const double *__restrict__ qtranie, double *__restrict__ rhs, double *__restrict__ xt5,
                          double *__restrict__ diag_true, double *__restrict__ diag_bc, const double *__restrict__ dtp)
{
    if (dtp) rdt = dtp[1];      // graph replay: {DELTAT, 1/DELTAT} of the current step live in device memory (the launch arguments are frozen)
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        double ax = dia_row(A, A.d[0], ptnew, k, n);
        double b = -ax - m2[k] * rdt * (pnew[k] - ptimep[k]) - m4[k] * rdt * (swnew[k] - swtimep[k]) - grav[k];
        xt5[k] = b;
        double dt_ = tetaf * A.d[0][k] + m2[k] * rdt + (m4[k] * et2[k]) * rdt;
        diag_true[k] = dt_;
        bool dir = is_dirichlet(k, nnod, ifatm, contp_flag);
        if (dir) b = 0.0;
        if (qneu) b += qneu[k];
        if (k < nnod && ifatm[k] == 0) b = b + (tetaf * atmact[k] + (1.0 - tetaf) * atmold[k]);
        b = b - qtranie[k];
        rhs[k] = b;
        diag_bc[k] = dir ? 1.0e-9 * RMAX_ : dt_;
    }
}
__global__ void k_scale(long long n, double a, double *__restrict__ v)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) v[i] *= a;
}

// plain SpMV y = A x (used by cathy_debug_spmv and the roofline measurement)
__global__ void k_spmv(int n, Diag A, const double *__restrict__ diag0, const double *__restrict__ x, double *__restrict__ y)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) y[k] = dia_row(A, diag0, x, k, n);
}


// ==========================================================================================
// Row-block partition of ONE mesh over several GPUs (BASELINE config 5).  Each rank holds a window of DEM rows (its owned
// node rows + DD_W ghost node rows per interior side) with the same layer-major numbering and the same 15-point DIA stencil.
// Ranks exchange data through peer memory over NVLink (CUDA IPC mapped "boxes"): halo rows are STORED straight into the
// neighbour's inbox by the kernel that produces them, all-reduces are slot writes + system-scope release/acquire flags,
// summed in rank order on every rank (bit-identical results on all ranks, hence identical control flow).
// ==========================================================================================
#define DD_W 2
#define DD_MAXW 8
#define DD_NRED 24
#define DD_TIMEOUT_CYCLES 12000000000LL   // ~6 s: a lost peer turns into an error, not a hang
struct DDBox {                               // lives in each rank's device memory, mapped by all peers
    unsigned int ar_flag[DD_MAXW];           // sequence number of the last all-reduce contribution of rank r
    unsigned int halo_flag[2];               // [0]: from the north neighbour, [1]: from the south neighbour
    int geom[4];                             // column-major layout of this rank: first owned row lo, end hi, local rows, halo rows (k_pcg_tma)
    unsigned int pad[2];
    double ar_slot[2][DD_MAXW][DD_NRED];     // [parity][rank][value]
};
struct DDCtx {
    int world, rank, north, south;           // neighbour ranks (-1: none)
    DDBox *me;
    DDBox *peer[DD_MAXW];                    // peer[rank] == me
    double *inbox_me;                        // [2 parities][2 sides][hcap]
    double *inbox_peer[DD_MAXW];
    long long hcap;
    int nc1, nlay, nnod, own_a, own_b;
    unsigned int *seq;                       // [0] all-reduce sequence, [1] halo sequence (device resident, advanced by the kernels)
    int *err;
};
__device__ __forceinline__ void dd_wait(const unsigned int *flag, unsigned int target, int *err, int site = 1)
{
    if (*(volatile int *)err) return;
    long long t0 = clock64();
    for (;;) {
        unsigned int v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int)(v - target) >= 0) return;
        if (clock64() - t0 > DD_TIMEOUT_CYCLES) { *(volatile int *)err = site + 10 * (int)(target & 0xffffffu); return; }
    }
}
__device__ __forceinline__ void dd_release(unsigned int *flag, unsigned int v)
{
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(v) : "memory");
}
// All-reduce (sum, rank order) of NV block-uniform values; called by every thread of every block of a kernel whose blocks
// all hold the same v[] (after a grid-wide local reduction).  sh: NV doubles of shared memory.
template <int NV>
__device__ __forceinline__ void dd_allreduce(const DDCtx &c, unsigned int &seq, double (&v)[NV], double *sh)
{
    ++seq;
    const int par = seq & 1u;
    if (blockIdx.x == 0 && (int)threadIdx.x < c.world) {
        DDBox *dst = c.peer[threadIdx.x];
#pragma unroll
        for (int q = 0; q < NV; ++q) dst->ar_slot[par][c.rank][q] = v[q];
        dd_release(&dst->ar_flag[c.rank], seq);
    }
    if ((int)threadIdx.x < c.world) dd_wait(&c.me->ar_flag[threadIdx.x], seq, c.err, 1);
    __syncthreads();
    if (threadIdx.x < NV) {
        double acc = 0.0;
        for (int r = 0; r < c.world; ++r) acc += *(volatile double *)&c.me->ar_slot[par][r][threadIdx.x];
        sh[threadIdx.x] = acc;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NV; ++q) v[q] = sh[q];
    __syncthreads();
}
// element e of a halo message <-> (layer, row offset, column)
__device__ __forceinline__ long long dd_index(const DDCtx &c, long long e, int row0)
{
    int j = (int)(e % c.nc1);
    long long t = e / c.nc1;
    int w = (int)(t % DD_W), l = (int)(t / DD_W);
    return (long long)l * c.nnod + (long long)(row0 + w) * c.nc1 + j;
}
__device__ __forceinline__ void dd_send_rows(const DDCtx &c, unsigned int next_seq, const double *vec, long long tid, long long nthreads)
{   // my first / last DD_W owned rows -> the neighbours' south / north inboxes
    const long long E = (long long)DD_W * c.nlay * c.nc1;
    const int par = next_seq & 1u;
    bool any = false;
    if (c.north >= 0) { double *dst = c.inbox_peer[c.north] + ((size_t)par * 2 + 1) * c.hcap; for (long long e = tid; e < E; e += nthreads) { dst[e] = vec[dd_index(c, e, c.own_a)]; any = true; } }
    if (c.south >= 0) { double *dst = c.inbox_peer[c.south] + ((size_t)par * 2 + 0) * c.hcap; for (long long e = tid; e < E; e += nthreads) { dst[e] = vec[dd_index(c, e, c.own_b - DD_W)]; any = true; } }
    if (any) __threadfence_system();
}
// after a grid-wide barrier that follows the sends: publish, wait for the neighbours' rows, copy them into the ghost rows
__device__ __forceinline__ void dd_recv_rows(const DDCtx &c, unsigned int &seq, double *vec, long long tid, long long nthreads)
{
    ++seq;
    const int par = seq & 1u;
    if (blockIdx.x == 0 && threadIdx.x == 0 && c.north >= 0) dd_release(&c.peer[c.north]->halo_flag[1], seq);
    if (blockIdx.x == 0 && threadIdx.x == 1 && c.south >= 0) dd_release(&c.peer[c.south]->halo_flag[0], seq);
    if (threadIdx.x == 0 && c.north >= 0) dd_wait(&c.me->halo_flag[0], seq, c.err, 2);
    if (threadIdx.x == 1 && c.south >= 0) dd_wait(&c.me->halo_flag[1], seq, c.err, 3);
    __syncthreads();
    const long long E = (long long)DD_W * c.nlay * c.nc1;
    if (c.north >= 0) { const double *src = c.inbox_me + ((size_t)par * 2 + 0) * c.hcap; for (long long e = tid; e < E; e += nthreads) vec[dd_index(c, e, c.own_a - DD_W)] = *(volatile const double *)&src[e]; }
    if (c.south >= 0) { const double *src = c.inbox_me + ((size_t)par * 2 + 1) * c.hcap; for (long long e = tid; e < E; e += nthreads) vec[dd_index(c, e, c.own_b)] = *(volatile const double *)&src[e]; }
}
// stand-alone halo exchange of one N-vector between kernels of the nonlinear loop (two launches: the kernel boundary is the
// grid-wide barrier between "all rows stored" and "flag published")
__global__ void k_dd_send(DDCtx c, const double *__restrict__ vec)
{
    dd_send_rows(c, c.seq[1] + 1u, vec, (long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x);
}
__global__ void k_dd_recv(DDCtx c, double *__restrict__ vec, unsigned int *counter)
{   // counter: arrival count so that the LAST block to finish advances the sequence number
    unsigned int seq = c.seq[1];
    dd_recv_rows(c, seq, vec, (long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(counter, 1u) == gridDim.x - 1) { c.seq[1] = seq; *counter = 0u; }
    }
}
// cross-rank combination of the per-iteration scalars (sums in rank order; max-norm with the reference's "last node wins" tie
// rule on GLOBAL node numbers) and of the per-step scalars; one block
__global__ void k_dd_combine_iter(DDCtx c, IterOut *__restrict__ io, int gnnod, int lo_shift)
{
    __shared__ double sh[DD_NRED];
    unsigned int seq = c.seq[0];
    // the two norms arrive squared-rooted from k_norms_final: square them back for the sum
    double v[7] = {io->pl2 * io->pl2, io->fl2 * io->fl2, io->dstore, io->adin, io->adout, io->anin, io->anout};
    dd_allreduce<7>(c, seq, v, sh);
    // max part: every rank publishes (pinf, global ik, pnew_ik, pold_ik, finf) in its slot, all pick the same winner
    int ikl = io->ikmax, lay = ikl / c.nnod;
    double gik = (double)((long long)lay * gnnod + (ikl - lay * c.nnod) + lo_shift);
    ++seq;
    const int par = seq & 1u;
    if ((int)threadIdx.x < c.world) {
        DDBox *dst = c.peer[threadIdx.x];
        double *sl = dst->ar_slot[par][c.rank];
        sl[0] = io->pinf; sl[1] = gik; sl[2] = io->pnew_ik; sl[3] = io->pold_ik; sl[4] = io->finf;
        dd_release(&dst->ar_flag[c.rank], seq);
    }
    if ((int)threadIdx.x < c.world) dd_wait(&c.me->ar_flag[threadIdx.x], seq, c.err, 4);
    __syncthreads();
    if (threadIdx.x == 0) {
        double pinf = -1.0, ik = -1.0, pn = 0.0, po = 0.0, finf = 0.0;
        for (int r = 0; r < c.world; ++r) {
            const volatile double *sl = c.me->ar_slot[par][r];
            if (sl[0] > pinf || (sl[0] == pinf && sl[1] > ik)) { pinf = sl[0]; ik = sl[1]; pn = sl[2]; po = sl[3]; }
            finf = fmax(finf, sl[4]);
        }
        io->pl2 = sqrt(v[0]); io->fl2 = sqrt(v[1]); io->dstore = v[2]; io->adin = v[3]; io->adout = v[4]; io->anin = v[5]; io->anout = v[6];
        io->pinf = pinf; io->ikmax = (int)ik; io->pnew_ik = pn; io->pold_ik = po; io->finf = finf;
        c.seq[0] = seq;
    }
}
__global__ void k_dd_combine_step(DDCtx c, StepOut *__restrict__ so, double *__restrict__ extra3)
{
    __shared__ double sh[DD_NRED];
    unsigned int seq = c.seq[0];
    double v[21];
    v[0] = so->store1; v[1] = so->apot; v[2] = so->aact; v[3] = so->ovflow; v[4] = so->reflow;
    v[5] = so->nhort; v[6] = so->ndunn; v[7] = so->npond; v[8] = so->nsat;
    for (int q = 0; q < 9; ++q) v[9 + q] = so->hgflag[q];
    for (int q = 0; q < 3; ++q) v[18 + q] = extra3 ? extra3[q] : 0.0;
    dd_allreduce<21>(c, seq, v, sh);
    if (threadIdx.x == 0) {
        so->store1 = v[0]; so->apot = v[1]; so->aact = v[2]; so->ovflow = v[3]; so->reflow = v[4];
        so->nhort = (int)v[5]; so->ndunn = (int)v[6]; so->npond = (int)v[7]; so->nsat = (int)v[8];
        for (int q = 0; q < 9; ++q) so->hgflag[q] = (int)v[9 + q];
        if (extra3) for (int q = 0; q < 3; ++q) extra3[q] = v[18 + q];
        c.seq[0] = seq;
    }
}

// ------------------------------------------------------------------------------------------
// K5-K7: the whole SYMSLV (SRC/solscal-extended.f:4669-4699) as ONE persistent cooperative
// kernel: preconditioner set-up, x0 = M^-1 b, and the GRADDP recurrence (:1260-1380) with two
// grid-wide barriers per iteration.  Reductions are fixed-order (block partials, then every block
// adds the partials in the same order), so results are bit-reproducible run to run.
//   phase A: p = z + beta p_old (recomputed on the fly for the neighbours), B = A p, (p.r), (p.B)
//   phase B: r -= alfa B, x += alfa p, z = M^-1 r, (B.z), ||r_free||^2
// Residual norm excludes Dirichlet rows exactly like GRADDP (:1286-1297, :1356-1371).
// ------------------------------------------------------------------------------------------
struct PcgArgs {
    int n, nnod, itmax;
    double tol;
    Diag A;
    const double *diag;      // main diagonal with the Dirichlet penalty
    const double *rhs;
    double *x, *r, *z, *p0, *p1, *bv;
    const int *ifatm;
    const unsigned char *contp_flag;
    double *partial;         // [3][gridDim.x]
    unsigned int *counter;   // grid barrier counter (monotonic)
    unsigned int epoch0;     // its value at launch
    IterOut *out;
    int prefetch;               // 1: software prefetch of the next grid-stride row into L2
    const unsigned char *own;   // row-block partition: bit0 = owned row, bit1 / bit2 = row is sent to the north / south neighbour
    DDCtx dd;
    int rows_cta;               // k_pcg_res: rows owned by one CTA (multiple of 32)
    int xres;                   // k_pcg_res: 1 = the solution vector lives in shared memory too
    int cm;                     // k_pcg: 1 = the arrays are in the column-major permutation (Dirichlet rows are recognised by their penalty diagonal)
};

// Grid-wide barrier for the persistent kernel: one arrival per block on a monotonically increasing counter
// (release), then a spin on an acquire load.  All blocks are co-resident (cooperative launch).
__device__ __forceinline__ void grid_barrier(unsigned int *counter, unsigned int &epoch)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch += gridDim.x;
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned int v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while ((int)(v - epoch) < 0);
    }
    __syncthreads();
}
// three sums at once: block partials (one shared-memory round), one grid barrier, then every block adds the
// partials in the same fixed order -> bit-reproducible and identical in all blocks
template <int BLOCK, bool CUSTOM>
__device__ __forceinline__ void grid_reduce3(cg::grid_group &grid, unsigned int *counter, unsigned int &epoch, double a, double b, double c,
                                             double *partial, double (*sh)[3], double &ra, double &rb, double &rc)
{
    const int nb = gridDim.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    if (lane == 0) { sh[w][0] = a; sh[w][1] = b; sh[w][2] = c; }
    __syncthreads();
    if (w == 0) {
        double t0 = lane < BLOCK / 32 ? sh[lane][0] : 0.0, t1 = lane < BLOCK / 32 ? sh[lane][1] : 0.0, t2 = lane < BLOCK / 32 ? sh[lane][2] : 0.0;
        t0 = warp_sum(t0); t1 = warp_sum(t1); t2 = warp_sum(t2);
        if (lane == 0) { partial[blockIdx.x] = t0; partial[nb + blockIdx.x] = t1; partial[2 * nb + blockIdx.x] = t2; }
    }
    if (CUSTOM) grid_barrier(counter, epoch); else grid.sync();
    if (w < 3) {
        double s0 = 0.0, s1 = 0.0;
        const volatile double *pp = partial + w * nb;
        int i = lane;
        for (; i + 32 < nb; i += 64) { s0 += pp[i]; s1 += pp[i + 32]; }
        if (i < nb) s0 += pp[i];
        double t = warp_sum(s0 + s1);
        if (lane == 0) sh[0][w] = t;
    }
    __syncthreads();
    ra = sh[0][0]; rb = sh[0][1]; rc = sh[0][2];
    __syncthreads();
}

// two sums at once, second generation (k_pcg_res, k_pcg_res2): half-warp butterflies, double-buffered partials, see k_pcg_res2
#define FULLMASK 0xffffffffu
template <int BLOCK>
__device__ __forceinline__ void grid_reduce2(unsigned int *counter, unsigned int &epoch, unsigned int &par, double a, double b, double *partial,
                                             double (*sh)[2], double (*res)[2], double &ra, double &rb)
{
    static_assert(BLOCK == 1024, "32 warps: the second level is one half-warp butterfly");
    const int nb = gridDim.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const bool hi = lane >= 16;
    double keep = hi ? b : a, send = hi ? a : b;
    keep += __shfl_xor_sync(FULLMASK, send, 16);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) keep += __shfl_xor_sync(FULLMASK, keep, o);
    if ((lane & 15) == 0) sh[w][hi] = keep;
    __syncthreads();
    double *pp = partial + (size_t)par * 2 * nb;   // [2][nb], buffer of this reduction
    if (w == 0) {
        double v = hi ? sh[lane - 16][1] + sh[lane][1] : sh[lane][0] + sh[lane + 16][0];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
        if ((lane & 15) == 0) pp[(hi ? nb : 0) + blockIdx.x] = v;
        __syncwarp();
        epoch += nb;
        // only thread 0 spins and nobody of its warp waits at a __syncwarp meanwhile: a lane spinning next to parked lanes of
        // the same warp costs +1.5 us per reduction on B200 (tools/bench_barrier4.cu)
        if (lane == 0) {
            __threadfence();
            atomicAdd(counter, 1u);
            unsigned int c;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(c) : "l"(counter) : "memory"); } while ((int)(c - epoch) < 0);
        }
    } else
        epoch += nb;
    __syncthreads();
    if (w < 2) {      // warp 0 sums the first quantity, warp 1 the second: independent loads, fixed order
        constexpr int MAXJ = 5;    // up to 160 CTAs (B200: 148)
        double v[MAXJ];
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) {
            const int i = lane + 32 * j;
            v[j] = 0.0;
            if (i < nb) asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v[j]) : "l"(pp + w * nb + i) : "memory");
        }
        double t = (((v[0] + v[1]) + v[2]) + v[3]) + v[4];
        for (int i = lane + 32 * MAXJ; i < nb; i += 32) t += ((volatile double *)pp)[w * nb + i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(FULLMASK, t, o);
        if (lane == 0) res[par][w] = t;
    }
    __syncthreads();
    ra = res[par][0]; rb = res[par][1];
    par ^= 1u;
}
// The same two sums when the whole solve runs in ONE thread-block cluster (small meshes): every CTA pushes its pair of partial sums
// into the slot it owns in every CTA's shared memory (st.shared::cluster), one hardware cluster barrier (release / acquire at
// cluster scope, ~0.2 us instead of the ~2 us of the global-memory barrier above), then every thread adds the slots in rank order
// -> the same value in all CTAs, bit-reproducible.  Double-buffered like grid_reduce2: one barrier per reduction suffices.
constexpr int PCG_CL_MAX = 16;
template <int BLOCK>
__device__ __forceinline__ void cluster_reduce2(cg::cluster_group &cl, unsigned int &par, double a, double b, double (*sh)[2],
                                                double (*cp)[PCG_CL_MAX][2], double &ra, double &rb)
{
    static_assert(BLOCK == 1024, "32 warps: the second level is one half-warp butterfly");
    const int nc = (int)cl.num_blocks(), lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const bool hi = lane >= 16;
    double keep = hi ? b : a, send = hi ? a : b;
    keep += __shfl_xor_sync(FULLMASK, send, 16);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) keep += __shfl_xor_sync(FULLMASK, keep, o);
    if ((lane & 15) == 0) sh[w][hi] = keep;
    __syncthreads();
    if (w == 0) {
        double v = hi ? sh[lane - 16][1] + sh[lane][1] : sh[lane][0] + sh[lane + 16][0];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
        const double va = __shfl_sync(FULLMASK, v, 0), vb = __shfl_sync(FULLMASK, v, 16);
        if (lane < nc) {
            double *dst = cl.map_shared_rank(&cp[par][cl.block_rank()][0], lane);
            *reinterpret_cast<double2 *>(dst) = make_double2(va, vb);
        }
    }
    cl.sync();
    double s0 = 0.0, s1 = 0.0;
    for (int c = 0; c < nc; ++c) { const double2 q = *reinterpret_cast<const double2 *>(&cp[par][c][0]); s0 += q.x; s1 += q.y; }
    ra = s0; rb = s1;
    par ^= 1u;
}
__device__ __forceinline__ void l2_prefetch(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
template <int BLOCK, bool CUSTOM, bool DD, int MINB = 1024 / BLOCK>
__global__ void __launch_bounds__(BLOCK, MINB) k_pcg(PcgArgs a)
{
    const bool PF = a.prefetch != 0;
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[BLOCK / 32][3];
    __shared__ double shdd[4];
    unsigned int epoch = a.epoch0;
    unsigned int seq_ar = 0, seq_h = 0;
    if (DD) { seq_ar = a.dd.seq[0]; seq_h = a.dd.seq[1]; }
    const int n = a.n, stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    const double *__restrict__ dg = a.diag;
    const unsigned char *__restrict__ own = a.own;
    // x0 = M^-1 b ; xlung = ||b_free||^2   (PRODDP call at :4686, XLUNG at :1286-1297)
    double xl = 0.0;
    for (int k = t0; k < n; k += stride) {
        double b = a.rhs[k];
        a.x[k] = b / dg[k];
        if ((!DD || (own[k] & 1)) && !(a.cm ? dg[k] > 1.0e80 : is_dirichlet(k, a.nnod, a.ifatm, a.contp_flag))) xl += b * b;
    }
    double xlung, d1, d2;
    grid_reduce3<BLOCK, CUSTOM>(grid, a.counter, epoch, xl, 0.0, 0.0, a.partial, sh, xlung, d1, d2);
    if (DD) {   // global ||b||^2, and x0 on the ghost rows from their owners
        double v[1] = {xlung};
        dd_allreduce<1>(a.dd, seq_ar, v, shdd);
        xlung = v[0];
        dd_send_rows(a.dd, seq_h + 1u, a.x, t0, stride);
        grid_barrier(a.counter, epoch);
        dd_recv_rows(a.dd, seq_h, a.x, t0, stride);
        grid_barrier(a.counter, epoch);
    }
    // r = b - A x0 ; z = M^-1 r ; p_old = 0 so that p = z in the first phase A
    for (int k = t0; k < n; k += stride) {
        double r = a.rhs[k] - dia_row(a.A, dg, a.x, k, n);
        a.r[k] = r;
        a.z[k] = r / dg[k];
        a.p0[k] = 0.0;
    }
    if (CUSTOM) grid_barrier(a.counter, epoch); else grid.sync();
    if (DD) {
        dd_send_rows(a.dd, seq_h + 1u, a.z, t0, stride);
        grid_barrier(a.counter, epoch);
        dd_recv_rows(a.dd, seq_h, a.z, t0, stride);
        grid_barrier(a.counter, epoch);
    }
    double beta = 0.0, err = 0.0;
    double *pold = a.p0, *pnew = a.p1;
    int niter = 1;
    for (;;) {
        // ---- phase A
        double s_pr = 0.0, s_pb = 0.0;
        {
            const double *z = a.z;      // NOT __restrict__/read-only: rewritten every iteration by other SMs
            const double *po = pold;
            for (int k = t0; k < n; k += stride) {
                if (PF && k + stride < n) {   // pull the next row's DRAM-bound streams into L2 while this row's FMA chain runs
                    const int kn = k + stride;
#pragma unroll
                    for (int d = 1; d < NDIAG; ++d) l2_prefetch(&a.A.d[d][kn]);
                    l2_prefetch(&dg[kn]); l2_prefetch(&z[kn]); l2_prefetch(&po[kn]); l2_prefetch(&a.r[kn]);
                }
                double pk = z[k] + beta * po[k];
                double acc = dg[k] * pk;
#pragma unroll
                for (int d = 1; d < NDIAG; ++d) {
                    const int o = a.A.off[d];
                    acc += a.A.d[d][k] * (z[k + o] + beta * po[k + o]);
                }
#pragma unroll
                for (int d = 1; d < NDIAG; ++d) {
                    const int o = a.A.off[d];
                    acc += a.A.d[d][k - o] * (z[k - o] + beta * po[k - o]);
                }
                pnew[k] = pk;
                a.bv[k] = acc;
                if (!DD || (own[k] & 1)) { s_pr += pk * a.r[k]; s_pb += pk * acc; }
            }
        }
        double pr, pb;
        grid_reduce3<BLOCK, CUSTOM>(grid, a.counter, epoch, s_pr, s_pb, 0.0, a.partial, sh, pr, pb, d1);
        if (DD) { double v[2] = {pr, pb}; dd_allreduce<2>(a.dd, seq_ar, v, shdd); pr = v[0]; pb = v[1]; }
        double alfa = pr / pb;
        // ---- phase B (row-block partition: the new z of my boundary rows goes straight into the neighbours' inboxes)
        double s_bz = 0.0, s_rr = 0.0;
        const int hpar = (seq_h + 1u) & 1u;
        bool sent = false;
        for (int k = t0; k < n; k += stride) {
            if (PF && k + stride < n) {
                const int kn = k + stride;
                l2_prefetch(&a.bv[kn]); l2_prefetch(&a.r[kn]); l2_prefetch(&a.x[kn]); l2_prefetch(&pnew[kn]); l2_prefetch(&dg[kn]);
            }
            double bk = a.bv[k];
            double r = a.r[k] - alfa * bk;
            a.r[k] = r;
            a.x[k] = a.x[k] + alfa * pnew[k];
            double zz = r / dg[k];
            a.z[k] = zz;
            if (!DD || (own[k] & 1)) {
                s_bz += bk * zz;
                if (!(a.cm ? dg[k] > 1.0e80 : is_dirichlet(k, a.nnod, a.ifatm, a.contp_flag))) s_rr += r * r;
            }
            if (DD && (own[k] & 6)) {
                const DDCtx &c = a.dd;
                int l = k / c.nnod, sidx = k - l * c.nnod, row = sidx / c.nc1, j = sidx - row * c.nc1;
                if ((own[k] & 2) && c.north >= 0) c.inbox_peer[c.north][((size_t)hpar * 2 + 1) * c.hcap + ((size_t)l * DD_W + (row - c.own_a)) * c.nc1 + j] = zz;
                if ((own[k] & 4) && c.south >= 0) c.inbox_peer[c.south][((size_t)hpar * 2 + 0) * c.hcap + ((size_t)l * DD_W + (row - (c.own_b - DD_W))) * c.nc1 + j] = zz;
                sent = true;
            }
        }
        if (DD && sent) __threadfence_system();
        double bz, rr;
        grid_reduce3<BLOCK, CUSTOM>(grid, a.counter, epoch, s_bz, s_rr, 0.0, a.partial, sh, bz, rr, d1);
        if (DD) {
            double v[2] = {bz, rr};
            dd_allreduce<2>(a.dd, seq_ar, v, shdd);
            bz = v[0]; rr = v[1];
            dd_recv_rows(a.dd, seq_h, a.z, t0, stride);     // publish my rows (stored in phase B), fetch the neighbours'
            grid_barrier(a.counter, epoch);
        }
        beta = -bz / pb;
        err = xlung > 0.0 ? sqrt(rr / xlung) : sqrt(rr / n);
        double *t = pold; pold = pnew; pnew = t;
        if (err > a.tol && niter < a.itmax && !(DD && *(volatile int *)a.dd.err)) { ++niter; continue; }
        break;
    }
    if (t0 == 0) {
        a.out->pcg_niter = niter; a.out->pcg_err = err; a.out->pad = (int)epoch;
        if (DD) { a.dd.seq[0] = seq_ar; a.dd.seq[1] = seq_h; }
    }
}



// ------------------------------------------------------------------------------------------
// SYMSLV with the CG vectors RESIDENT IN SHARED MEMORY (default whenever they fit: n <= #CTAs x ~9.6k rows, i.e. up to
// ~1.4 M nodes on one B200).  Every CTA owns a contiguous block of rows_cta rows for the whole solve and keeps r, p and
// B = A p (and x when there is room) of its rows in its 227 KB of shared memory; only z = M^-1 r, which the neighbours'
// stencils need, goes through global memory (L2).  The search direction is never gathered: by linearity
//     p = z + beta p_old   =>   B = A p = A z + beta B_old,
// so phase A is ONE stencil product on z (15 gathered operands per row instead of 29) and two shared-memory recurrences.
// Same recurrence otherwise (GRADDP, SRC/solscal-extended.f:1260-1380): x0 = M^-1 b, alfa = (p.r)/(p.B),
// beta = -(B.z)/(p.B), residual test on the non-Dirichlet rows; two grid barriers per iteration, fixed-order reductions.
// Per row and iteration the kernel moves 8 diagonals + z (read, write) + the diagonal again in phase B = 88 B
// (+16 B for x when it is not resident) instead of 168 B.
// ------------------------------------------------------------------------------------------
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, 1) k_pcg_res(PcgArgs a)
{
    extern __shared__ __align__(16) double smv[];
    __shared__ double sh[BLOCK / 32][2];
    __shared__ double res[2][2];
    unsigned int epoch = a.epoch0, par = 0;
    const bool PF = a.prefetch != 0;
    const int R = a.rows_cta, row0 = blockIdx.x * R, cnt = max(0, min(R, a.n - row0)), tid = threadIdx.x;
    double *rs = smv, *ps = smv + R, *bs = smv + 2 * (size_t)R, *xs = a.xres ? smv + 3 * (size_t)R : a.x + row0;
    const double *__restrict__ dg = a.diag;
    // x0 = M^-1 b ; xlung = ||b_free||^2 ; Dirichlet rows of this thread as a bit mask (row j*BLOCK + tid -> bit j)
    unsigned int dmask = 0;
    double xl = 0.0;
    for (int i = tid, j = 0; i < cnt; i += BLOCK, ++j) {
        const int k = row0 + i;
        double b = a.rhs[k];
        a.x[k] = b / dg[k];
        if (is_dirichlet(k, a.nnod, a.ifatm, a.contp_flag)) dmask |= 1u << j; else xl += b * b;
    }
    double xlung, d1;
    grid_reduce2<BLOCK>(a.counter, epoch, par, xl, 0.0, a.partial, sh, res, xlung, d1);
    // r = b - A x0 ; z = M^-1 r ; p = B = 0
    for (int i = tid; i < cnt; i += BLOCK) {
        const int k = row0 + i;
        double r = a.rhs[k] - dia_row(a.A, dg, a.x, k, a.n);
        rs[i] = r;
        a.z[k] = r / dg[k];
        ps[i] = 0.0;
        bs[i] = 0.0;
        if (a.xres) xs[i] = a.x[k];
    }
    grid_barrier(a.counter, epoch);
    double beta = 0.0, err = 0.0;
    int niter = 1;
    const double *z = a.z;      // NOT __restrict__/read-only: rewritten every iteration by other SMs
    for (;;) {
        // ---- phase A: B = A z + beta B, p = z + beta p, (p.r), (p.B)
        double s_pr = 0.0, s_pb = 0.0;
        for (int i = tid; i < cnt; i += BLOCK) {
            const int k = row0 + i;
            if (PF && i + BLOCK < cnt) {
#pragma unroll
                for (int d = 1; d < NDIAG; ++d) l2_prefetch(&a.A.d[d][k + BLOCK]);
                l2_prefetch(&dg[k + BLOCK]);
            }
            const double zk = z[k];
            double acc = dg[k] * zk;
#pragma unroll
            for (int d = 1; d < NDIAG; ++d) acc += a.A.d[d][k] * z[k + a.A.off[d]];
#pragma unroll
            for (int d = 1; d < NDIAG; ++d) acc += a.A.d[d][k - a.A.off[d]] * z[k - a.A.off[d]];
            const double pk = zk + beta * ps[i], bk = acc + beta * bs[i];
            ps[i] = pk;
            bs[i] = bk;
            s_pr += pk * rs[i];
            s_pb += pk * bk;
        }
        double pr, pb;
        grid_reduce2<BLOCK>(a.counter, epoch, par, s_pr, s_pb, a.partial, sh, res, pr, pb);
        const double alfa = pr / pb;
        // ---- phase B: r -= alfa B, x += alfa p, z = M^-1 r, (B.z), ||r_free||^2
        double s_bz = 0.0, s_rr = 0.0;
        for (int i = tid, j = 0; i < cnt; i += BLOCK, ++j) {
            const int k = row0 + i;
            const double bk = bs[i], r = rs[i] - alfa * bk;
            rs[i] = r;
            xs[i] += alfa * ps[i];
            const double zz = r / dg[k];
            a.z[k] = zz;
            s_bz += bk * zz;
            if (!((dmask >> j) & 1u)) s_rr += r * r;
        }
        double bz, rr;
        grid_reduce2<BLOCK>(a.counter, epoch, par, s_bz, s_rr, a.partial, sh, res, bz, rr);
        beta = -bz / pb;
        err = xlung > 0.0 ? sqrt(rr / xlung) : sqrt(rr / a.n);
        if (err > a.tol && niter < a.itmax) { ++niter; continue; }
        break;
    }
    if (a.xres) for (int i = tid; i < cnt; i += BLOCK) a.x[row0 + i] = xs[i];
    if (blockIdx.x == 0 && tid == 0) { a.out->pcg_niter = niter; a.out->pcg_err = err; a.out->pad = (int)epoch; }
}

// ------------------------------------------------------------------------------------------
// k_pcg_res2 (default, CATHY_PCG_ALGO=4): the resident-vector PCG above, re-cut after ncu showed k_pcg_res bound by the L1/LSU
// data pipe (l1tex__data_pipe_lsu_wavefronts 55 % of peak over the whole launch, ~90 % inside the phases) and by the grid
// reduction (tools/bench_barrier*.cu: 2.9 us each = 1400 cycles of fp64 shuffles + 870 fence + 1480 arrive/poll + 1000 re-read):
//  * every thread owns TWO consecutive rows (k0 even, k0+1) and loads 16-byte aligned pairs; the element that a misaligned
//    window lacks comes from the neighbouring lane by shuffle (edge lanes fetch it themselves).  The stencil offsets come in
//    pairs (o, o+1) -- {-1,0,1}, {NC1,NC1+1}, {NNOD-NC1-1,NNOD-NC1}, {NNOD-1,NNOD} -- so one 3-element z window serves two
//    diagonals of both rows: ~60 instead of 81 LSU wavefronts per 32 rows;
//  * M^-1 is applied as a multiplication with the reciprocal diagonal computed once per solve (no fp64 division per row);
//  * the grid reduction sums two quantities in ONE half-warp butterfly (a in lanes 0-15, b in lanes 16-31), the partials are
//    double-buffered (a fast CTA can no longer overwrite what a slow one still reads) and fetched with independent loads.
// Same recurrence and stopping test as k_pcg_res; inside a row the products are summed pair of diagonals by pair of diagonals.
// ------------------------------------------------------------------------------------------
// paired-row loads: this thread needs p[0..1] (pair) or p[0..2] (win3); lanes own consecutive pairs of rows, so lane+1 needs
// p[2..], lane-1 p[-2..].  ODD (compile time, uniform): p is 8 but not 16 bytes aligned.  The 16-byte aligned pair is loaded, the
// missing element comes from the neighbouring lane by shuffle; edge_lo / edge_hi: the lane below / above does not hold the
// continuation (lane 0 / lane 31 or the last active pair) and the element is fetched directly.  Loads (`*_ld`) and shuffles
// (`*_fin`) are separate calls so that all loads of a group are in flight before the first shuffle waits for one of them.
struct PairLd { double2 q; double e; };
template <bool ODD> __device__ __forceinline__ PairLd pair_ld(const double *p, bool edge_hi)
{
    PairLd r; r.e = 0.0;
    if (!ODD) r.q = *reinterpret_cast<const double2 *>(p);
    else { r.q = *reinterpret_cast<const double2 *>(p - 1); if (edge_hi) r.e = p[1]; }
    return r;
}
template <bool ODD> __device__ __forceinline__ void pair_fin(const PairLd &r, bool edge_hi, double &v0, double &v1)
{
    if (!ODD) { v0 = r.q.x; v1 = r.q.y; }
    else { v0 = r.q.y; const double t = __shfl_down_sync(FULLMASK, r.q.x, 1); v1 = edge_hi ? r.e : t; }
}
template <bool ODD> __device__ __forceinline__ PairLd win3_ld(const double *p, bool edge_lo, bool edge_hi)
{
    PairLd r; r.e = 0.0;
    if (!ODD) { r.q = *reinterpret_cast<const double2 *>(p); if (edge_hi) r.e = p[2]; }
    else { r.q = *reinterpret_cast<const double2 *>(p + 1); if (edge_lo) r.e = p[0]; }
    return r;
}
template <bool ODD> __device__ __forceinline__ void win3_fin(const PairLd &r, bool edge_lo, bool edge_hi, double &v0, double &v1, double &v2)
{
    if (!ODD) { v0 = r.q.x; v1 = r.q.y; const double t = __shfl_down_sync(FULLMASK, r.q.x, 1); v2 = edge_hi ? r.e : t; }
    else { v1 = r.q.x; v2 = r.q.y; const double t = __shfl_up_sync(FULLMASK, r.q.y, 1); v0 = edge_lo ? r.e : t; }
}
// one pair of diagonals (o, o+1) = (da, da+1): upper and lower products of rows k, k+1
template <bool ODD>
__device__ __forceinline__ void pair_group(const Diag &A, const double *z, int da, int o, int k, bool elo, bool ehi, double &a0, double &a1)
{
    const double2 ua = *reinterpret_cast<const double2 *>(A.d[da] + k), ub = *reinterpret_cast<const double2 *>(A.d[da + 1] + k);
    const PairLd rw = win3_ld<ODD>(z + k + o, elo, ehi), rm = win3_ld<!ODD>(z + k - o - 1, elo, ehi);
    const PairLd ra = pair_ld<ODD>(A.d[da] + k - o, ehi), rb = pair_ld<!ODD>(A.d[da + 1] + k - o - 1, ehi);   // L_d = (A_d[k - off_d], A_d[k + 1 - off_d])
    double w0, w1, w2, m0, m1, m2, la0, la1, lb0, lb1;
    win3_fin<ODD>(rw, elo, ehi, w0, w1, w2);
    win3_fin<!ODD>(rm, elo, ehi, m0, m1, m2);
    pair_fin<ODD>(ra, ehi, la0, la1);
    pair_fin<!ODD>(rb, ehi, lb0, lb1);
    a0 += ua.x * w0;  a1 += ua.y * w1;
    a0 += ub.x * w1;  a1 += ub.y * w2;
    a0 += la0 * m1;   a1 += la1 * m2;
    a0 += lb0 * m0;   a1 += lb1 * m1;
}
// CL = true: the grid is ONE thread-block cluster (meshes of a few thousand to a few ten thousand rows, e.g. BASELINE config 1 and the
// members of small-catchment ensembles): reductions and barriers are cluster-scope (cluster_reduce2), so an iteration costs ~2 us
// instead of ~7 us, and a solve occupies only its cluster's SMs -- other members' solves run beside it.
template <int BLOCK, int PAR, bool CL = false>     // PAR: parities of the offsets off[2], off[4], off[6] (bits 0, 1, 2)
__global__ void __launch_bounds__(BLOCK, 1) k_pcg_res2(PcgArgs a)
{
    extern __shared__ __align__(16) double smv[];
    __shared__ double sh[BLOCK / 32][2];
    __shared__ double res[2][2];
    __shared__ __align__(16) double cpart[2][PCG_CL_MAX][2];
    cg::cluster_group cl = cg::this_cluster();
    unsigned int epoch = a.epoch0, par = 0;
    const int R = a.rows_cta, row0 = blockIdx.x * R, cnt = max(0, min(R, a.n - row0)), tid = threadIdx.x, lane = tid & 31;
    double *rs = smv, *ps = smv + R, *bs = smv + 2 * (size_t)R, *xs = a.xres ? smv + 3 * (size_t)R : a.x + row0;
    const double *__restrict__ dg = a.diag;
    double *dinv = a.p0;          // k_pcg's search-direction buffer is free here: reciprocal diagonal
    // x0 = M^-1 b ; xlung = ||b_free||^2 ; Dirichlet rows of this thread as a bit mask (pass j: rows 2 tid + 2 BLOCK j + {0,1} -> bits 2j, 2j+1)
    unsigned int dmask = 0;
    double xl = 0.0;
    for (int i = 2 * tid, j = 0; i < cnt; i += 2 * BLOCK, ++j)
#pragma unroll
        for (int q = 0; q < 2; ++q)
            if (i + q < cnt) {
                const int k = row0 + i + q;
                const double b = a.rhs[k], dv = 1.0 / dg[k];
                dinv[k] = dv;
                a.x[k] = b * dv;
                if (is_dirichlet(k, a.nnod, a.ifatm, a.contp_flag)) dmask |= 1u << (2 * j + q); else xl += b * b;
            }
    double xlung, d1;
    if (CL) cluster_reduce2<BLOCK>(cl, par, xl, 0.0, sh, cpart, xlung, d1);
    else grid_reduce2<BLOCK>(a.counter, epoch, par, xl, 0.0, a.partial, sh, res, xlung, d1);
    // r = b - A x0 ; z = M^-1 r ; p = B = 0
    for (int i = tid; i < cnt; i += BLOCK) {
        const int k = row0 + i;
        const double r = a.rhs[k] - dia_row(a.A, dg, a.x, k, a.n);
        rs[i] = r;
        a.z[k] = r * dinv[k];
        ps[i] = 0.0;
        bs[i] = 0.0;
        if (a.xres) xs[i] = a.x[k];
    }
    if (CL) cl.sync(); else grid_barrier(a.counter, epoch);
    double beta = 0.0, err = 0.0;
    int niter = 1;
    const double *z = a.z;      // NOT __restrict__/read-only: rewritten every iteration by other SMs
    const int o2 = a.A.off[2], o4 = a.A.off[4], o6 = a.A.off[6];      // off[1] = 1, off[3] = o2 + 1, off[5] = o4 + 1, off[7] = o6 + 1 (checked by the host)
    const int last = (cnt - 1) & ~1;                                   // first row of the last pair
    const int iwarp_end = cnt;                                         // a warp runs a pass while its first pair exists
    for (;;) {
        // ---- phase A: B = A z + beta B, p = z + beta p, (p.r), (p.B)
        double s_pr = 0.0, s_pb = 0.0;
        for (int iw = 2 * (tid - lane); iw < iwarp_end; iw += 2 * BLOCK) {
            const int i_own = iw + 2 * lane;
            const bool act = i_own < cnt, ok1 = i_own + 1 < cnt;
            const int i = act ? i_own : last;                          // idle lanes of the last warp shadow the last pair (their shuffles feed nobody)
            const bool ehi = lane == 31 || i_own + 2 >= cnt, elo = lane == 0;
            const int k = row0 + i;
            // centre window z[k-1..k+2]: (z[k], z[k+1]) is the aligned pair
            const double2 zc = *reinterpret_cast<const double2 *>(z + k);
            double zm = __shfl_up_sync(FULLMASK, zc.y, 1), zp = __shfl_down_sync(FULLMASK, zc.x, 1);
            if (elo) zm = z[k - 1];
            if (ehi) zp = z[k + 2];
            const double2 dd = *reinterpret_cast<const double2 *>(dg + k);
            const double2 u1 = *reinterpret_cast<const double2 *>(a.A.d[1] + k);
            // lower part of diagonal 1: A1[k-1] (from the lane below), A1[k] = u1.x
            double l1 = __shfl_up_sync(FULLMASK, u1.y, 1);
            if (elo) l1 = a.A.d[1][k - 1];
            double a0 = dd.x * zc.x, a1 = dd.y * zc.y;
            a0 += u1.x * zc.y;  a1 += u1.y * zp;
            a0 += l1 * zm;      a1 += u1.x * zc.x;
            // the three offset pairs (o, o+1), one after the other (keeps the live registers under the 64 a 1024-thread CTA gets)
            pair_group<(PAR & 1) != 0>(a.A, z, 2, o2, k, elo, ehi, a0, a1);
            pair_group<(PAR & 2) != 0>(a.A, z, 4, o4, k, elo, ehi, a0, a1);
            pair_group<(PAR & 4) != 0>(a.A, z, 6, o6, k, elo, ehi, a0, a1);
            if (act) {
                double2 pv = *reinterpret_cast<double2 *>(ps + i), bv = *reinterpret_cast<double2 *>(bs + i);
                const double2 rv = *reinterpret_cast<const double2 *>(rs + i);
                pv.x = zc.x + beta * pv.x; pv.y = zc.y + beta * pv.y;
                bv.x = a0 + beta * bv.x;   bv.y = a1 + beta * bv.y;
                s_pr += pv.x * rv.x; s_pb += pv.x * bv.x;
                if (ok1) {
                    s_pr += pv.y * rv.y; s_pb += pv.y * bv.y;
                    *reinterpret_cast<double2 *>(ps + i) = pv; *reinterpret_cast<double2 *>(bs + i) = bv;
                } else { ps[i] = pv.x; bs[i] = bv.x; }
            }
        }
        double pr, pb;
        if (CL) cluster_reduce2<BLOCK>(cl, par, s_pr, s_pb, sh, cpart, pr, pb);
        else grid_reduce2<BLOCK>(a.counter, epoch, par, s_pr, s_pb, a.partial, sh, res, pr, pb);
        const double alfa = pr / pb;
        // ---- phase B: r -= alfa B, x += alfa p, z = M^-1 r, (B.z), ||r_free||^2
        double s_bz = 0.0, s_rr = 0.0;
        for (int i = 2 * tid, j = 0; i < cnt; i += 2 * BLOCK, ++j) {
            const int k = row0 + i;
            const bool ok1 = i + 1 < cnt;
            const double2 bv = *reinterpret_cast<const double2 *>(bs + i), pv = *reinterpret_cast<const double2 *>(ps + i);
            double2 rv = *reinterpret_cast<double2 *>(rs + i);
            const double2 dv = *reinterpret_cast<const double2 *>(dinv + k);
            rv.x -= alfa * bv.x; rv.y -= alfa * bv.y;
            double2 zz; zz.x = rv.x * dv.x; zz.y = rv.y * dv.y;
            s_bz += bv.x * zz.x;
            if (!((dmask >> (2 * j)) & 1u)) s_rr += rv.x * rv.x;
            if (ok1) {
                double2 xv = *reinterpret_cast<double2 *>(xs + i);
                xv.x += alfa * pv.x; xv.y += alfa * pv.y;
                *reinterpret_cast<double2 *>(xs + i) = xv;
                *reinterpret_cast<double2 *>(rs + i) = rv;
                *reinterpret_cast<double2 *>(a.z + k) = zz;
                s_bz += bv.y * zz.y;
                if (!((dmask >> (2 * j + 1)) & 1u)) s_rr += rv.y * rv.y;
            } else { xs[i] += alfa * pv.x; rs[i] = rv.x; a.z[k] = zz.x; }
        }
        double bz, rr;
        if (CL) cluster_reduce2<BLOCK>(cl, par, s_bz, s_rr, sh, cpart, bz, rr);
        else grid_reduce2<BLOCK>(a.counter, epoch, par, s_bz, s_rr, a.partial, sh, res, bz, rr);
        beta = -bz / pb;
        err = xlung > 0.0 ? sqrt(rr / xlung) : sqrt(rr / a.n);
        if (err > a.tol && niter < a.itmax) { ++niter; continue; }
        break;
    }
    if (a.xres) for (int i = tid; i < cnt; i += BLOCK) a.x[row0 + i] = xs[i];
    if (blockIdx.x == 0 && tid == 0) { a.out->pcg_niter = niter; a.out->pcg_err = err; a.out->pad = (int)epoch; }
}

#include "pcg_cluster.cuh"

// ------------------------------------------------------------------------------------------
// SYMSLV, second formulation (opt-in, CATHY_PCG_ALGO=2; measured slower than k_pcg on B200 except on tiny meshes, see
// profiles/r1_pcg_experiments.md): the system is scaled symmetrically, As = D^-1/2 A D^-1/2 (unit
// diagonal, y = D^1/2 x), so that the Jacobi-preconditioned CG of k_pcg becomes plain CG without the z vector and without the
// diagonal; and the recurrence is the single-reduction form of CG (Chronopoulos & Gear): with w = As r,
//     gamma = (r,r), delta = (w,r);  beta = gamma/gamma_old;  alpha = gamma / (delta - beta*gamma/alpha_old)
//     p = r + beta p;  s = w + beta s;  y += alpha p;  r -= alpha s;  w = As r
// The new w needs the new r of the 14 neighbours, which every thread recomputes on the fly from the OLD r, w, s
// (r_j - alpha (w_j + beta s_j)); r, w, s are double-buffered.  ONE grid-wide barrier per iteration (inside the reduction)
// instead of two, 144 instead of 168 bytes per row and iteration.  Same iterates as SYMSLV/GRADDP in exact arithmetic (same
// x0 = M^-1 b, same stopping test on the unscaled residual, Dirichlet rows excluded).
// ------------------------------------------------------------------------------------------
__global__ void k_sym_scale(int n, Diag A, const double *__restrict__ diag_bc, double *__restrict__ dis)
{   // pass 1: dis = 1/sqrt(diag)
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) dis[k] = 1.0 / sqrt(diag_bc[k]);
}
__global__ void k_sym_scale2(int n, Diag A, const double *__restrict__ dis)
{   // pass 2: off-diagonals in place (dis carries a halo; the entries that reach into it are structurally zero)
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const double dk = dis[k];
#pragma unroll
        for (int d = 1; d < NDIAG; ++d) A.d[d][k] = (A.d[d][k] * dk) * dis[k + A.off[d]];
    }
}
struct Pcg2Args {
    int n, nnod, itmax, prefetch;
    double tol;
    Diag A;                  // scaled off-diagonals in d[1..7]
    const double *dis;       // 1/sqrt(diagonal with the Dirichlet penalty)
    const double *rhs;
    double *y, *p, *r0, *r1, *w0, *w1, *s0, *s1;
    const int *ifatm;
    const unsigned char *contp_flag;
    double *partial;
    unsigned int *counter;
    unsigned int epoch0;
    IterOut *out;
};
__device__ __forceinline__ double dia_offrow(const Diag &A, const double *x, int k)
{   // sum over the 14 off-diagonal entries of row k (unit diagonal not included)
    double acc = 0.0;
#pragma unroll
    for (int d = 1; d < NDIAG; ++d) acc += A.d[d][k] * x[k + A.off[d]];
#pragma unroll
    for (int d = 1; d < NDIAG; ++d) acc += A.d[d][k - A.off[d]] * x[k - A.off[d]];
    return acc;
}
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, 1024 / BLOCK) k_pcg2(Pcg2Args a)
{
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[BLOCK / 32][3];
    unsigned int epoch = a.epoch0;
    const int n = a.n, stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    const double *__restrict__ dis = a.dis;
    const bool PF = a.prefetch != 0;
    // y0 = D^1/2 x0 = b/sqrt(d) (x0 = M^-1 b, :4686);  xlung = ||b_free||^2 (:1286-1297)
    double xl = 0.0;
    for (int k = t0; k < n; k += stride) {
        double b = a.rhs[k];
        a.y[k] = b * dis[k];
        a.p[k] = 0.0; a.s0[k] = 0.0;
        if (!is_dirichlet(k, a.nnod, a.ifatm, a.contp_flag)) xl += b * b;
    }
    double xlung, g0, d0;
    grid_reduce3<BLOCK, true>(grid, a.counter, epoch, xl, 0.0, 0.0, a.partial, sh, xlung, g0, d0);
    // r = b~ - As y0
    for (int k = t0; k < n; k += stride) a.r0[k] = a.rhs[k] * dis[k] - (a.y[k] + dia_offrow(a.A, a.y, k));
    grid_barrier(a.counter, epoch);
    // w = As r ; gamma = (r,r) ; delta = (w,r)
    double sg = 0.0, sd = 0.0;
    for (int k = t0; k < n; k += stride) {
        double r = a.r0[k], w = r + dia_offrow(a.A, a.r0, k);
        a.w0[k] = w;
        sg += r * r; sd += w * r;
    }
    double gamma, delta, rr;
    grid_reduce3<BLOCK, true>(grid, a.counter, epoch, sg, sd, 0.0, a.partial, sh, gamma, delta, rr);
    double alpha = gamma / delta, beta = 0.0, err = 0.0;
    double *rc = a.r0, *rn = a.r1, *wc = a.w0, *wn = a.w1, *sc = a.s0, *sn = a.s1;
    int niter = 1;
    for (;;) {
        const double ab = alpha * beta;
        double s_g = 0.0, s_d = 0.0, s_rr = 0.0;
        for (int k = t0; k < n; k += stride) {
            if (PF && k + stride < n) {
                const int kn = k + stride;
#pragma unroll
                for (int d = 1; d < NDIAG; ++d) l2_prefetch(&a.A.d[d][kn]);
                l2_prefetch(&rc[kn]); l2_prefetch(&wc[kn]); l2_prefetch(&sc[kn]); l2_prefetch(&a.p[kn]); l2_prefetch(&a.y[kn]); l2_prefetch(&dis[kn]);
            }
            const double r = rc[k], w = wc[k], so = sc[k];
            const double s = w + beta * so;
            const double p = r + beta * a.p[k];
            const double r2 = (r - alpha * w) - ab * so;      // = r - alpha s, in the very form the neighbours use below
            double acc = r2;                                 // unit diagonal
#pragma unroll
            for (int d = 1; d < NDIAG; ++d) {
                const int j = k + a.A.off[d];
                acc += a.A.d[d][k] * ((rc[j] - alpha * wc[j]) - ab * sc[j]);
            }
#pragma unroll
            for (int d = 1; d < NDIAG; ++d) {
                const int j = k - a.A.off[d];
                acc += a.A.d[d][j] * ((rc[j] - alpha * wc[j]) - ab * sc[j]);
            }
            sn[k] = s; a.p[k] = p; a.y[k] = a.y[k] + alpha * p; rn[k] = r2; wn[k] = acc;
            s_g += r2 * r2; s_d += acc * r2;
            if (!is_dirichlet(k, a.nnod, a.ifatm, a.contp_flag)) { double di = dis[k]; s_rr += (r2 * r2) / (di * di); }   // unscaled residual
        }
        double g1, d1;
        grid_reduce3<BLOCK, true>(grid, a.counter, epoch, s_g, s_d, s_rr, a.partial, sh, g1, d1, rr);
        err = xlung > 0.0 ? sqrt(rr / xlung) : sqrt(rr / n);
        double *t;
        t = rc; rc = rn; rn = t; t = wc; wc = wn; wn = t; t = sc; sc = sn; sn = t;
        if (err > a.tol && niter < a.itmax) {
            beta = g1 / gamma;
            alpha = g1 / (d1 - beta * g1 / alpha);
            gamma = g1;
            ++niter;
            continue;
        }
        break;
    }
    // x = D^-1/2 y
    for (int k = t0; k < n; k += stride) a.y[k] = a.y[k] * dis[k];
    if (t0 == 0) { a.out->pcg_niter = niter; a.out->pcg_err = err; a.out->pad = (int)epoch; }
}

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

#include "bicg_res.cuh"
#include "pcg_tma.cuh"

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

// ------------------------------------------------------------------------------------------
// surface routing (SURF_FLOWTRA, SRC/surf_flowtra.f:38-196)
// ------------------------------------------------------------------------------------------
// NOD_CELL + TRANSFER_F3D_SURF (SRC/nod_cell.f, SRC/transfer_f3d_surf.f); OVFLNOD is divided by the
// nodal area IN PLACE first (separate launch), exactly as the reference does.
__global__ void k_div_area(int nnod, const double *__restrict__ arenod, double *__restrict__ ovfl)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nnod; i += gridDim.x * blockDim.x) ovfl[i] = ovfl[i] / arenod[i];
}
__global__ void k_nod_cell(int nrow, int ncol, double dx, double dy, const double *__restrict__ ovfl, double *__restrict__ sw_sn)
{
    int ncell = nrow * ncol, nc1 = ncol + 1;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncell; c += gridDim.x * blockDim.x) {
        int i = c / ncol, j = c - i * ncol;
        int n00 = i * nc1 + j, n10 = n00 + nc1, n11 = n10 + 1, n01 = n00 + 1;
        double cc = 0.0;
        cc = cc + ovfl[n00]; cc = cc + ovfl[n10]; cc = cc + ovfl[n11]; cc = cc + ovfl[n01];
        cc = cc * 0.25;
        int jr = nrow - 1 - i;                 // row counted from the south
        sw_sn[j * nrow + jr] = cc * dx * dy;   // routing index (I-1)*NROW+J
    }
}
// CELL_NOD + TRANSFER_SURF_F3D (SRC/cell_nod.f, SRC/transfer_surf_f3d.f): ponding head per node =
// mean over the adjacent triangles, accumulated in triangle order
__global__ void k_cell_nod(int nrow, int ncol, const double *__restrict__ h_sn, double *__restrict__ pondnod)
{
    int nc1 = ncol + 1, nnod = (nrow + 1) * nc1;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nnod; s += gridDim.x * blockDim.x) {
        int i = s / nc1, j = s - i * nc1;
        double acc = 0.0;
        int cnt = 0;
        auto cellv = [&](int ci, int cj) { return h_sn[cj * nrow + (nrow - 1 - ci)]; };
        if (i > 0 && j > 0) { double v = cellv(i - 1, j - 1); acc = acc + v; acc = acc + v; cnt += 2; }
        if (i > 0 && j < ncol) { acc = acc + cellv(i - 1, j); cnt += 1; }
        if (i < nrow && j > 0) { acc = acc + cellv(i, j - 1); cnt += 1; }
        if (i < nrow && j < ncol) { double v = cellv(i, j); acc = acc + v; acc = acc + v; cnt += 2; }
        pondnod[s] = acc / cnt;
    }
}

struct RouteArgs {
    int ncell, nlevel;
    const int *level_ptr;     // [nlevel+1] cells grouped by drainage level (level-scheduled tree)
    const int *level_cell;    // [ncell] routing index I_BASIN (0-based)
    const int *seq;           // [ncell] position of the cell in QOI order (for the AK_MAX tie rule)
    const int *don_ptr;       // [ncell+1] donors of each cell in QOI order
    const int *don_cell;      // donor routing index
    const unsigned char *don_dir; // 0: donor's direction-1 outflow, 1: direction-2
    const int *don_code;      // per donor entry: (index << 3) | (direction << 2) | kind, see k_route
    const double *w1, *w2, *sl1, *sl2, *epl1, *epl2, *ks1, *ks2, *ws1, *ws2, *b1, *y1, *nrc;
    double *sw_sn, *q_in_kk, *q_in_kkp1, *q_out_kk_1, *q_out_kk_2, *q_out_kkp1_1, *q_out_kkp1_2;
    double *volume_kk, *volume_kkp1, *h_water;
    double *ak_max;           // in/out
    int *nsurf_out;
    double deltat, cellarea;
    double *ckf1, *ckf2, *dhd1, *dhd2;   // static factors of MC per cell and direction (k_route_static)
};
// Muskingum-Cunge for one cell and direction (MC, SRC/mc.f).  Of the kinematic celerity
//   CK = 5/(3 G) KS^(3/5) W^(-2/5) sin(BETA)^(3/10) QC^(1 - 3G/5)   and   DH = QC^(1 - B1) / (2 G W tan(BETA))
// only the powers of QC change during a run: the leading product `ckf` and the denominator `dhd` are evaluated once per cell and
// direction by k_route_static with the same operations in the same order (products associate left to right), so hoisting them
// leaves every result bit-identical and takes 3 of the 5 pow() calls, atan, sin and tan out of each cell's dependent chain --
// the routing runs level by level on ONE SM, where this chain is the critical path.
__device__ __forceinline__ void mc_static(double slope, double ks, double w, double b1, double y1, double &ckf, double &dhd)
{
    double beta = atan(slope);
    double g = (1.0 - y1 + 2.0 / 3.0 * b1);
    ckf = 5.0 / (3.0 * g) * pow(ks, 3.0 / 5.0) * pow(w, -2.0 / 5.0) * pow(sin(beta), 3.0 / 1.0e1);
    dhd = 2 * g * w * tan(beta);
}
__device__ __forceinline__ double mc_cell(double ckf, double dhd, double epl, double b1, double y1, double dt,
                                          double q_in_kk, double q_in_kkp1, double q_out_kk, double q_over, double &cu, double &ak)
{
    double qc = 1.0 / 3.0 * (q_in_kk + q_in_kkp1 + q_out_kk);
    if (qc <= 1.0e-05) qc = 1.0e-05;
    double g = (1.0 - y1 + 2.0 / 3.0 * b1);
    double ck = ckf * pow(qc, 1.0 - 3.0 * g / 5.0);
    ak = ck / epl;
    cu = ck * dt / epl;
    double dh = pow(qc, 1.0 - b1) / dhd;
    if (dh < (1.0 - cu)) dh = 1.0 - cu;
    double xx = 0.50 - dh / (ck * epl);
    double den = 2.0 * (1.0 - xx) + cu;
    double c1 = (cu - 2.0 * xx) / den, c2 = (cu + 2.0 * xx) / den, c3 = (2.0 * (1.0 - xx) - cu) / den, c4 = (2.0 * ck * dt) / den;
    return c1 * q_in_kkp1 + c2 * q_in_kk + c3 * q_out_kk + c4 * q_over;
}
__global__ void k_route_static(RouteArgs a)
{
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < a.ncell; c += gridDim.x * blockDim.x) {
        mc_static(a.sl1[c], a.ks1[c], a.ws1[c], a.b1[c], a.y1[c], a.ckf1[c], a.dhd1[c]);
        mc_static(a.sl2[c], a.ks2[c], a.ws2[c], a.b1[c], a.y1[c], a.ckf2[c], a.dhd2[c]);
    }
}
// All NSURF sub-steps of ROUTE + ALTEZZE (SRC/route.f:47-253, SRC/altezze.f) in ONE launch of one CTA:
// cells are processed level by level down the drainage tree (a cell's inflow is the ordered sum of its
// donors' outflows, so the result equals the reference's sequential descending-elevation sweep).
// The levels are short (a few hundred cells) and strictly dependent, so the time per level is the latency of one thread's chain
// level_cell -> don_ptr -> don_cell -> donor outflow -> MC.  Everything in that chain that does not depend on the previous level
// (indices, donor lists, the cell's parameters and old-time-level values) is loaded one level AHEAD, while the current level
// computes: after the barrier only the donors' outflows remain to be fetched.
// Donor kinds (don_code & 3): 0 = the donor sits at least two levels up: its outflow is final when the loads of the next level are
// issued, so it is fetched ahead with everything else; 1 = the donor was computed in the level just finished by the thread whose
// slot is the index: its outflow is read from the CTA's shared stash (a few cycles instead of an L2 round trip on the critical
// path); 2 = previous level but beyond the stash (levels wider than the CTA): read from global memory after the barrier.
constexpr int ROUTE_BLOCK = 384, ROUTE_RD = 4;
struct RouteCell {
    int ib, d0, nd, seq;
    int dc[ROUTE_RD];        // don_code of the first ROUTE_RD donors
    double dq[ROUTE_RD];     // outflows of the kind-0 donors among them
    double w[2], epl[2], ckf[2], dhd[2], qok[2], nrc, b1, y1, sw, qik;
};
__device__ __forceinline__ void route_load(const RouteArgs &a, int q, RouteCell &c)
{
    const int ib = a.level_cell[q];
    c.ib = ib; c.seq = a.seq[ib];
    c.d0 = a.don_ptr[ib]; c.nd = a.don_ptr[ib + 1] - c.d0;
#pragma unroll
    for (int j = 0; j < ROUTE_RD; ++j) {
        const int code = j < c.nd ? a.don_code[c.d0 + j] : 1;
        c.dc[j] = code;
        c.dq[j] = (code & 3) == 0 ? ((code & 4) ? a.q_out_kkp1_2[code >> 3] : a.q_out_kkp1_1[code >> 3]) : 0.0;
    }
    c.w[0] = a.w1[ib]; c.w[1] = a.w2[ib]; c.epl[0] = a.epl1[ib]; c.epl[1] = a.epl2[ib];
    c.ckf[0] = a.ckf1[ib]; c.ckf[1] = a.ckf2[ib]; c.dhd[0] = a.dhd1[ib]; c.dhd[1] = a.dhd2[ib];
    c.qok[0] = a.q_out_kk_1[ib]; c.qok[1] = a.q_out_kk_2[ib];
    c.nrc = a.nrc[ib]; c.b1 = a.b1[ib]; c.y1 = a.y1[ib]; c.sw = a.sw_sn[ib]; c.qik = a.q_in_kk[ib];
}
// prev: stash written by the previous level, mine: this level's stash, slot: this thread's stash slot or -1
__device__ __forceinline__ void route_cell(const RouteArgs &a, const RouteCell &c, double dt, const double (*prev)[ROUTE_BLOCK],
                                           double (*mine)[ROUTE_BLOCK], int slot, double &best_cu, double &best_ak, int &best_seq)
{
    const int ib = c.ib;
    double qin = 0.0;
#pragma unroll
    for (int j = 0; j < ROUTE_RD; ++j)
        if (j < c.nd) {
            const int code = c.dc[j], kind = code & 3, idx = code >> 3, dr = (code >> 2) & 1;
            const double v = kind == 0 ? c.dq[j] : kind == 1 ? prev[dr][idx] : (dr ? a.q_out_kkp1_2[idx] : a.q_out_kkp1_1[idx]);
            qin = qin + v;
        }
    for (int dn = c.d0 + ROUTE_RD; dn < c.d0 + c.nd; ++dn)
        qin = qin + (a.don_dir[dn] ? a.q_out_kkp1_2[a.don_cell[dn]] : a.q_out_kkp1_1[a.don_cell[dn]]);
    a.q_in_kkp1[ib] = qin;
    const double nrc = c.nrc, swv = c.sw / nrc;
#pragma unroll
    for (int dir = 0; dir < 2; ++dir) {
        const double w = c.w[dir];
        double *qo_kkp1 = dir ? a.q_out_kkp1_2 : a.q_out_kkp1_1;
        if (w == 0.0) continue;
        const double epl = c.epl[dir];
        double q_over = swv * w * (1.0 / epl);
        double q_in_kk = c.qik * w / nrc, q_out_kk = c.qok[dir] / nrc;
        double q_in_kkp1 = qin * w / nrc, cu, ak;
        double qo = mc_cell(c.ckf[dir], c.dhd[dir], epl, c.b1, c.y1, dt, q_in_kk, q_in_kkp1, q_out_kk, q_over, cu, ak);
        if (qo < 0.0) qo = 0.0;
        qo_kkp1[ib] = qo * nrc;
        if (slot >= 0) mine[dir][slot] = qo * nrc;
        int sq = 2 * c.seq + dir;
        if (cu > best_cu || (cu == best_cu && sq > best_seq)) { best_cu = cu; best_ak = ak; best_seq = sq; }
    }
}
__global__ void __launch_bounds__(ROUTE_BLOCK) k_route(RouteArgs a, const int *handled)
{
    if (handled && *handled) return;             // k_route_wave (route_wave.cuh) did this step
    __shared__ double s_cu[32], s_ak[32];
    __shared__ int s_seq[32];
    __shared__ double s_akmax;
    __shared__ int s_nsurf;
    __shared__ double s_dt;
    __shared__ double s_q[2][2][ROUTE_BLOCK];      // [level parity][direction][slot]: outflows of the level's first ROUTE_BLOCK cells
    if (threadIdx.x == 0) {
        double akm = *a.ak_max, cu_max = akm * a.deltat, dts;
        int ns;
        if (cu_max > 1.0) { dts = 1.0 / akm; ns = (int)(a.deltat / dts) + 1; dts = a.deltat / ns; }
        else { dts = a.deltat; ns = 1; }
        s_nsurf = ns; s_dt = dts; s_akmax = akm;
    }
    __syncthreads();
    const int nsurf = s_nsurf;
    const double dt = s_dt;
    const int *__restrict__ lp = a.level_ptr;
    for (int sub = 1; sub <= nsurf; ++sub) {
        double best_cu = -1.0, best_ak = 0.0;
        int best_seq = -1;
        RouteCell nxt;
        bool have = (int)threadIdx.x < lp[1] - lp[0];
        if (have) route_load(a, lp[0] + threadIdx.x, nxt);
        for (int lv = 0; lv < a.nlevel; ++lv) {
            const int beg = lp[lv], end = lp[lv + 1];
            const RouteCell cur = nxt;
            const bool hc = have;
            have = false;
            if (lv + 1 < a.nlevel) {       // the next level's first cell of this thread: nothing here depends on this level's results
                const int q = end + threadIdx.x;
                have = q < lp[lv + 2];
                if (have) route_load(a, q, nxt);
            }
            if (hc) route_cell(a, cur, dt, s_q[(lv + 1) & 1], s_q[lv & 1], (int)threadIdx.x, best_cu, best_ak, best_seq);
            for (int q = beg + threadIdx.x + blockDim.x; q < end; q += blockDim.x) {
                RouteCell t;
                route_load(a, q, t);
                route_cell(a, t, dt, s_q[(lv + 1) & 1], s_q[lv & 1], -1, best_cu, best_ak, best_seq);
            }
            __syncthreads();
        }
        // AK_MAX = celerity of the LAST cell (in sequential order) attaining the max Courant number
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double oc = __shfl_down_sync(0xffffffffu, best_cu, o), oa = __shfl_down_sync(0xffffffffu, best_ak, o);
            int os = __shfl_down_sync(0xffffffffu, best_seq, o);
            if (oc > best_cu || (oc == best_cu && os > best_seq)) { best_cu = oc; best_ak = oa; best_seq = os; }
        }
        int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        if (lane == 0) { s_cu[wid] = best_cu; s_ak[wid] = best_ak; s_seq[wid] = best_seq; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int q = 1; q < (int)(blockDim.x >> 5); ++q)
                if (s_cu[q] > best_cu || (s_cu[q] == best_cu && s_seq[q] > best_seq)) { best_cu = s_cu[q]; best_ak = s_ak[q]; best_seq = s_seq[q]; }
            if (best_seq >= 0) s_akmax = best_ak;
        }
        // ALTEZZE: volume balance and water depth per cell
        for (int c = threadIdx.x; c < a.ncell; c += blockDim.x) {
            double dv = (a.q_in_kk[c] + a.q_in_kkp1[c]) / 2 * dt + a.sw_sn[c] * dt - (a.q_out_kk_1[c] + a.q_out_kk_2[c]) / 2 * dt
                        - (a.q_out_kkp1_1[c] + a.q_out_kkp1_2[c]) / 2 * dt;
            double v1 = a.volume_kk[c] + dv, h;
            if (v1 >= 0.0) h = v1 / a.cellarea; else { v1 = 0.0; h = 0.0; }
            a.volume_kkp1[c] = v1;
            a.h_water[c] = h;
            if (nsurf > 1 && sub < nsurf) {      // shift time levels for the next sub-step (:171-195)
                a.q_in_kk[c] = a.q_in_kkp1[c]; a.q_in_kkp1[c] = 0.0;
                a.q_out_kk_1[c] = a.q_out_kkp1_1[c]; a.q_out_kkp1_1[c] = 0.0;
                a.q_out_kk_2[c] = a.q_out_kkp1_2[c]; a.q_out_kkp1_2[c] = 0.0;
                a.volume_kk[c] = v1; a.volume_kkp1[c] = 0.0;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { *a.ak_max = s_akmax; *a.nsurf_out = nsurf; }
}

#include "route_wave.cuh"

// k_route with FOUR lanes per cell: the four fp64 pow() of a cell (two directions x celerity / diffusivity) are dependent chains of
// ~1 us each and bound the time per drainage level when one thread evaluates them one after the other (5.3 us per level, 2.1 ms per
// call on the 200x200 bench DEM = a third of the coupled step).  Here lane q of a quad evaluates power q of its cell, the partner
// lane's power arrives by shuffle, and the even lanes finish their direction with mc_finish (route_wave.cuh) -- operation for operation
// the arithmetic of mc_cell.  Everything else is k_route: one CTA, level after level, next level's records fetched ahead, outflows of
// the level just finished read from the shared stash.
// MEASURED NEGATIVE (profiles/micro/r2f_route4.log): the coupled bench workload goes from 6.65 to 7.86 ms per step -- the compiler
// already interleaves the four independent pow() chains of one thread, and with 128 cells per pass the 200-cell levels of the bench DEM
// need two passes.  Kept as an opt-in (CATHY_ROUTE_LANES=4) with its parity tests green; k_route stays the default.
constexpr int ROUTE4_BLOCK = 512, ROUTE4_CELLS = ROUTE4_BLOCK / 4;
__global__ void __launch_bounds__(ROUTE4_BLOCK) k_route4(RouteArgs a, const int *handled)
{
    if (handled && *handled) return;             // k_route_wave (route_wave.cuh) did this step
    __shared__ double s_cu[32], s_ak[32];
    __shared__ int s_seq[32];
    __shared__ double s_akmax;
    __shared__ int s_nsurf;
    __shared__ double s_dt;
    __shared__ double s_q[2][2][ROUTE_BLOCK];      // [level parity][direction][slot]: outflows of the level's first ROUTE_BLOCK cells
    if (threadIdx.x == 0) {
        double akm = *a.ak_max, cu_max = akm * a.deltat, dts;
        int ns;
        if (cu_max > 1.0) { dts = 1.0 / akm; ns = (int)(a.deltat / dts) + 1; dts = a.deltat / ns; }
        else { dts = a.deltat; ns = 1; }
        s_nsurf = ns; s_dt = dts; s_akmax = akm;
    }
    __syncthreads();
    const int nsurf = s_nsurf;
    const double dt = s_dt;
    const int *__restrict__ lp = a.level_ptr;
    const int cell = threadIdx.x >> 2, quad = threadIdx.x & 3, dir = quad >> 1;
    // one cell of a level: all four lanes hold the record (same addresses: one transaction), lane q raises the reference discharge of
    // direction q / 2 to the celerity (q even) or diffusivity (q odd) exponent
    auto do_cell = [&](const RouteCell &c, bool act, const double (*prev)[ROUTE_BLOCK], double (*mine)[ROUTE_BLOCK], int slot,
                       double &best_cu, double &best_ak, int &best_seq) {
        double qin = 0.0;
        if (act) {
#pragma unroll
            for (int j = 0; j < ROUTE_RD; ++j)
                if (j < c.nd) {
                    const int code = c.dc[j], kind = code & 3, idx = code >> 3, dr = (code >> 2) & 1;
                    const double v = kind == 0 ? c.dq[j] : kind == 1 ? prev[dr][idx] : (dr ? a.q_out_kkp1_2[idx] : a.q_out_kkp1_1[idx]);
                    qin = qin + v;
                }
            for (int dn = c.d0 + ROUTE_RD; dn < c.d0 + c.nd; ++dn)
                qin = qin + (a.don_dir[dn] ? a.q_out_kkp1_2[a.don_cell[dn]] : a.q_out_kkp1_1[a.don_cell[dn]]);
            if (quad == 0) a.q_in_kkp1[c.ib] = qin;
        }
        const double nrc = act ? c.nrc : 1.0, w = act ? c.w[dir] : 0.0, epl = act ? c.epl[dir] : 1.0;
        const bool on = act && w != 0.0;
        const double q_in_kk = c.qik * w / nrc, q_out_kk = c.qok[dir] / nrc, q_in_kkp1 = qin * w / nrc;
        double pw = 0.0;
        if (on) {
            const double qc = mc_qc(q_in_kk, q_in_kkp1, q_out_kk), g = (1.0 - c.y1 + 2.0 / 3.0 * c.b1);
            pw = pow(qc, (quad & 1) ? 1.0 - c.b1 : 1.0 - 3.0 * g / 5.0);
        }
        const double p_dh = __shfl_xor_sync(0xffffffffu, pw, 1);
        if (on && !(quad & 1)) {
            const double swv = c.sw / nrc, q_over = swv * w * (1.0 / epl);
            double cu, ak;
            double qo = mc_finish(c.ckf[dir], c.dhd[dir], epl, dt, pw, p_dh, q_in_kk, q_in_kkp1, q_out_kk, q_over, cu, ak);
            if (qo < 0.0) qo = 0.0;
            (dir ? a.q_out_kkp1_2 : a.q_out_kkp1_1)[c.ib] = qo * nrc;
            if (slot >= 0) mine[dir][slot] = qo * nrc;
            const int sq = 2 * c.seq + dir;
            if (cu > best_cu || (cu == best_cu && sq > best_seq)) { best_cu = cu; best_ak = ak; best_seq = sq; }
        }
    };
    for (int sub = 1; sub <= nsurf; ++sub) {
        double best_cu = -1.0, best_ak = 0.0;
        int best_seq = -1;
        RouteCell nxt;
        bool have = cell < lp[1] - lp[0];
        if (have) route_load(a, lp[0] + cell, nxt);
        for (int lv = 0; lv < a.nlevel; ++lv) {
            const int beg = lp[lv], end = lp[lv + 1];
            const RouteCell cur = nxt;
            const bool hc = have;
            have = false;
            if (lv + 1 < a.nlevel) {       // the next level's first cell of this quad: nothing here depends on this level's results
                const int q = end + cell;
                have = q < lp[lv + 2];
                if (have) route_load(a, q, nxt);
            }
            do_cell(cur, hc, s_q[(lv + 1) & 1], s_q[lv & 1], cell, best_cu, best_ak, best_seq);
            for (int q0 = beg + ROUTE4_CELLS; q0 < end; q0 += ROUTE4_CELLS) {      // levels wider than one pass (warp-uniform trip count)
                const int q = q0 + cell;
                const bool act = q < end;
                RouteCell t;
                if (act) route_load(a, q, t);
                do_cell(t, act, s_q[(lv + 1) & 1], s_q[lv & 1], (act && q - beg < ROUTE_BLOCK) ? q - beg : -1, best_cu, best_ak, best_seq);
            }
            __syncthreads();
        }
        // AK_MAX = celerity of the LAST cell (in sequential order) attaining the max Courant number
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double oc = __shfl_down_sync(0xffffffffu, best_cu, o), oa = __shfl_down_sync(0xffffffffu, best_ak, o);
            int os = __shfl_down_sync(0xffffffffu, best_seq, o);
            if (oc > best_cu || (oc == best_cu && os > best_seq)) { best_cu = oc; best_ak = oa; best_seq = os; }
        }
        int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        if (lane == 0) { s_cu[wid] = best_cu; s_ak[wid] = best_ak; s_seq[wid] = best_seq; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int q = 1; q < (int)(blockDim.x >> 5); ++q)
                if (s_cu[q] > best_cu || (s_cu[q] == best_cu && s_seq[q] > best_seq)) { best_cu = s_cu[q]; best_ak = s_ak[q]; best_seq = s_seq[q]; }
            if (best_seq >= 0) s_akmax = best_ak;
        }
        // ALTEZZE: volume balance and water depth per cell
        for (int c = threadIdx.x; c < a.ncell; c += blockDim.x) {
            double dv = (a.q_in_kk[c] + a.q_in_kkp1[c]) / 2 * dt + a.sw_sn[c] * dt - (a.q_out_kk_1[c] + a.q_out_kk_2[c]) / 2 * dt
                        - (a.q_out_kkp1_1[c] + a.q_out_kkp1_2[c]) / 2 * dt;
            double v1 = a.volume_kk[c] + dv, h;
            if (v1 >= 0.0) h = v1 / a.cellarea; else { v1 = 0.0; h = 0.0; }
            a.volume_kkp1[c] = v1;
            a.h_water[c] = h;
            if (nsurf > 1 && sub < nsurf) {      // shift time levels for the next sub-step (:171-195)
                a.q_in_kk[c] = a.q_in_kkp1[c]; a.q_in_kkp1[c] = 0.0;
                a.q_out_kk_1[c] = a.q_out_kkp1_1[c]; a.q_out_kkp1_1[c] = 0.0;
                a.q_out_kk_2[c] = a.q_out_kkp1_2[c]; a.q_out_kkp1_2[c] = 0.0;
                a.volume_kk[c] = v1; a.volume_kkp1[c] = 0.0;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { *a.ak_max = s_akmax; *a.nsurf_out = nsurf; }
}

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
        LAUNCH(S, k_curves_chord, nblk(n, S->grid_n), RED_BLOCK, n, make_soil(S), S->p.kslope, S->p.tolksl, S->ptnew.p, S->ptold.p, S->pnew.p, S->ptimep.p, S->timep_dirty,
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
    void *args[] = {&a};
    CK(cudaEventRecord(S->evp0, S->st));
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
        CK(cudaEventRecord(S->evp1, S->st));
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
    S->barrier_epoch = (unsigned int)o.pad;
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

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

const char *cathy_last_error(void) { return g_err; }
int64_t cathy_sizeof_problem(void) { return (int64_t)sizeof(CathyProblem); }
int64_t cathy_sizeof_report(void) { return (int64_t)sizeof(CathyStepReport); }
int32_t cathy_abi_version(void) { return CATHY_ABI_VERSION; }

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
    if (const char *e = getenv("CATHY_GRAPH")) S->graph_mode = S->graph_mode && atoi(e) != 0;
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
    if (prob->kslope != 0 && !((prob->kslope == 1 || prob->kslope == 2) && prob->ivghu == 0 && prob->iopt == 1 && prob->dd_world <= 1))
        FAIL(-2, "KSLOPE=%d: chord slopes (1, 2) are implemented for van Genuchten curves (IVGHU=0) under Picard on one GPU; localized slopes (3, 4) are not", prob->kslope);
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
