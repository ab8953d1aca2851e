// Seepage faces on the device (included by cathy_b200.cu after the back-calculated flux kernels).
//
// Reference: input/sfbc read by SFVONE (SRC/sfvone.f:27-66), initial exit points SFINIT (SRC/sfinit.f:26-56), the SFEX branches of
// BCPIC / BCNEW (SRC/bcpic.f:46-58, SRC/bcnew.f:43-57), SHLPIC / SHLNEW (SRC/shlpic.f:37-49), BKPIC / BKNEW (SRC/bkpic.f:41-48,
// SRC/bknew.f:37-52), FLUXMB (SRC/fluxmb.f:54-66), EXTALL + EXTCVG (SRC/extall.f:32-66, SRC/extcvg.f:21-26).
//
// All faces are flattened into one list of potential seepage nodes (face order, then position along the face -- the order of every
// reference loop).  An ACTUAL seepage node (SFEX = 1) is a Dirichlet node at psi = 0: it is expressed through the same dense
// flag / value arrays the prescribed-head nodes use (bit 1 of contp_flag), so assembly, the linear solvers and the head update need
// no seepage-specific code; the face lists are short (tens to thousands of nodes), every kernel here is one small launch.
#pragma once

// the reference spells the thresholds as single-precision literals (`1.0e-8`)
#define SF_EPS ((double)1.0e-8f)

struct SfOut { double sfflw; int ksf; int flag5; };

// SFINIT: nodes that start with a positive head become actual seepage nodes at psi = 0
__global__ void k_sf_init(int m, const int *__restrict__ node, int *__restrict__ sfex, int *__restrict__ sfexp, int *__restrict__ sfexit,
                          double *__restrict__ ptimep, double *__restrict__ pnew)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) {
        int k = node[j], e = 0;
        if (ptimep[k] >= SF_EPS) { e = 1; ptimep[k] = 0.0; pnew[k] = 0.0; }
        sfex[j] = e; sfexp[j] = e; sfexit[j] = e;
    }
}
// potential seepage nodes on the surface leave the atmospheric state machine (SRC/atmone.f:112-120, SRC/atmnxt.f:94-101)
__global__ void k_sf_mark_nonatm(int m, const int *__restrict__ node, int nnod, int *__restrict__ ifatm, int *__restrict__ ifatmp)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) {
        int k = node[j];
        if (k < nnod) { ifatm[k] = -1; if (ifatmp) ifatmp[k] = -1; }
    }
}
// BCPIC / BCNEW: bit 1 of the Dirichlet flag follows SFEX; the prescribed value of an actual seepage node is 0
__global__ void k_sf_apply(int m, const int *__restrict__ node, const int *__restrict__ sfex, unsigned char *__restrict__ flag,
                           double *__restrict__ val)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) {
        int k = node[j];
        unsigned char f = flag[k] & 1;
        if (sfex[j] == 1) { f |= 2; if (!(f & 1)) val[k] = 0.0; }
        flag[k] = f;
    }
}
// BKPIC / BKNEW + FLUXMB: SFQ of the actual seepage nodes, then SFFLW summed by ONE thread in list order (the reference's order)
template <bool NEWTON>
__global__ void k_sf_flux(int m, int n, const int *__restrict__ node, const int *__restrict__ sfex, Diag A, Diag L,
                          const double *__restrict__ diag_true, const double *__restrict__ dis, const double *__restrict__ pdiff,
                          const double *__restrict__ xt5, double tetaf, const double *__restrict__ sfqp, double *__restrict__ sfq,
                          SfOut *__restrict__ out)
{
    for (int j = threadIdx.x; j < m; j += blockDim.x)
        if (sfex[j] == 1) {
            int k = node[j];
            double scr = (NEWTON ? dia_row_n(A, L, pdiff, k) : dia_row_orig(A, diag_true, dis, pdiff, k, n)) - xt5[k];
            sfq[j] = (scr - (1.0 - tetaf) * sfqp[j]) * (1.0 / tetaf);
        }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        int f5 = 0;
        for (int j = 0; j < m; ++j)
            if (sfex[j] == 1) { s = s + sfq[j]; if (sfq[j] >= 0.0) ++f5; }
        out->sfflw = s; out->flag5 = f5;
    }
}
// EXTALL + EXTCVG: actual nodes with inflow become potential ones, potential nodes with a positive head become actual ones at psi = 0
__global__ void k_sf_extall(int m, const int *__restrict__ node, int *__restrict__ sfex, const int *__restrict__ sfexit,
                            double *__restrict__ sfq, double *__restrict__ pnew, SfOut *__restrict__ out)
{
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    int c = 0;
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        int k = node[j], e = sfex[j];
        if (e == 1) { if (sfq[j] >= 0.0) { e = 0; sfq[j] = 0.0; } }
        else if (pnew[k] > SF_EPS) { pnew[k] = 0.0; e = 1; }
        sfex[j] = e;
        if (e != sfexit[j]) ++c;
    }
    if (c) atomicAdd(&cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) out->ksf = cnt;
}
