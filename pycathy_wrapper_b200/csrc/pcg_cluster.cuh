// k_pcg_cl: SYMSLV / GRADDP (SRC/solscal-extended.f:4669-4699, :1260-1380) for SMALL meshes -- the whole solve inside ONE thread-block
// cluster, matrix and CG vectors resident in the cluster's shared memory (included by cathy_b200.cu after k_pcg_res2).
//
// Why: on a mesh of a few thousand rows (BASELINE config 1: 7,056 rows; members of small-catchment ensembles) the full-grid resident
// kernel is bound by latency, not by traffic: 7 us per iteration = four dependent L2 round trips for the stencil operands plus two
// global-memory grid barriers, on a Jacobi-preconditioned solve of ~150 iterations.  Here every CTA of the cluster (<= 16 CTAs,
// one row per thread) first copies the 8 upper diagonals of its rows AND of the H = NNOD rows before them (the lower triangle of a
// row is the upper triangle of the rows it couples to) into shared memory; after that an iteration touches global memory only for
// z = M^-1 r (one L2 round trip of 15 independent loads per row), and the two reductions are cluster-scope (cluster_reduce2:
// partial sums pushed into every CTA's shared memory, one hardware cluster barrier).  A solve occupies its cluster's SMs only, so
// the solves of other ensemble members run beside it.  Same recurrence, start vector and stopping test as k_pcg / k_pcg_res2.
#pragma once

struct PclArgs {
    int n, nnod, itmax, R, H;      // R rows per CTA (<= blockDim), H = largest stencil offset
    double tol;
    Diag A;
    const double *diag;            // main diagonal with the Dirichlet penalty
    const double *rhs;
    double *x, *z;
    const int *ifatm;
    const unsigned char *contp_flag;
    IterOut *out;
    unsigned int epoch0;
};

__global__ void __launch_bounds__(1024, 1) k_pcg_cl(PclArgs a)
{
    extern __shared__ __align__(16) double smv[];
    __shared__ double sh[32][2];
    __shared__ __align__(16) double cpart[2][PCG_CL_MAX][2];
    cg::cluster_group cl = cg::this_cluster();
    const int R = a.R, H = a.H, W = R + H, tid = threadIdx.x;
    const int row0 = blockIdx.x * R, cnt = max(0, min(R, a.n - row0)), k = row0 + tid;
    const bool act = tid < cnt;
    double *Au = smv;                                          // [NDIAG][W]: entry j of diagonal d is A_d[row0 - H + j]
    double *rs = Au + (size_t)NDIAG * W, *ps = rs + R, *bs = ps + R, *xs = bs + R, *dv = xs + R;
    for (int d = 0; d < NDIAG; ++d) {
        const double *src = d == 0 ? a.diag : a.A.d[d];
        for (int j = tid; j < W; j += blockDim.x) {
            const int g = row0 - H + j;
            Au[(size_t)d * W + j] = (g >= 0 && g < a.n) ? src[g] : 0.0;
        }
    }
    unsigned int par = 0;
    // x0 = M^-1 b ; xlung = ||b_free||^2
    bool dir = false;
    double xl = 0.0;
    if (act) {
        const double b = a.rhs[k], d = 1.0 / a.diag[k];
        dv[tid] = d; xs[tid] = b * d; a.x[k] = b * d;
        dir = is_dirichlet(k, a.nnod, a.ifatm, a.contp_flag);
        if (!dir) xl = b * b;
    }
    double xlung, d1;
    cluster_reduce2<1024>(cl, par, xl, 0.0, sh, cpart, xlung, d1);        // its barrier also publishes x0 and the staged matrix
    // r = b - A x0 ; z = M^-1 r ; p = B = 0
    if (act) {
        const double r = a.rhs[k] - dia_row(a.A, a.diag, a.x, k, a.n);
        rs[tid] = r; a.z[k] = r * dv[tid]; ps[tid] = 0.0; bs[tid] = 0.0;
    }
    cl.sync();
    const double *z = a.z;      // NOT __restrict__/read-only: rewritten every iteration by the other CTAs
    int off[NDIAG];
#pragma unroll
    for (int d = 0; d < NDIAG; ++d) off[d] = a.A.off[d];
    double beta = 0.0, err = 0.0;
    int niter = 1;
    for (;;) {
        // ---- phase A: B = A z + beta B, p = z + beta p, (p.r), (p.B)
        double s_pr = 0.0, s_pb = 0.0;
        if (act) {
            double zu[NDIAG], zl[NDIAG];
            const double zc = z[k];
#pragma unroll
            for (int d = 1; d < NDIAG; ++d) { zu[d] = z[k + off[d]]; zl[d] = z[k - off[d]]; }      // 14 independent loads: one L2 round trip
            double acc = Au[tid + H] * zc;
#pragma unroll
            for (int d = 1; d < NDIAG; ++d) acc += Au[(size_t)d * W + tid + H] * zu[d];
#pragma unroll
            for (int d = 1; d < NDIAG; ++d) acc += Au[(size_t)d * W + tid + H - off[d]] * zl[d];
            const double p = zc + beta * ps[tid], bq = acc + beta * bs[tid];
            ps[tid] = p; bs[tid] = bq;
            s_pr = p * rs[tid]; s_pb = p * bq;
        }
        double pr, pb;
        cluster_reduce2<1024>(cl, par, s_pr, s_pb, sh, cpart, pr, pb);
        const double alfa = pr / pb;
        // ---- phase B: r -= alfa B, x += alfa p, z = M^-1 r, (B.z), ||r_free||^2
        double s_bz = 0.0, s_rr = 0.0;
        if (act) {
            const double bq = bs[tid], r = rs[tid] - alfa * bq, zz = r * dv[tid];
            rs[tid] = r; xs[tid] += alfa * ps[tid];
            a.z[k] = zz;
            s_bz = bq * zz;
            if (!dir) s_rr = r * r;
        }
        double bz, rr;
        cluster_reduce2<1024>(cl, par, s_bz, s_rr, sh, cpart, bz, rr);     // its barrier publishes the new z
        beta = -bz / pb;
        err = xlung > 0.0 ? sqrt(rr / xlung) : sqrt(rr / a.n);
        if (err > a.tol && niter < a.itmax) { ++niter; continue; }
        break;
    }
    if (act) a.x[k] = xs[tid];
    if (blockIdx.x == 0 && tid == 0) { a.out->pcg_niter = niter; a.out->pcg_err = err; a.out->pad = (int)a.epoch0; }
}
