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

// ------------------------------------------------------------------------------------------------------------------------------
// k_pcg_cl2: the same solve with ONE cluster barrier per iteration and no global-memory traffic inside the iteration.
//
// Recurrence: the single-reduction form of preconditioned CG (Chronopoulos & Gear; k_pcg2 uses it on the full grid): with
// u = M^-1 r and w = A u,
//     gamma = (r,u), delta = (w,u);   beta = gamma / gamma_old;   alpha = gamma / (delta - beta gamma / alpha_old)
//     p = u + beta p;  s = w + beta s;  x += alpha p;  r -= alpha s;  u = M^-1 r;  w = A u
// -- the same iterates as GRADDP in exact arithmetic (same x0 = M^-1 b, same stopping test on ||r_free|| / ||b_free||).
// Every CTA keeps r, s, u and M^-1 not only for its R rows but also for the H = NNOD rows on either side (its stencil window) and
// updates them redundantly, so the only data a neighbour must supply per iteration is w on those halo rows: the owner pushes it
// straight into the neighbour's shared memory (st.shared::cluster) while it computes it, double-buffered by iteration parity,
// and the reduction's cluster barrier publishes it.  Needs H <= R (only the two neighbouring CTAs hold a row's copies).
// ------------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cluster_reduce3(cg::cluster_group &cl, unsigned int &par, double a, double b, double c, double (*sh)[3],
                                                double (*cp)[PCG_CL_MAX][4], double &ra, double &rb, double &rc)
{
    const int nc = (int)cl.num_blocks(), lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(FULLMASK, a, o); b += __shfl_xor_sync(FULLMASK, b, o); c += __shfl_xor_sync(FULLMASK, c, o); }
    if (lane == 0) { sh[w][0] = a; sh[w][1] = b; sh[w][2] = c; }
    __syncthreads();
    if (w == 0) {
        double t0 = lane < nw ? sh[lane][0] : 0.0, t1 = lane < nw ? sh[lane][1] : 0.0, t2 = lane < nw ? sh[lane][2] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { t0 += __shfl_xor_sync(FULLMASK, t0, o); t1 += __shfl_xor_sync(FULLMASK, t1, o); t2 += __shfl_xor_sync(FULLMASK, t2, o); }
        if (lane < nc) {
            double *dst = cl.map_shared_rank(&cp[par][cl.block_rank()][0], lane);
            *reinterpret_cast<double2 *>(dst) = make_double2(t0, t1);
            dst[2] = t2;
        }
    }
    cl.sync();
    // the nc slots are added in rank order by lane 0..nc-1 -> butterfly would change the order between cluster sizes only, not between
    // CTAs: every CTA holds the same slots, so every CTA gets the same sums
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int q = 0; q < nc; ++q) { const double2 v = *reinterpret_cast<const double2 *>(&cp[par][q][0]); s0 += v.x; s1 += v.y; s2 += cp[par][q][2]; }
    ra = s0; rb = s1; rc = s2;
    par ^= 1u;
}

// Launched with FEW warps per CTA (default 256 threads, several rows per thread): the iteration is bound by instruction issue -- the
// reduction, barrier and loop-control instructions are per warp, not per row (ncu on k_pcg_cl with 1024 threads: 430 instructions
// per warp and iteration for one row per thread, issue slots 40 % busy, top stall ERRBAR/barrier) -- so fewer, fatter warps win.
__global__ void __launch_bounds__(1024, 1) k_pcg_cl2(PclArgs a)
{
    extern __shared__ __align__(16) double smv[];
    __shared__ double sh[32][3];
    __shared__ __align__(16) double cpart[2][PCG_CL_MAX][4];
    cg::cluster_group cl = cg::this_cluster();
    const int R = a.R, H = a.H, W = R + H, W2 = R + 2 * H, tid = threadIdx.x, nt = blockDim.x, nc = (int)cl.num_blocks(), rank = (int)cl.block_rank();
    const int row0 = rank * R, cnt = max(0, min(R, a.n - row0)), g0 = row0 - H;      // window row j <-> global row g0 + j
    double *Au = smv;                                          // [NDIAG][W]: entry j of diagonal d is A_d[g0 + j]
    double *rw = Au + (size_t)NDIAG * W, *sw = rw + W2, *uw = sw + W2, *dw = uw + W2, *w0 = dw + W2, *w1 = w0 + W2;   // windows
    double *xs = w1 + W2, *ps = xs + R, *fs = ps + R;          // own rows: x, p, 1 = free row / 0 = Dirichlet row
    for (int d = 0; d < NDIAG; ++d) {
        const double *src = d == 0 ? a.diag : a.A.d[d];
        for (int j = tid; j < W; j += nt) {
            const int g = g0 + j;
            Au[(size_t)d * W + j] = (g >= 0 && g < a.n) ? src[g] : 0.0;
        }
    }
    unsigned int par = 0;
    double xl = 0.0;
    for (int i = tid; i < cnt; i += nt) {
        const int k = row0 + i;
        const double b = a.rhs[k], d = 1.0 / a.diag[k];
        xs[i] = b * d; a.x[k] = b * d; ps[i] = 0.0;
        const bool dir = is_dirichlet(k, a.nnod, a.ifatm, a.contp_flag);
        fs[i] = dir ? 0.0 : 1.0;
        if (!dir) xl += b * b;
    }
    double xlung, d1, d2;
    cluster_reduce3(cl, par, xl, 0.0, 0.0, sh, cpart, xlung, d1, d2);       // its barrier publishes x0 (global) and the staged matrix
    // window: r = b - A x0, u = M^-1 r, s = 0
    for (int j = tid; j < W2; j += nt) {
        const int g = g0 + j;
        double r = 0.0, dv = 0.0;
        if (g >= 0 && g < a.n) { dv = 1.0 / a.diag[g]; r = a.rhs[g] - dia_row(a.A, a.diag, a.x, g, a.n); }
        rw[j] = r; dw[j] = dv; uw[j] = r * dv; sw[j] = 0.0; w0[j] = 0.0; w1[j] = 0.0;
    }
    cl.sync();        // nobody pushes into a window that is still being initialised
    int off[NDIAG];
#pragma unroll
    for (int d = 0; d < NDIAG; ++d) off[d] = a.A.off[d];
    double *nb_lo0 = rank > 0 ? cl.map_shared_rank(w0, rank - 1) : nullptr, *nb_lo1 = rank > 0 ? cl.map_shared_rank(w1, rank - 1) : nullptr;
    double *nb_hi0 = rank + 1 < nc ? cl.map_shared_rank(w0, rank + 1) : nullptr, *nb_hi1 = rank + 1 < nc ? cl.map_shared_rank(w1, rank + 1) : nullptr;
    double alfa = 0.0, beta = 0.0, gam = 1.0, rr_last = 0.0;
    const double den = xlung > 0.0 ? xlung : (double)a.n, tol2 = a.tol * a.tol * den;
    int niter = 0;
    unsigned int wb = 0;      // parity of the w buffer being WRITTEN
    for (;;) {
        double *wn = wb ? w1 : w0;
        double *plo = wb ? nb_lo1 : nb_lo0, *phi = wb ? nb_hi1 : nb_hi0;
        // ---- w = A u on the own rows (the u window is complete), pushed into the neighbours' halos; (r,u), (w,u), ||r_free||^2
        double s_g = 0.0, s_d = 0.0, s_r = 0.0;
        for (int i = tid; i < cnt; i += nt) {
            const int j = i + H;
            double acc = Au[j] * uw[j];
#pragma unroll
            for (int d = 1; d < NDIAG; ++d) acc += Au[(size_t)d * W + j] * uw[j + off[d]];
#pragma unroll
            for (int d = 1; d < NDIAG; ++d) acc += Au[(size_t)d * W + j - off[d]] * uw[j - off[d]];
            wn[j] = acc;
            if (i < H && plo) plo[i + R + H] = acc;            // upper halo of the CTA that owns the rows before these (rank - 1)
            if (i >= R - H && phi) phi[i - R + H] = acc;       // lower halo of the CTA that owns the rows after these (rank + 1)
            const double r = rw[j], u = uw[j];
            s_g += r * u; s_d += acc * u; s_r += fs[i] * (r * r);
        }
        double gn, dl, rr;
        cluster_reduce3(cl, par, s_g, s_d, s_r, sh, cpart, gn, dl, rr);     // its barrier publishes the pushed halo rows of w
        // stopping test err = sqrt(rr / xlung) > tol without the square root and the division (they sit on the serial path of
        // every iteration): rr > tol^2 xlung; the reported error is formed once, at the exit
        rr_last = rr;
        if (niter > 0 && !(rr > tol2 && niter < a.itmax)) break;
        if (niter == 0) { beta = 0.0; alfa = gn / dl; }
        else { beta = gn / gam; alfa = gn / (dl - beta * gn / alfa); }
        gam = gn;
        ++niter;
        // ---- p = u + beta p, x += alfa p (own rows); s = w + beta s, r -= alfa s, u = M^-1 r (whole window, redundantly)
        for (int i = tid; i < cnt; i += nt) { const double p = uw[i + H] + beta * ps[i]; ps[i] = p; xs[i] += alfa * p; }
        __syncthreads();          // p read the OLD u of the own rows
        for (int j = tid; j < W2; j += nt) {
            const double s = wn[j] + beta * sw[j], r = rw[j] - alfa * s;
            sw[j] = s; rw[j] = r; uw[j] = r * dw[j];
        }
        __syncthreads();
        wb ^= 1u;
    }
    for (int i = tid; i < cnt; i += nt) a.x[row0 + i] = xs[i];
    if (blockIdx.x == 0 && tid == 0) { a.out->pcg_niter = max(niter, 1); a.out->pcg_err = sqrt(rr_last / den); a.out->pad = (int)a.epoch0; }
}
