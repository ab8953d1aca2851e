// bicg_res.cuh -- NSYSLV (SRC/solscal-extended.f:3063-3240; recurrence of GCSTAS, :1010-1128) with the Krylov vectors RESIDENT IN
// SHARED MEMORY and a vertical-line preconditioner, for Jacobians whose four resident vectors fit (n <= #CTAs x ~7 k rows,
// i.e. up to ~1 M nodes on one B200).  Included by cathy_b200.cu (shares Diag, IterOut, grid_barrier and the paired-row loads).
//
// Numbering.  The rest of the code numbers nodes layer-major like the reference (k = l NNOD + s).  This solver works in the
// COLUMN-major permutation k' = s L + l (L = NSTR + 1 node layers): the nodes of one DEM column are contiguous, so
//   * a CTA that owns a contiguous block of rows owns whole columns, and the preconditioner -- the block diagonal of J with one
//     nonsymmetric tridiagonal block per DEM column (layers are thin against the cell size, so the vertical coupling dominates;
//     the reference's ILU(0) is a sequential sweep, SRC/solscal-extended.f:2194-2267) -- becomes a Thomas solve on data that
//     sits in the CTA's shared memory: ~1 us per application instead of the 21 us of two latency-bound global sweeps;
//   * the half bandwidth drops from NNOD to (NC1 + 1) L, and the 15-point stencil keeps its pair structure: offsets
//     {1 | L-1, L | NC1 L - 1, NC1 L | (NC1+1) L - 1, (NC1+1) L}, so the aligned 16-byte paired-row loads of k_pcg_res2 apply.
// k_permute_cols transposes the 15 Jacobian diagonals, the right-hand side and the Dirichlet mask into that numbering once per
// solve (2 x 120 B/row, ~1 % of a solve); k_unpermute_cols brings the solution back.
//
// Per iteration (right-preconditioned BiCGSTAB, M = vertical-line blocks):
//   P1  v = J ph                    sigma = (rt, v)                                  -> grid reduction
//   P2  s = r - alpha v             sh = M^-1 s (Thomas, shared memory)              -> grid barrier
//   P3  t = J sh                    (t,s) (t,t) (rt,s) (rt,t)                        -> grid reduction
//   P4  x += alpha ph + omega sh    r = s - omega t, ||r||^2, p = r + beta (p - omega v), ph = M^-1 p   -> grid reduction
// r/s, v and a work vector (t between P3 and P4, the Thomas input / output otherwise) live in shared memory for the whole solve, and so
// do the line factors (fp32); ph and sh (which the neighbours' stencils need), the shadow residual rt, x and p (touched in P4 only) go
// through L2.  Global bytes per row and iteration: 2 x 15 diagonals + 2 x (vector window ~8 + rt 8) + sh, ph written (16) +
// P4: x, p read + written, ph, sh read (48) ~ 336 B (k_bicgstab: 448 B + two global line sweeps).  Dirichlet rows carry dinv = 0:
// all Krylov vectors stay exactly zero there.

struct BresArgs {
    int n, itmax;
    double tol;
    Diag U, L;                      // permuted Jacobian: U.d[0..7] = J(k, k + off), L.d[1..7][k] = J(k + off, k); off = permuted offsets
    const double *rhs, *dinv;       // permuted
    double *x, *ph, *sh, *rt, *p;   // permuted global vectors (ph, sh, x carry zero halos)
    double *partial;                // [2][5][gridDim.x]
    unsigned int *counter;
    unsigned int epoch0;
    IterOut *out;
    int rows_cta, nl, cols_cta;     // rows per CTA (= cols_cta * nl, even), node layers, columns per CTA
    int zigzag;
    unsigned long long *prof;       // diagnostic (CATHY_BRES_PROF=1): ns spent by CTA 0 in each phase, accumulated over all solves
    const unsigned char *symf;      // [CTA][pass][warp]: 1 = for the 64 rows of that warp-pass every lower-triangle entry equals its transpose
                                    // (k_bres_sym_flags): the product reads the upper arrays instead -- see row_pair_n
    int prefetch;                   // 1: L2 prefetch of the matrix streams of the thread's next pass
    int point;                      // 1: point Jacobi instead of the line blocks (diagnostic, CATHY_BRES_POINT=1)
};

// NV sums at once (NV <= 5): the scheme of grid_reduce2 -- block partials, ONE thread arrives and spins (alone in its warp: no lane
// of warp 0 waits at a warp barrier meanwhile), double-buffered partials, warp w re-reads the partials of quantity w with
// independent loads in a fixed order -> bit-reproducible, identical in all CTAs
template <int BLOCK, int NV>
__device__ __forceinline__ void grid_reduce_v(unsigned int *counter, unsigned int &epoch, unsigned int &par, const double *in, double *partial,
                                              double (*sh)[5], double (*res)[5], double *out)
{
    static_assert(BLOCK == 1024 && NV <= 5, "32 warps");
    const int nb = gridDim.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double v[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        v[q] = in[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(FULLMASK, v[q], o);
    }
    if (lane == 0)
#pragma unroll
        for (int q = 0; q < NV; ++q) sh[w][q] = v[q];
    __syncthreads();
    double *pp = partial + (size_t)par * 5 * nb;
    if (w == 0) {
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            double t = sh[lane][q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(FULLMASK, t, o);
            if (lane == 0) pp[q * nb + blockIdx.x] = t;
        }
        __syncwarp();
        epoch += nb;
        if (lane == 0) {
            __threadfence();
            atomicAdd(counter, 1u);
            unsigned int c;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(c) : "l"(counter) : "memory"); } while ((int)(c - epoch) < 0);
        }
    } else
        epoch += nb;
    __syncthreads();
    if (w < NV) {
        constexpr int MAXJ = 5;    // up to 160 CTAs (B200: 148)
        double u[MAXJ];
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) {
            const int i = lane + 32 * j;
            u[j] = 0.0;
            if (i < nb) asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(u[j]) : "l"(pp + w * nb + i) : "memory");
        }
        double t = (((u[0] + u[1]) + u[2]) + u[3]) + u[4];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(FULLMASK, t, o);
        if (lane == 0) res[par][w] = t;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NV; ++q) out[q] = res[par][q];
    par ^= 1u;
}

// one pair of diagonals (o, o+1) = (da, da+1) of a NONSYMMETRIC matrix: upper products from U, lower from Lw, rows k and k+1
template <bool ODD>
__device__ __forceinline__ void pair_group_n(const Diag &U, const Diag &Lw, const double *x, int da, int o, int k, bool elo, bool ehi, double &a0, double &a1)
{
    const double2 ua = *reinterpret_cast<const double2 *>(U.d[da] + k), ub = *reinterpret_cast<const double2 *>(U.d[da + 1] + k);
    const PairLd rw = win3_ld<ODD>(x + k + o, elo, ehi), rm = win3_ld<!ODD>(x + k - o - 1, elo, ehi);
    const PairLd ra = pair_ld<ODD>(Lw.d[da] + k - o, ehi), rb = pair_ld<!ODD>(Lw.d[da + 1] + k - o - 1, ehi);
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
// the matrix streams of rows kn, kn+1 (the pair this thread handles in its NEXT pass) are pulled into L2 while the current pass runs
__device__ __forceinline__ void prefetch_pair_n(const Diag &U, const Diag &Lw, int kn)
{
#pragma unroll
    for (int d = 0; d < NDIAG; ++d) l2_prefetch(U.d[d] + kn);
#pragma unroll
    for (int d = 1; d < NDIAG; ++d) l2_prefetch(Lw.d[d] + kn - U.off[d]);
}
// rows k (even) and k+1 of J x
template <int PAR>
__device__ __forceinline__ void row_pair_n(const Diag &U, const Diag &Lw, const double *x, int k, bool elo, bool ehi, int o2, int o4, int o6, double &a0, double &a1)
{
    const double2 xc = *reinterpret_cast<const double2 *>(x + k);
    double xm = __shfl_up_sync(FULLMASK, xc.y, 1), xp = __shfl_down_sync(FULLMASK, xc.x, 1);
    if (elo) xm = x[k - 1];
    if (ehi) xp = x[k + 2];
    const double2 dd = *reinterpret_cast<const double2 *>(U.d[0] + k), u1 = *reinterpret_cast<const double2 *>(U.d[1] + k);
    const double2 l1 = *reinterpret_cast<const double2 *>(Lw.d[1] + k);
    double l1m = __shfl_up_sync(FULLMASK, l1.y, 1);             // J(k, k-1) is stored at column index k-1
    if (elo) l1m = Lw.d[1][k - 1];
    a0 = dd.x * xc.x;   a1 = dd.y * xc.y;
    a0 += u1.x * xc.y;  a1 += u1.y * xp;
    a0 += l1m * xm;     a1 += l1.x * xc.x;
    pair_group_n<(PAR & 1) != 0>(U, Lw, x, 2, o2, k, elo, ehi, a0, a1);
    pair_group_n<(PAR & 2) != 0>(U, Lw, x, 4, o4, k, elo, ehi, a0, a1);
    pair_group_n<(PAR & 4) != 0>(U, Lw, x, 6, o6, k, elo, ehi, a0, a1);
}

// v <- M^-1 v IN PLACE for the columns of this CTA, everything in shared memory: forward and backward Thomas sweep, one thread per
// column.  The factors (reciprocal pivot, sub-diagonal, eliminated super-diagonal) are kept as fp32 (a preconditioner only has to be
// a fixed linear operator), [column][layer] with an odd stride LP so that the lanes of a warp hit different banks.
__device__ __forceinline__ void line_apply_inplace(const BresArgs &a, double *v, const float *fi, const float *fl, const float *fc, int ncol, int LP)
{
    const int L = a.nl;
    if (a.point) {
        for (int i = threadIdx.x; i < ncol * L; i += blockDim.x) v[i] *= a.dinv[blockIdx.x * a.rows_cta + i];
        return;
    }
    for (int c = threadIdx.x; c < ncol; c += blockDim.x) {
        const int b = c * L, f = c * LP;
        double y = v[b] * (double)fi[f];
        v[b] = y;
#pragma unroll 4
        for (int l = 1; l < L; ++l) {
            y = (v[b + l] - (double)fl[f + l - 1] * y) * (double)fi[f + l];
            v[b + l] = y;
        }
#pragma unroll 4
        for (int l = L - 2; l >= 0; --l) {
            y = v[b + l] - (double)fc[f + l] * y;
            v[b + l] = y;
        }
    }
}

// Where the derivative part of the Jacobian vanishes (saturated zones: d kr / d psi = d eta / d psi = 0, SRC/assnew.f) J is
// symmetric, J(k, k - o) = J(k - o, k): the product can take the lower-triangle entry of row k from the UPPER array at k - o, which
// another row streams anyway.  One flag per group of 64 rows (a warp-pass of k_bicgstab_res): on a fully saturated system (the coupled
// storm of BASELINE config 3) the solve then touches 8 diagonals (54 MB at 848 k rows, L2-resident) instead of 15 (102 MB, streamed
// from HBM twice per iteration).  Same arithmetic, same operands bit for bit -- the flag only says where they are equal.
__global__ void k_bres_sym_flags(int n, int rows_cta, int npass, Diag U, Diag L, unsigned char *flags)
{
    // one warp per group; group g = (cta, pass, warp): rows cta * rows_cta + pass * 2048 + warp * 64 + [0, 64)
    const int lane = threadIdx.x & 31;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long ngroup = (long long)gridDim.y * npass * 32;
    (void)ngroup;
    for (long long g = gw; g < (long long)((n + rows_cta - 1) / rows_cta) * npass * 32; g += nwarp) {
        const int cta = (int)(g / (npass * 32)), rem = (int)(g % (npass * 32)), pass = rem >> 5, warp = rem & 31;
        const int i0 = pass * 2048 + warp * 64, cnt = min(rows_cta, n - cta * rows_cta);
        bool ok = true;
        for (int q = 0; q < 2; ++q) {
            const int i = i0 + 2 * lane + q;
            if (i < cnt) {
                const int k = cta * rows_cta + i;
#pragma unroll
                for (int d = 1; d < NDIAG; ++d) { const int c = k - U.off[d]; ok = ok && (L.d[d][c] == U.d[d][c]); }
            }
        }
        ok = __all_sync(0xffffffffu, ok);
        if (lane == 0) flags[g] = ok ? 1 : 0;
    }
}

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define BRES_TICK(slot)                                                                         \
    do {                                                                                        \
        if (a.prof && blockIdx.x == 0 && threadIdx.x == 0) { const unsigned long long t_ = gtime(); a.prof[slot] += t_ - tprof; tprof = t_; } \
    } while (0)

template <int BLOCK, int PAR>     // PAR: parities of the permuted offsets off[2], off[4], off[6] (bits 0, 1, 2)
__global__ void __launch_bounds__(BLOCK, 1) k_bicgstab_res(BresArgs a)
{
    unsigned long long tprof = a.prof ? gtime() : 0ull;
    extern __shared__ __align__(16) double smv[];
    __shared__ double sh[BLOCK / 32][5];
    __shared__ double res[2][5];
    unsigned int epoch = a.epoch0, par = 0;
    const int R = a.rows_cta, row0 = blockIdx.x * R, cnt = max(0, min(R, a.n - row0)), tid = threadIdx.x, lane = tid & 31;
    const int ncol = cnt / a.nl, LP = a.nl | 1, FS = a.cols_cta * LP;
    // resident: r (s after P2), v, and a work vector w that holds t between P3 and P4 and the Thomas input / output otherwise;
    // the search direction p is only touched in P4 and lives in global memory
    double *rs = smv, *vs = smv + R, *ws = smv + 2 * (size_t)R;
    float *ffi = reinterpret_cast<float *>(smv + 3 * (size_t)R), *ffl = ffi + FS, *ffc = ffl + FS;
    const double *__restrict__ di = a.dinv;
    const int o2 = a.U.off[2], o4 = a.U.off[4], o6 = a.U.off[6];
    const int last = (cnt - 1) & ~1;
    const int npass = (cnt + 2 * BLOCK - 1) / (2 * BLOCK), npass_max = (R + 2 * BLOCK - 1) / (2 * BLOCK);
    const bool PF = a.prefetch != 0;
    const unsigned char *__restrict__ symf = a.symf + (size_t)blockIdx.x * npass_max * 32;
    double in[5] = {0, 0, 0, 0, 0}, out[5];
    // ---- x0 = D^-1 b, ||b_free||^2, Dirichlet rows of this thread as a bit mask (pass j: rows 2 tid + 2 BLOCK j + {0,1} -> bits 2j, 2j+1)
    unsigned int dmask = 0;
    for (int i = 2 * tid, j = 0; i < cnt; i += 2 * BLOCK, ++j)
#pragma unroll
        for (int q = 0; q < 2; ++q)
            if (i + q < cnt) {
                const int k = row0 + i + q;
                const double b = a.rhs[k], d = di[k];
                a.x[k] = b * d;
                if (d == 0.0) dmask |= 1u << (2 * j + q); else in[0] += b * b;
            }
    // ---- Thomas factors of this CTA's columns (once per solve, fp64 arithmetic, stored as fp32)
    {
        const int L = a.nl;
        for (int c = tid; c < ncol; c += BLOCK) {
            double cprev = 0.0, loprev = 0.0;
            for (int l = 0, k = row0 + c * L; l < L; ++l, ++k) {
                double idn = 0.0, cc = 0.0;
                const double lo = l + 1 < L ? a.L.d[1][k] : 0.0;
                if (di[k] != 0.0) {
                    const double piv = a.U.d[0][k] - loprev * cprev;
                    idn = 1.0 / piv;
                    cc = l + 1 < L ? a.U.d[1][k] * idn : 0.0;
                }
                ffi[c * LP + l] = (float)idn; ffc[c * LP + l] = (float)cc; ffl[c * LP + l] = (float)lo;
                cprev = cc; loprev = lo;
            }
        }
    }
    grid_reduce_v<BLOCK, 1>(a.counter, epoch, par, in, a.partial, sh, res, out);
    const double xlung = out[0];
    // ---- r0 = b - J x0 (zero on Dirichlet rows), rt = r0, p = r0, v = 0; rho = (rt, r0); ph = M^-1 p
    in[0] = 0.0;
    for (int iw = 2 * (tid - lane), j = 0; iw < cnt; iw += 2 * BLOCK, ++j) {
        const int i_own = iw + 2 * lane;
        const bool act = i_own < cnt, ok1 = i_own + 1 < cnt;
        const int i = act ? i_own : last;
        const bool ehi = lane == 31 || i_own + 2 >= cnt, elo = lane == 0;
        const int k = row0 + i;
        double a0, a1;
        row_pair_n<PAR>(a.U, symf[j * 32 + (tid >> 5)] ? a.U : a.L, a.x, k, elo, ehi, o2, o4, o6, a0, a1);
        if (act) {
            const double r0 = ((dmask >> (2 * j)) & 1u) ? 0.0 : a.rhs[k] - a0;
            const double r1 = (!ok1 || ((dmask >> (2 * j + 1)) & 1u)) ? 0.0 : a.rhs[k + 1] - a1;
            rs[i] = r0; ws[i] = r0; vs[i] = 0.0; a.rt[k] = r0; a.p[k] = r0;
            in[0] += r0 * r0;
            if (ok1) { rs[i + 1] = r1; ws[i + 1] = r1; vs[i + 1] = 0.0; a.rt[k + 1] = r1; a.p[k + 1] = r1; in[0] += r1 * r1; }
        }
    }
    __syncthreads();
    line_apply_inplace(a, ws, ffi, ffl, ffc, ncol, LP);
    __syncthreads();
    for (int i = tid; i < cnt; i += BLOCK) a.ph[row0 + i] = ws[i];
    grid_reduce_v<BLOCK, 1>(a.counter, epoch, par, in, a.partial, sh, res, out);
    double rho = out[0], err = xlung > 0.0 ? sqrt(out[0] / xlung) : sqrt(out[0] / a.n);
    int niter = 0;
    if (rho == 0.0 || err <= a.tol) {
        if (blockIdx.x == 0 && tid == 0) { a.out->pcg_niter = 1; a.out->pcg_err = err; a.out->pad = (int)epoch; }
        return;
    }
    bool breakdown = false;
    BRES_TICK(0);       // set-up
    for (;;) {
        ++niter;
        // ---- P1: v = J ph, sigma = (rt, v)
        in[0] = 0.0;
        for (int iw = 2 * (tid - lane), j = 0; iw < cnt; iw += 2 * BLOCK, ++j) {
            const int i_own = iw + 2 * lane;
            const bool act = i_own < cnt, ok1 = i_own + 1 < cnt;
            const int i = act ? i_own : last;
            const bool ehi = lane == 31 || i_own + 2 >= cnt, elo = lane == 0;
            const int k = row0 + i;
            if (PF && i_own + 2 * BLOCK < cnt) prefetch_pair_n(a.U, a.L, k + 2 * BLOCK);
            double a0, a1;
            row_pair_n<PAR>(a.U, symf[j * 32 + (tid >> 5)] ? a.U : a.L, a.ph, k, elo, ehi, o2, o4, o6, a0, a1);
            if (act) {
                if ((dmask >> (2 * j)) & 1u) a0 = 0.0;
                if ((dmask >> (2 * j + 1)) & 1u) a1 = 0.0;
                if (ok1) {
                    const double2 rtv = *reinterpret_cast<const double2 *>(a.rt + k);
                    *reinterpret_cast<double2 *>(vs + i) = make_double2(a0, a1);
                    in[0] += rtv.x * a0; in[0] += rtv.y * a1;
                } else { vs[i] = a0; in[0] += a.rt[k] * a0; }
            }
        }
        BRES_TICK(1);   // P1 product
        grid_reduce_v<BLOCK, 1>(a.counter, epoch, par, in, a.partial, sh, res, out);
        BRES_TICK(2);   // reduction 1
        const double alpha = rho / out[0];
        // ---- P2: s = r - alpha v (in place of r, copy in w), sh = M^-1 s
        for (int i = 2 * tid; i < cnt; i += 2 * BLOCK) {
            if (i + 1 < cnt) {
                double2 r = *reinterpret_cast<double2 *>(rs + i);
                const double2 v = *reinterpret_cast<const double2 *>(vs + i);
                r.x -= alpha * v.x; r.y -= alpha * v.y;
                *reinterpret_cast<double2 *>(rs + i) = r;
                *reinterpret_cast<double2 *>(ws + i) = r;
            } else { const double r = rs[i] - alpha * vs[i]; rs[i] = r; ws[i] = r; }
        }
        __syncthreads();
        BRES_TICK(3);   // s update
        line_apply_inplace(a, ws, ffi, ffl, ffc, ncol, LP);
        __syncthreads();
        BRES_TICK(4);   // Thomas
        for (int i = 2 * tid; i < cnt; i += 2 * BLOCK) {
            if (i + 1 < cnt) *reinterpret_cast<double2 *>(a.sh + row0 + i) = *reinterpret_cast<const double2 *>(ws + i);
            else a.sh[row0 + i] = ws[i];
        }
        grid_barrier(a.counter, epoch);
        BRES_TICK(5);   // sh copy-out + grid barrier
        // ---- P3: t = J sh (into w); (t,s), (t,t), (rt,s), (rt,t).  zigzag: the passes run from the last to the first, so the sweep
        // starts on the part of the Jacobian that P1 read last and that is still in the L2
        in[0] = in[1] = in[2] = in[3] = 0.0;
        for (int jj = 0; jj < npass; ++jj) {
            const int j = a.zigzag ? npass - 1 - jj : jj;
            const int iw = 2 * (tid - lane) + 2 * BLOCK * j;
            if (iw >= cnt) continue;                 // warp-uniform
            const int i_own = iw + 2 * lane;
            const bool act = i_own < cnt, ok1 = i_own + 1 < cnt;
            const int i = act ? i_own : last;
            const bool ehi = lane == 31 || i_own + 2 >= cnt, elo = lane == 0;
            const int k = row0 + i;
            if (PF) { const int in_ = a.zigzag ? i_own - 2 * BLOCK : i_own + 2 * BLOCK; if (in_ >= 0 && in_ < cnt) prefetch_pair_n(a.U, a.L, row0 + in_); }
            double a0, a1;
            row_pair_n<PAR>(a.U, symf[j * 32 + (tid >> 5)] ? a.U : a.L, a.sh, k, elo, ehi, o2, o4, o6, a0, a1);
            if (act) {
                if ((dmask >> (2 * j)) & 1u) a0 = 0.0;
                if ((dmask >> (2 * j + 1)) & 1u) a1 = 0.0;
                if (ok1) {
                    const double2 rtv = *reinterpret_cast<const double2 *>(a.rt + k), s = *reinterpret_cast<const double2 *>(rs + i);
                    *reinterpret_cast<double2 *>(ws + i) = make_double2(a0, a1);
                    in[0] += a0 * s.x; in[1] += a0 * a0; in[2] += rtv.x * s.x; in[3] += rtv.x * a0;
                    in[0] += a1 * s.y; in[1] += a1 * a1; in[2] += rtv.y * s.y; in[3] += rtv.y * a1;
                } else {
                    const double rtv = a.rt[k], s = rs[i];
                    ws[i] = a0;
                    in[0] += a0 * s; in[1] += a0 * a0; in[2] += rtv * s; in[3] += rtv * a0;
                }
            }
        }
        BRES_TICK(6);   // P3 product
        grid_reduce_v<BLOCK, 4>(a.counter, epoch, par, in, a.partial, sh, res, out);
        BRES_TICK(7);   // reduction 4
        const double omega = out[1] > 0.0 ? out[0] / out[1] : 0.0;
        const double rho_new = out[2] - omega * out[3];
        breakdown = omega == 0.0 || rho_new == 0.0;
        const double beta = breakdown ? 0.0 : (rho_new / rho) * (alpha / omega);
        // ---- P4: x += alpha ph + omega sh; r = s - omega t; ||r||^2 summed directly (the algebraic form cancels on ill-conditioned
        // systems); p = r + beta (p - omega v) -> global and w; ph = M^-1 p
        in[0] = 0.0;
        for (int i = 2 * tid; i < cnt; i += 2 * BLOCK) {
            const int k = row0 + i;
            if (i + 1 < cnt) {
                double2 x = *reinterpret_cast<double2 *>(a.x + k), p = *reinterpret_cast<double2 *>(a.p + k);
                const double2 ph = *reinterpret_cast<const double2 *>(a.ph + k), shv = *reinterpret_cast<const double2 *>(a.sh + k);
                x.x = x.x + alpha * ph.x + omega * shv.x; x.y = x.y + alpha * ph.y + omega * shv.y;
                *reinterpret_cast<double2 *>(a.x + k) = x;
                double2 r = *reinterpret_cast<double2 *>(rs + i);
                const double2 t = *reinterpret_cast<const double2 *>(ws + i), v = *reinterpret_cast<const double2 *>(vs + i);
                r.x -= omega * t.x; r.y -= omega * t.y;
                in[0] += r.x * r.x; in[0] += r.y * r.y;
                p.x = r.x + beta * (p.x - omega * v.x); p.y = r.y + beta * (p.y - omega * v.y);
                *reinterpret_cast<double2 *>(rs + i) = r; *reinterpret_cast<double2 *>(ws + i) = p; *reinterpret_cast<double2 *>(a.p + k) = p;
            } else {
                a.x[k] = a.x[k] + alpha * a.ph[k] + omega * a.sh[k];
                const double r = rs[i] - omega * ws[i];
                rs[i] = r; in[0] += r * r;
                const double p = r + beta * (a.p[k] - omega * vs[i]);
                ws[i] = p; a.p[k] = p;
            }
        }
        __syncthreads();
        BRES_TICK(8);   // P4 vector updates
        line_apply_inplace(a, ws, ffi, ffl, ffc, ncol, LP);
        __syncthreads();
        BRES_TICK(9);   // Thomas
        for (int i = 2 * tid; i < cnt; i += 2 * BLOCK) {
            if (i + 1 < cnt) *reinterpret_cast<double2 *>(a.ph + row0 + i) = *reinterpret_cast<const double2 *>(ws + i);
            else a.ph[row0 + i] = ws[i];
        }
        grid_reduce_v<BLOCK, 1>(a.counter, epoch, par, in, a.partial, sh, res, out);
        BRES_TICK(10);  // ph copy-out + reduction
        if (a.prof && blockIdx.x == 0 && tid == 0) a.prof[15] += 1;
        err = xlung > 0.0 ? sqrt(out[0] / xlung) : sqrt(out[0] / a.n);
        if (!(err > a.tol) || niter >= a.itmax || breakdown) break;
        rho = rho_new;
    }
    // a breakdown (NaN / zero inner products) without convergence is reported as "ITMXCG reached" so that FLOW3D back-steps
    if (blockIdx.x == 0 && tid == 0) { a.out->pcg_niter = (err > a.tol || !(err == err)) ? max(niter, a.itmax) : niter; a.out->pcg_err = err; a.out->pad = (int)epoch; }
}

// layer-major [L][NNOD] -> column-major [NNOD][L] for up to 17 arrays at once (blockIdx.y = array): 32 surface nodes x all layers per
// block through a shared-memory tile, both sides coalesced.  shift: the destination index is lowered by it (families whose
// upper / lower role swaps under the permutation are indexed by the other end of the entry); out-of-range destinations are
// structurally zero entries and are skipped.
struct PermArgs { const double *src[17]; double *dst[17]; int shift[17]; int nnod, nl, n; };
__global__ void k_permute_cols(PermArgs a)
{
    extern __shared__ double tile[];   // [nl][33]
    const int q = blockIdx.y, s0 = blockIdx.x * 32, cntx = 32 * a.nl;
    const double *__restrict__ src = a.src[q];
    for (int idx = threadIdx.x; idx < cntx; idx += blockDim.x) {
        const int l = idx >> 5, sx = idx & 31, s = s0 + sx;
        tile[l * 33 + sx] = s < a.nnod ? src[(size_t)l * a.nnod + s] : 0.0;
    }
    __syncthreads();
    double *dst = a.dst[q];
    const int shift = a.shift[q];
    for (int idx = threadIdx.x; idx < cntx; idx += blockDim.x) {
        const int sx = idx / a.nl, l = idx - sx * a.nl, s = s0 + sx;
        const long long kd = (long long)s * a.nl + l - shift;
        if (s < a.nnod && kd >= 0 && kd < a.n) dst[kd] = tile[l * 33 + sx];
    }
}
__global__ void k_unpermute_cols(int nnod, int nl, const double *__restrict__ src, double *__restrict__ dst)
{
    extern __shared__ double tile[];   // [nl][33]
    const int s0 = blockIdx.x * 32, cntx = 32 * nl;
    for (int idx = threadIdx.x; idx < cntx; idx += blockDim.x) {
        const int sx = idx / nl, l = idx - sx * nl, s = s0 + sx;
        if (s < nnod) tile[l * 33 + sx] = src[(size_t)s * nl + l];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < cntx; idx += blockDim.x) {
        const int l = idx >> 5, sx = idx & 31, s = s0 + sx;
        if (s < nnod) dst[(size_t)l * nnod + s] = tile[l * 33 + sx];
    }
}
