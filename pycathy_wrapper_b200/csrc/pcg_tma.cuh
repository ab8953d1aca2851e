// pcg_tma.cuh -- SYMSLV (SRC/solscal-extended.f:4669-4699; recurrence of GRADDP, :1260-1380) for systems that do NOT fit the caches:
// streaming PCG in the column-major permutation k' = s L + l, every operand of the stencil phase staged into shared memory by
// TMA bulk copies (cp.async.bulk + mbarrier), one persistent CTA per SM, two-stage pipeline.  Included by cathy_b200.cu.
//
// Why column-major: layer-major, the 15-point stencil reaches NNOD rows up and down; on large meshes (config 5: NNOD = 1 M rows, ~100 MB
// of matrix + vectors per layer) the gathers of the neighbouring layers miss the L2 and come from HBM again.  Column-major, the half
// bandwidth is (NC1 + 1) L rows (31 k rows = 3 MB at config 5), and the stencil offsets form three clusters -- {0, +-1, +-(L-1), +-L},
// +-{NC1 L - 1, NC1 L, (NC1+1) L - 1, (NC1+1) L} -- so the z values a tile of T rows needs are THREE contiguous windows of ~T + L
// elements, and the lower-triangle entries of the tile are contiguous slices of the same 7 diagonal arrays: everything is a 1-D bulk copy.
//
// Recurrence (as k_pcg_res2): by linearity B = A p = A z + beta B_old, so only z is gathered.
//   phase A, per tile: TMA: 7 upper + diagonal + 7 lower slices, 3 z windows, p, B, r tiles -> p = z + beta p, B = A z + beta B, (p.r), (p.B)
//            (a hybrid that keeps the streaming operands in registers, loaded one tile ahead, and stages only the gathered ones
//            was measured slower: 134.9 against 126.3 us per iteration at 3.4 M rows).
//            Tiles are dealt round-robin (tile t -> CTA t mod #CTAs): at any time the CTAs read ONE contiguous band of every array.
//            Warp-specialised: one producer warp waits for a free stage (mbarrier `empty`) and issues the copies (mbarrier `full`,
//            complete_tx), 16 consumer warps (one row of the tile per thread) wait on `full`, compute, arrive on `empty`
//   phase B, streaming: r -= alfa B, x += alfa p, z = M^-1 r, (B.z), ||r_free||^2        (M = diag A; Dirichlet rows: penalty diagonal)
// Algorithmic bytes per row and iteration: 168; this kernel moves 64 (matrix) + 8 (z) + 24 (p, B, r in) + 16 (p, B out)
// + 64 (phase B) = 176 from HBM, plus 56 (lower slices) + 16 (two more z windows) that are L2 hits.
//
// Row-block partition (several GPUs, BASELINE config 5): a rank's rows [lo, hi) are contiguous in this numbering and so are its halo
// rows; phase B stores the z of the rank's first / last two node rows straight into the neighbours' ghost region of z (peer memory
// over NVLink, coalesced), the flags follow after the grid reduction, and in phase A every CTA runs the tiles whose windows lie inside
// the rank FIRST and only then waits for the neighbours' flag (interior first, boundary tiles after the peer flag).
#define TMA_T 512          // rows per tile = consumer threads
#define TMA_NS 2           // pipeline stages
#define TMA_BLOCK 544      // 16 consumer warps + 1 producer warp (its lane 0 issues the bulk copies)

struct TmaArgs {
    int n;                          // rows of the local arrays (window incl. ghost rows when partitioned)
    int lo, hi;                     // rows this rank owns and computes: [lo, hi), lo even
    int itmax;
    double tol;
    Diag A;                         // permuted upper off-diagonals d[1..7] (d[0] unused), off = permuted offsets
    const double *dg;               // diagonal with the Dirichlet penalty
    const double *rhs;
    double *dinv, *x, *r, *z, *p, *bv;
    double *partial;                // [3][gridDim.x]
    unsigned int *counter;
    unsigned int epoch0;
    IterOut *out;
    int tiles_cta;                  // tiles per CTA
    int nl;                         // node layers L
    // row-block partition
    int dd_on;
    DDCtx dd;
    double *zpeer_n, *zpeer_s;      // the neighbours' z arrays (their element 0)
    long long ndst0, sdst0;         // where my first / last nbr rows go in them (their south / north ghost region)
    int nbr;                        // halo rows: DD_W node rows = DD_W * NC1 * L
    unsigned long long *prof;       // diagnostic (CATHY_TMA_PROF=1): ns spent by CTA 0 per phase, accumulated
};

__device__ __forceinline__ unsigned int smem_u32(const void *p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned int bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned int parity)
{
    unsigned int ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
// 1-D TMA bulk copy global -> shared (SASS: UBLKCP); src, dst 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void tma_load(void *dst, const void *src, unsigned int bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// shared-memory image of one tile (offsets in doubles from the stage base)
struct TmaLayout {
    int up, dgt, lo, zn, zl, zh, pt, bt, rt, total;      // up: 7 x T, lo: 7 x (T + 2)
    int wn, wf;                                          // window lengths (even)
    int nb, fb, hb;                                      // index of row 0 of the tile inside the near / far-low / far-high window
};
__host__ __device__ inline TmaLayout tma_layout(const int *off, int L)
{
    TmaLayout t;
    const int T = TMA_T;
    t.wn = (T + 2 * L + 2 + 1) & ~1;
    t.wf = (T + (off[7] - off[4]) + 2 + 1) & ~1;
    int o = 0;
    t.up = o; o += 7 * T;
    t.dgt = o; o += T;
    t.lo = o; o += 7 * (T + 2);
    t.zn = o; o += t.wn;
    t.zl = o; o += t.wf;
    t.zh = o; o += t.wf;
    t.pt = o; o += T;
    t.bt = o; o += T;
    t.rt = o; o += T;
    t.total = (o + 15) & ~15;            // stages stay 128-byte aligned
    t.nb = L + (L & 1);                  // near window starts at (k0 - L) rounded down to even
    t.fb = off[7] + (off[7] & 1);        // far-low window starts at (k0 - off7) rounded down to even
    t.hb = -(off[4] - (off[4] & 1));     // far-high window starts at (k0 + off4) rounded down to even: z(k + o) = zh[i + o + hb]
    return t;
}

// producer: request everything tile [k0, k0 + T) needs; gv = the gathered vector (z; x0 in the set-up product), rsrc = r (rhs in the set-up)
__device__ __forceinline__ void tma_issue(const TmaArgs &a, const TmaLayout &ly, double *st, unsigned long long *bar, int k0, const double *gv,
                                          const double *rsrc, bool with_pb)
{
    const int T = TMA_T;
    const unsigned int bt = T * 8u, bl = (T + 2) * 8u;
    const unsigned int total = 8u * bt + 7u * bl + (unsigned int)ly.wn * 8u + 2u * (unsigned int)ly.wf * 8u + (with_pb ? 2u * bt : 0u) + (rsrc ? bt : 0u);
    mbar_expect_tx(bar, total);
#pragma unroll
    for (int d = 1; d < NDIAG; ++d) tma_load(st + ly.up + (d - 1) * T, a.A.d[d] + k0, bt, bar);
    tma_load(st + ly.dgt, a.dg + k0, bt, bar);
#pragma unroll
    for (int d = 1; d < NDIAG; ++d) tma_load(st + ly.lo + (d - 1) * (T + 2), a.A.d[d] + ((k0 - a.A.off[d]) & ~1), bl, bar);
    tma_load(st + ly.zn, gv + ((k0 - a.nl) & ~1), (unsigned int)ly.wn * 8u, bar);
    tma_load(st + ly.zl, gv + ((k0 - a.A.off[7]) & ~1), (unsigned int)ly.wf * 8u, bar);
    tma_load(st + ly.zh, gv + ((k0 + a.A.off[4]) & ~1), (unsigned int)ly.wf * 8u, bar);
    if (with_pb) { tma_load(st + ly.pt, a.p + k0, bt, bar); tma_load(st + ly.bt, a.bv + k0, bt, bar); }
    if (rsrc) tma_load(st + ly.rt, rsrc + k0, bt, bar);
}
// (A gv)_k for row i of the tile in stage st
__device__ __forceinline__ double tma_row(const TmaLayout &ly, const double *st, int i, const int *off, int L, double &zc)
{
    const int T = TMA_T;
    const double *zn = st + ly.zn + ly.nb + i, *zl = st + ly.zl + ly.fb + i, *zh = st + ly.zh + ly.hb + i;
    const double *up = st + ly.up + i, *lo = st + ly.lo + i;
    zc = zn[0];
    double acc = st[ly.dgt + i] * zc;
    acc += up[0 * T] * zn[1];
    acc += up[1 * T] * zn[L - 1];
    acc += up[2 * T] * zn[L];
    acc += up[3 * T] * zh[off[4]];
    acc += up[4 * T] * zh[off[5]];
    acc += up[5 * T] * zh[off[6]];
    acc += up[6 * T] * zh[off[7]];
    acc += lo[0 * (T + 2) + (off[1] & 1)] * zn[-1];
    acc += lo[1 * (T + 2) + (off[2] & 1)] * zn[-(L - 1)];
    acc += lo[2 * (T + 2) + (off[3] & 1)] * zn[-L];
    acc += lo[3 * (T + 2) + (off[4] & 1)] * zl[-off[4]];
    acc += lo[4 * (T + 2) + (off[5] & 1)] * zl[-off[5]];
    acc += lo[5 * (T + 2) + (off[6] & 1)] * zl[-off[6]];
    acc += lo[6 * (T + 2) + (off[7] & 1)] * zl[-off[7]];
    return acc;
}

template <bool DD>
__global__ void __launch_bounds__(TMA_BLOCK, 1) k_pcg_tma(TmaArgs a)
{
    extern __shared__ __align__(128) double stg[];            // TMA_NS stages
    __shared__ __align__(8) unsigned long long full[TMA_NS], empty[TMA_NS];
    __shared__ double sh[TMA_BLOCK / 32][3];
    __shared__ double shdd[4];
    cg::grid_group grid = cg::this_grid();
    constexpr int T = TMA_T, BLOCK = TMA_BLOCK, NCW = TMA_T / 32;      // NCW consumer warps, warp NCW produces
    const int tid = threadIdx.x, L = a.nl, lane = tid & 31, warp = tid >> 5;
    int off[NDIAG];
#pragma unroll
    for (int d = 0; d < NDIAG; ++d) off[d] = a.A.off[d];
    const TmaLayout ly = tma_layout(off, L);
    unsigned int epoch = a.epoch0, seq_ar = 0, seq_h = 0;
    if (DD) { seq_ar = a.dd.seq[0]; seq_h = a.dd.seq[1]; }
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < TMA_NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NCW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // tiles of this CTA: tile t of the rank (rows lo + t T ...) belongs to CTA t mod #CTAs; local tile j <-> t = blockIdx.x + j #CTAs
    const int ntile = (a.hi - a.lo + T - 1) / T, G = gridDim.x;
    const int nq = (int)blockIdx.x < ntile ? (ntile - (int)blockIdx.x + G - 1) / G : 0;
    auto k0_of = [&](int j) { return a.lo + ((int)blockIdx.x + j * G) * T; };
    // interior first: local tiles [ja, jb) read no ghost row; [0, ja) need the north neighbour's rows, [jb, nq) the south one's
    int ja = 0, jb = nq;
    if (DD) {
        const int reach = off[7] + 2;
        if (a.dd.north >= 0) while (ja < nq && k0_of(ja) - reach < a.lo) ++ja;
        if (a.dd.south >= 0) while (jb > ja && k0_of(jb - 1) + T + reach > a.hi) --jb;
    }
    auto tile_of = [&](int q) { const int ni = jb - ja; if (q < ni) return ja + q; const int r = q - ni; return r < ja ? r : jb + (r - ja); };
    // the rows of this CTA, tile after tile: element e of the CTA <-> row k0_of(e / T) + e % T
    const long long nrow_cta = (long long)nq * T;
    unsigned int gq = 0;                                      // tiles consumed so far (stage = gq % NS, parity = (gq / NS) & 1)
    double d1, d2;
    const bool has_n = DD && a.dd.north >= 0, has_s = DD && a.dd.south >= 0;
    // z (or x0) of my first / last nbr rows also goes into the neighbours' ghost region
    auto store_z = [&](int k, double v) -> bool {
        a.z[k] = v;
        bool sent = false;
        if (DD) {
            if (has_n && k < a.lo + a.nbr) { a.zpeer_n[a.ndst0 + (k - a.lo)] = v; sent = true; }
            if (has_s && k >= a.hi - a.nbr) { a.zpeer_s[a.sdst0 + (k - (a.hi - a.nbr))] = v; sent = true; }
        }
        return sent;
    };
    // after a grid-wide barrier that follows the stores: tell the neighbours; the matching wait sits in front of the first ghost-reading tile
    auto publish = [&]() {
        if (DD) {
            ++seq_h;
            if (blockIdx.x == 0 && tid == 0 && has_n) dd_release(&a.dd.peer[a.dd.north]->halo_flag[1], seq_h);
            if (blockIdx.x == 0 && tid == 1 && has_s) dd_release(&a.dd.peer[a.dd.south]->halo_flag[0], seq_h);
        }
    };
    auto wait_halo = [&]() {      // thread 0 only
        if (DD) {
            if (has_n) dd_wait(&a.dd.me->halo_flag[0], seq_h, a.dd.err, 2);
            if (has_s) dd_wait(&a.dd.me->halo_flag[1], seq_h, a.dd.err, 3);
            fence_proxy_async();
        }
    };
    // one stencil sweep over this CTA's tiles; MODE 0: set-up residual r = b - A gv;  MODE 1: phase A
    double s1 = 0.0, s2 = 0.0, beta = 0.0;
    auto sweep = [&](const double *gv, const double *rsrc, int mode) {
        if (warp == NCW) {
            if (lane == 0) {
                fence_proxy_async();                          // the gathered vector was written with ordinary stores (other SMs, acquired at the last barrier)
                for (int q = 0; q < nq; ++q) {
                    const unsigned int g = gq + q, s = g % TMA_NS;
                    mbar_wait(&empty[s], ((g / TMA_NS) & 1u) ^ 1u);      // the consumers have released the stage (passes at once the first time round)
                    if (DD && q == jb - ja) wait_halo();      // first tile that reads ghost rows
                    tma_issue(a, ly, stg + (size_t)s * ly.total, &full[s], k0_of(tile_of(q)), gv, rsrc, mode == 1);
                }
            }
            __syncwarp();
        } else {
            for (int q = 0; q < nq; ++q) {
                const unsigned int g = gq + q, s = g % TMA_NS;
                mbar_wait(&full[s], (g / TMA_NS) & 1u);
                const double *st = stg + (size_t)s * ly.total;
                const int i = tid, k = k0_of(tile_of(q)) + i;
                if (k < a.hi) {
                    double zc;
                    const double acc = tma_row(ly, st, i, off, L, zc);
                    if (mode == 0) a.r[k] = st[ly.rt + i] - acc;
                    else {
                        const double pk = zc + beta * st[ly.pt + i], bk = acc + beta * st[ly.bt + i];
                        a.p[k] = pk; a.bv[k] = bk;
                        s1 += pk * st[ly.rt + i]; s2 += pk * bk;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);        // this warp is done with the stage
            }
        }
        gq += nq;
    };
    // ---- set-up: x0 = M^-1 b (also into z, the gathered vector of the first sweep), ||b_free||^2, p = B = 0
    double xl = 0.0;
    bool sent = false;
    for (long long e = tid; e < nrow_cta; e += BLOCK) {
        const int k = k0_of((int)(e / T)) + (int)(e % T);
        if (k >= a.hi) continue;
        const double b = a.rhs[k], dgk = a.dg[k], dv = 1.0 / dgk, x0 = b * dv;
        a.dinv[k] = dv; a.x[k] = x0; a.p[k] = 0.0; a.bv[k] = 0.0;
        sent |= store_z(k, x0);
        if (!(dgk > 1.0e80)) xl += b * b;
    }
    if (DD && sent) __threadfence_system();
    double xlung;
    grid_reduce3<BLOCK, true>(grid, a.counter, epoch, xl, 0.0, 0.0, a.partial, sh, xlung, d1, d2);
    if (DD) { double v[1] = {xlung}; dd_allreduce<1>(a.dd, seq_ar, v, shdd); xlung = v[0]; }
    publish();
    sweep(a.z, a.rhs, 0);                                     // r = b - A x0
    grid_barrier(a.counter, epoch);
    if (DD) { double v[1] = {0.0}; dd_allreduce<1>(a.dd, seq_ar, v, shdd); }      // every rank has finished reading x0 from z
    sent = false;
    for (long long e = tid; e < nrow_cta; e += BLOCK) {
        const int k = k0_of((int)(e / T)) + (int)(e % T);
        if (k < a.hi) sent |= store_z(k, a.r[k] * a.dinv[k]);
    }
    if (DD && sent) __threadfence_system();
    grid_barrier(a.counter, epoch);
    publish();
    double err = 0.0;
    int niter = 1;
    unsigned long long tprof = 0;
    auto tick = [&](int slot) {
        if (a.prof && blockIdx.x == 0 && tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); if (slot >= 0) a.prof[slot] += t - tprof; tprof = t; }
    };
    tick(-1);
    for (;;) {
        // ---- phase A
        s1 = s2 = 0.0;
        sweep(a.z, a.r, 1);
        __syncthreads();
        tick(0);
        double pr, pb;
        grid_reduce3<BLOCK, true>(grid, a.counter, epoch, s1, s2, 0.0, a.partial, sh, pr, pb, d1);
        tick(1);
        if (DD) { double v[2] = {pr, pb}; dd_allreduce<2>(a.dd, seq_ar, v, shdd); pr = v[0]; pb = v[1]; }
        const double alfa = pr / pb;
        // ---- phase B: streaming, two rows per thread
        double s_bz = 0.0, s_rr = 0.0;
        sent = false;
        auto pairB = [&](int k, const double2 bk, const double2 pk, const double2 dv, double2 r, double2 x) {
            r.x -= alfa * bk.x; r.y -= alfa * bk.y;
            x.x += alfa * pk.x; x.y += alfa * pk.y;
            *reinterpret_cast<double2 *>(a.r + k) = r; *reinterpret_cast<double2 *>(a.x + k) = x;
            const double z0 = r.x * dv.x, z1 = r.y * dv.y;
            if (DD) { sent |= store_z(k, z0); sent |= store_z(k + 1, z1); }
            else *reinterpret_cast<double2 *>(a.z + k) = make_double2(z0, z1);
            s_bz += bk.x * z0; s_bz += bk.y * z1;
            if (dv.x > 1.0e-80) s_rr += r.x * r.x;
            if (dv.y > 1.0e-80) s_rr += r.y * r.y;
        };
        const long long npair = nrow_cta / 2;                    // T is even: a pair never straddles two tiles
        auto pair_row = [&](long long pi) { return k0_of((int)(pi / (T / 2))) + 2 * (int)(pi % (T / 2)); };
        long long pi = tid;
        for (; pi + BLOCK < npair; pi += 2 * BLOCK) {            // two pairs per trip: ten 16-byte loads in flight per thread
            const int k = pair_row(pi), k2 = pair_row(pi + BLOCK);
            if (k2 + 1 >= a.hi) break;                           // the ragged end goes through the loop below
            const double2 b0 = *reinterpret_cast<const double2 *>(a.bv + k), p0 = *reinterpret_cast<const double2 *>(a.p + k);
            const double2 d0 = *reinterpret_cast<const double2 *>(a.dinv + k), r0 = *reinterpret_cast<const double2 *>(a.r + k), x0 = *reinterpret_cast<const double2 *>(a.x + k);
            const double2 b1 = *reinterpret_cast<const double2 *>(a.bv + k2), p1 = *reinterpret_cast<const double2 *>(a.p + k2);
            const double2 d1v = *reinterpret_cast<const double2 *>(a.dinv + k2), r1 = *reinterpret_cast<const double2 *>(a.r + k2), x1 = *reinterpret_cast<const double2 *>(a.x + k2);
            pairB(k, b0, p0, d0, r0, x0);
            pairB(k2, b1, p1, d1v, r1, x1);
        }
        for (; pi < npair; pi += BLOCK) {
            const int k = pair_row(pi);
            if (k + 1 < a.hi)
                pairB(k, *reinterpret_cast<const double2 *>(a.bv + k), *reinterpret_cast<const double2 *>(a.p + k), *reinterpret_cast<const double2 *>(a.dinv + k),
                      *reinterpret_cast<const double2 *>(a.r + k), *reinterpret_cast<const double2 *>(a.x + k));
            else if (k < a.hi) {
                const double bk = a.bv[k], dv = a.dinv[k];
                const double r = a.r[k] - alfa * bk;
                a.r[k] = r; a.x[k] += alfa * a.p[k];
                const double z0 = r * dv;
                sent |= store_z(k, z0);
                s_bz += bk * z0;
                if (dv > 1.0e-80) s_rr += r * r;
            }
        }
        if (DD && sent) __threadfence_system();
        __syncthreads();
        tick(2);
        double bz, rr;
        grid_reduce3<BLOCK, true>(grid, a.counter, epoch, s_bz, s_rr, 0.0, a.partial, sh, bz, rr, d1);
        tick(3);
        if (a.prof && blockIdx.x == 0 && tid == 0) a.prof[15] += 1;
        if (DD) { double v[2] = {bz, rr}; dd_allreduce<2>(a.dd, seq_ar, v, shdd); bz = v[0]; rr = v[1]; }
        publish();
        beta = -bz / pb;
        err = xlung > 0.0 ? sqrt(rr / xlung) : sqrt(rr / a.n);
        if (err > a.tol && niter < a.itmax && !(DD && *(volatile int *)a.dd.err)) { ++niter; continue; }
        break;
    }
    if (DD) {
        // the solution of my boundary rows goes to the neighbours' ghost rows (BKPIC forms row products that reach one ghost row), through
        // the ghost region of z: nobody reads z any more (every rank has contributed to the last all-reduce, i.e. finished its last sweep)
        sent = false;
        for (long long e = tid; e < nrow_cta; e += BLOCK) {
            const int k = k0_of((int)(e / T)) + (int)(e % T);
            if (k >= a.hi) continue;
            const double xv = a.x[k];
            if (has_n && k < a.lo + a.nbr) { a.zpeer_n[a.ndst0 + (k - a.lo)] = xv; sent = true; }
            if (has_s && k >= a.hi - a.nbr) { a.zpeer_s[a.sdst0 + (k - (a.hi - a.nbr))] = xv; sent = true; }
        }
        if (sent) __threadfence_system();
        grid_barrier(a.counter, epoch);
        publish();
        if (tid == 0) wait_halo();
        __syncthreads();
        const int gtid = blockIdx.x * BLOCK + tid, gstride = gridDim.x * BLOCK;
        if (has_n) for (int k = a.lo - a.nbr + gtid; k < a.lo; k += gstride) a.x[k] = *(volatile double *)&a.z[k];
        if (has_s) for (int k = a.hi + gtid; k < a.hi + a.nbr; k += gstride) a.x[k] = *(volatile double *)&a.z[k];
    }
    if (blockIdx.x == 0 && tid == 0) {
        a.out->pcg_niter = niter; a.out->pcg_err = err; a.out->pad = (int)epoch;
        if (DD) { a.dd.seq[0] = seq_ar; a.dd.seq[1] = seq_h; }
    }
}

// y = A x alone, with the same staging as phase A of k_pcg_tma (the SpMV of GRADDP, SRC/solscal-extended.f:1326-1334, on HBM-resident
// systems): 7 upper + diagonal slices and the x windows from HBM (72 B/row), the lower slices and the two far windows from L2, y
// written once (8 B/row).  No grid barrier: one sweep per launch.  cathy_debug_spmv's kernel whenever the solver is k_pcg_tma.
__global__ void __launch_bounds__(TMA_BLOCK, 1) k_spmv_tma(TmaArgs a, const double *x, double *y)
{
    extern __shared__ __align__(128) double stg[];
    __shared__ __align__(8) unsigned long long full[TMA_NS], empty[TMA_NS];
    constexpr int T = TMA_T, NCW = TMA_T / 32;
    const int tid = threadIdx.x, L = a.nl, lane = tid & 31, warp = tid >> 5;
    int off[NDIAG];
#pragma unroll
    for (int d = 0; d < NDIAG; ++d) off[d] = a.A.off[d];
    const TmaLayout ly = tma_layout(off, L);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < TMA_NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NCW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int ntile = (a.hi - a.lo + T - 1) / T, G = gridDim.x;
    const int nq = (int)blockIdx.x < ntile ? (ntile - (int)blockIdx.x + G - 1) / G : 0;
    if (warp == NCW) {
        if (lane == 0)
            for (int q = 0; q < nq; ++q) {
                const unsigned int s = q % TMA_NS;
                mbar_wait(&empty[s], ((q / TMA_NS) & 1u) ^ 1u);
                tma_issue(a, ly, stg + (size_t)s * ly.total, &full[s], a.lo + ((int)blockIdx.x + q * G) * T, x, nullptr, false);
            }
        __syncwarp();
    } else {
        for (int q = 0; q < nq; ++q) {
            const unsigned int s = q % TMA_NS;
            mbar_wait(&full[s], (q / TMA_NS) & 1u);
            const int k = a.lo + ((int)blockIdx.x + q * G) * T + tid;
            if (k < a.hi) { double zc; y[k] = tma_row(ly, stg + (size_t)s * ly.total, tid, off, L, zc); }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
    }
}
