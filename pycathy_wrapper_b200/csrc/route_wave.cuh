// route_wave.cuh -- ROUTE + ALTEZZE (SRC/route.f:47-253, SRC/altezze.f, SRC/mc.f) as a WAVEFRONT over (drainage level, sub-step).
// Included by cathy_b200.cu (shares RouteArgs, mc_cell).
//
// The reference sweeps the cells in descending-elevation order, NSURF sub-steps one after the other.  k_route runs that as
// NSURF x NLEVEL strictly dependent level steps on one CTA (config 3 bench DEM: 398 levels x ~8 sub-steps = 3,200 barriers, 1.5 ms).
// But cell c at sub-step s only needs (i) the outflows of its donors -- cells of LOWER levels -- at the same sub-step and (ii) its own
// state at sub-step s-1.  So all tasks (level l, sub-step s) with l + s = w are independent: NLEVEL + NSURF - 1 wavefronts instead of
// NLEVEL x NSURF steps, each wavefront holding the cells of up to NSURF levels (several hundred to a few thousand tasks).
// A task costs four fp64 pow() (Muskingum-Cunge celerity and diffusivity of two directions): on ONE SM the sweep is bound by the fp64
// pipe (ncu: k_route 2.1 ms on the bench DEM for ~40 k cells x ~10 sub-steps), so the wavefront runs on a THREAD-BLOCK CLUSTER of up to
// 16 CTAs (16 SMs) that meet at the hardware cluster barrier (barrier.cluster, ~0.2 us) after every wavefront;
//   * the outflows are kept per sub-step (qo[dir][s][pos], a receiver may sit many levels below its donor), a cell's own inflow and
//     volume in a 2-deep ring;
//   * per-cell constants live in ONE 128-byte record in level order (one coalesced line per task, L1-resident for the following
//     sub-steps of the same cell); the records of the level that enters the wavefront next are prefetched into L1;
//   * the arithmetic of a task is route_cell / ALTEZZE of k_route operation for operation and the donor sum keeps the reference's
//     order; results equal the sequential sweep to rounding (config 1, 1,693 sub-steps: same accepted steps, heads 2e-15 m apart --
//     the compiler contracts the split Muskingum-Cunge formula into other multiply-adds).
// With ONE sub-step per call there is nothing to overlap (wavefronts = levels) and k_route's level-ahead register prefetch is the
// faster sweep (5.3 against 10 us per level of 200 cells): the host launches the wavefront only when the previous call needed two
// or more sub-steps.  More than ROUTE_NSMAX sub-steps (or no memory for the history): the launch leaves a flag and k_route does the
// step instead.  Both kernels give the same results to rounding, so switching between them from call to call is safe.
#define ROUTE_NSMAX 64
struct __align__(16) RouteS {        // 128 bytes, level order
    double w[2], epl[2], ckf[2], dhd[2], nrc, b1, y1;
    int ib, seq, nd, d0;             // routing index I_BASIN, position in QOI order, donors, first overflow donor entry
    int dc[4];                       // first four donors: (level-order position << 1) | direction
    double pad;
};
struct RouteWArgs {
    RouteArgs r;
    const RouteS *rs;                // [ncell] level order
    const int *dcx;                  // overflow donor codes (donors beyond the fourth), indexed d0 + j
    double *qo;                      // [2][ROUTE_NSMAX][ncell] outflows per direction and sub-step, level order
    double *qin_ring, *vol_ring;     // [2][ncell]
    int *handled;                    // out: 1 = this launch did the routing step
    double *best;                    // [3 x cluster size] per-CTA (Courant number, celerity, sequence) of the last sub-step
    int nsmax;
    unsigned long long *prof;        // diagnostic (CATHY_ROUTE_DEBUG): globaltimer at the end of wavefront w, w < 1024
};
__global__ void k_route_fill_static(int ncell, RouteS *rs, const double *ckf1, const double *ckf2, const double *dhd1, const double *dhd2)
{
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < ncell; p += gridDim.x * blockDim.x) {
        const int ib = rs[p].ib;
        rs[p].ckf[0] = ckf1[ib]; rs[p].ckf[1] = ckf2[ib]; rs[p].dhd[0] = dhd1[ib]; rs[p].dhd[1] = dhd2[ib];
    }
}
// mc_cell with its two pow() passed in: the four powers of a cell (two directions x celerity / diffusivity) are evaluated by FOUR lanes
// side by side -- a double-precision pow() is a dependent chain of ~1 us, and four of them one after the other were the 5 us per
// drainage level that bound the routing sweep; operation for operation the rest is mc_cell.
__device__ __forceinline__ double mc_qc(double q_in_kk, double q_in_kkp1, double q_out_kk)
{
    double qc = 1.0 / 3.0 * (q_in_kk + q_in_kkp1 + q_out_kk);
    if (qc <= 1.0e-05) qc = 1.0e-05;
    return qc;
}
__device__ __forceinline__ double mc_finish(double ckf, double dhd, double epl, double dt, double p_ck, double p_dh, double q_in_kk, double q_in_kkp1,
                                            double q_out_kk, double q_over, double &cu, double &ak)
{
    double ck = ckf * p_ck;
    ak = ck / epl;
    cu = ck * dt / epl;
    double dh = p_dh / dhd;
    if (dh < (1.0 - cu)) dh = 1.0 - cu;
    double xx = 0.50 - dh / (ck * epl);
    double den = 2.0 * (1.0 - xx) + cu;
    double c1 = (cu - 2.0 * xx) / den, c2 = (cu + 2.0 * xx) / den, c3 = (2.0 * (1.0 - xx) - cu) / den, c4 = (2.0 * ck * dt) / den;
    return c1 * q_in_kkp1 + c2 * q_in_kk + c3 * q_out_kk + c4 * q_over;
}
constexpr int ROUTE_WBLOCK = 512;
__global__ void __launch_bounds__(ROUTE_WBLOCK) k_route_wave(RouteWArgs A)
{
    const RouteArgs &a = A.r;
    cg::cluster_group cl = cg::this_cluster();
    const int crank = (int)cl.block_rank(), ncta = (int)cl.num_blocks();
    __shared__ double s_cu[32], s_ak[32];
    __shared__ int s_seq[32];
    __shared__ int s_nsurf;
    __shared__ double s_dt;
    if (threadIdx.x == 0) {
        double akm = *a.ak_max, cu_max = akm * a.deltat, dts;
        int ns;
        if (cu_max > 1.0) { dts = 1.0 / akm; ns = (int)(a.deltat / dts) + 1; dts = a.deltat / ns; }
        else { dts = a.deltat; ns = 1; }
        s_nsurf = ns; s_dt = dts;
        if (crank == 0) *A.handled = ns <= A.nsmax ? 1 : 0;
    }
    __syncthreads();
    const int nsurf = s_nsurf;
    if (nsurf > A.nsmax) return;                     // k_route takes the step
    const double dt = s_dt;
    const int nc = a.ncell, nlev = a.nlevel;
    const int *__restrict__ lp = a.level_ptr;
    const size_t slab = (size_t)A.nsmax * nc;
    const bool multi = ncta > 1;
    double best_cu = -1.0, best_ak = 0.0;
    int best_seq = -1;
    const int quad = threadIdx.x & 3, qbase = (threadIdx.x & 31) & ~3;           // four lanes per task
    const int per_round = ncta * (int)(blockDim.x >> 2);
    // dynamic data written by another SM of the cluster one wavefront ago must be read past the L1
    auto ldd = [&](const double *p) { return multi ? __ldcg(p) : *p; };
    for (int w = 0; w < nlev + nsurf - 1; ++w) {
        const int s_lo = max(0, w - nlev + 1), s_hi = min(nsurf - 1, w);
        int ntask = 0;
        for (int s = s_lo; s <= s_hi; ++s) ntask += lp[w - s + 1] - lp[w - s];
        if (!multi && w + 1 < nlev)                  // the records of the level that enters the wavefront next -> L1
            for (int q = lp[w + 1] + threadIdx.x; q < lp[w + 2]; q += blockDim.x) asm volatile("prefetch.global.L1 [%0];" ::"l"(A.rs + q));
        for (int base = 0; base < ntask; base += per_round) {
            // task -> (sub-step s, position pos): the levels w - s_lo, w - s_lo - 1, ... one after the other
            const int task = base + crank * (int)(blockDim.x >> 2) + (int)(threadIdx.x >> 2);
            int rem = task, s = s_lo, pos = -1;
            if (task < ntask)
                for (; s <= s_hi; ++s) {
                    const int l = w - s, cnt = lp[l + 1] - lp[l];
                    if (rem < cnt) { pos = lp[l] + rem; break; }
                    rem -= cnt;
                }
            const bool act = pos >= 0;
            const int dir = quad >> 1;
            double qik = 0.0, qok[2] = {0.0, 0.0}, vkk = 0.0, swsn = 0.0, qin = 0.0, pw = 0.0;
            const RouteS *S = A.rs + (act ? pos : 0);            // 128-byte record
            int ib = 0;
            if (act) {
                ib = S->ib;
                if (s == 0) { qik = a.q_in_kk[ib]; qok[0] = a.q_out_kk_1[ib]; qok[1] = a.q_out_kk_2[ib]; vkk = a.volume_kk[ib]; }
                else {
                    qik = ldd(A.qin_ring + (size_t)((s - 1) & 1) * nc + pos); vkk = ldd(A.vol_ring + (size_t)((s - 1) & 1) * nc + pos);
                    qok[0] = ldd(A.qo + (size_t)(s - 1) * nc + pos); qok[1] = ldd(A.qo + slab + (size_t)(s - 1) * nc + pos);
                }
                swsn = a.sw_sn[ib];
                // inflow: the donors' outflows of THIS sub-step, summed in the reference's order
                const int nd = S->nd;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j < nd) { const int code = S->dc[j]; qin = qin + ldd(A.qo + (size_t)(code & 1) * slab + (size_t)s * nc + (code >> 1)); }
                for (int j = 4; j < nd; ++j) { const int code = A.dcx[S->d0 + j - 4]; qin = qin + ldd(A.qo + (size_t)(code & 1) * slab + (size_t)s * nc + (code >> 1)); }
                // this lane's power: direction quad >> 1, celerity exponent (quad & 1 == 0) or diffusivity exponent
                const double wd = S->w[dir];
                if (wd != 0.0) {
                    const double nrc = S->nrc, b1 = S->b1, g = (1.0 - S->y1 + 2.0 / 3.0 * b1);
                    const double qc = mc_qc(qik * wd / nrc, qin * wd / nrc, qok[dir] / nrc);
                    pw = pow(qc, (quad & 1) ? 1.0 - b1 : 1.0 - 3.0 * g / 5.0);
                }
            }
            const double p00 = __shfl_sync(0xffffffffu, pw, qbase), p01 = __shfl_sync(0xffffffffu, pw, qbase + 1);
            const double p10 = __shfl_sync(0xffffffffu, pw, qbase + 2), p11 = __shfl_sync(0xffffffffu, pw, qbase + 3);
            if (act && quad == 0) {
                const double nrc = S->nrc, swv = swsn / nrc;
                double qo[2] = {0.0, 0.0};
#pragma unroll
                for (int d = 0; d < 2; ++d) {
                    const double wd = S->w[d];
                    if (wd == 0.0) continue;
                    const double epl = S->epl[d];
                    double q_over = swv * wd * (1.0 / epl);
                    double q_in_kk = qik * wd / nrc, q_out_kk = qok[d] / nrc;
                    double q_in_kkp1 = qin * wd / nrc, cu, ak;
                    double q = mc_finish(S->ckf[d], S->dhd[d], epl, dt, d ? p10 : p00, d ? p11 : p01, q_in_kk, q_in_kkp1, q_out_kk, q_over, cu, ak);
                    if (q < 0.0) q = 0.0;
                    qo[d] = q * nrc;
                    if (s == nsurf - 1) {
                        const int sq = 2 * S->seq + d;
                        if (cu > best_cu || (cu == best_cu && sq > best_seq)) { best_cu = cu; best_ak = ak; best_seq = sq; }
                    }
                }
                A.qo[(size_t)s * nc + pos] = qo[0]; A.qo[slab + (size_t)s * nc + pos] = qo[1];
                A.qin_ring[(size_t)(s & 1) * nc + pos] = qin;
                // ALTEZZE: volume balance and water depth of the cell
                const double dv = (qik + qin) / 2 * dt + swsn * dt - (qok[0] + qok[1]) / 2 * dt - (qo[0] + qo[1]) / 2 * dt;
                double v1 = vkk + dv, h;
                if (v1 >= 0.0) h = v1 / a.cellarea; else { v1 = 0.0; h = 0.0; }
                A.vol_ring[(size_t)(s & 1) * nc + pos] = v1;
                if (s == nsurf - 1) {
                    a.q_in_kkp1[ib] = qin; a.q_out_kkp1_1[ib] = qo[0]; a.q_out_kkp1_2[ib] = qo[1]; a.volume_kkp1[ib] = v1; a.h_water[ib] = h;
                }
                if (nsurf > 1 && s == nsurf - 2) {   // what the reference's time-level shift leaves in the KK arrays (:171-195)
                    a.q_in_kk[ib] = qin; a.q_out_kk_1[ib] = qo[0]; a.q_out_kk_2[ib] = qo[1]; a.volume_kk[ib] = v1;
                }
            }
        }
        if (multi) cl.sync(); else __syncthreads();  // barrier.cluster: release / acquire over the SMs of the cluster
        if (A.prof && crank == 0 && threadIdx.x == 0 && w < 1024) { unsigned long long t_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_)); A.prof[w] = t_; }
    }
    // AK_MAX = celerity of the LAST cell (in sequential order) attaining the max Courant number of the last sub-step
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double oc = __shfl_down_sync(0xffffffffu, best_cu, o), oa = __shfl_down_sync(0xffffffffu, best_ak, o);
        int os = __shfl_down_sync(0xffffffffu, best_seq, o);
        if (oc > best_cu || (oc == best_cu && os > best_seq)) { best_cu = oc; best_ak = oa; best_seq = os; }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { s_cu[wid] = best_cu; s_ak[wid] = best_ak; s_seq[wid] = best_seq; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < (int)(blockDim.x >> 5); ++q)
            if (s_cu[q] > best_cu || (s_cu[q] == best_cu && s_seq[q] > best_seq)) { best_cu = s_cu[q]; best_ak = s_ak[q]; best_seq = s_seq[q]; }
        A.best[3 * crank] = best_cu; A.best[3 * crank + 1] = best_ak; A.best[3 * crank + 2] = (double)best_seq;
    }
    cl.sync();
    if (crank == 0 && threadIdx.x == 0) {
        best_cu = -1.0; best_seq = -1;
        for (int q = 0; q < ncta; ++q) {
            const double c_ = __ldcg(A.best + 3 * q), k_ = __ldcg(A.best + 3 * q + 1);
            const int s_ = (int)__ldcg(A.best + 3 * q + 2);
            if (c_ > best_cu || (c_ == best_cu && s_ > best_seq)) { best_cu = c_; best_ak = k_; best_seq = s_; }
        }
        if (best_seq >= 0) *a.ak_max = best_ak;
        *a.nsurf_out = nsurf;
    }
}
