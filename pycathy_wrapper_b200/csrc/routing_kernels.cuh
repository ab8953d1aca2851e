// routing_kernels.cuh -- surface routing (SURF_FLOWTRA): node <-> cell transfers, Muskingum-Cunge static factors, the level-scheduled sweeps k_route / k_route4 and the wavefront kernel (route_wave.cuh).
// Part of the single translation unit cathy_b200.cu (included in dependency order; shares its structs and helpers).
#pragma once

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
