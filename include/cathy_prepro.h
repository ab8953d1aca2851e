/* cathy_prepro.h -- C ABI of the B200-native CATHY pre-processor (part of libcathy_b200.so).
 *
 * Boundary being replaced: pyCATHY runs the Fortran pre-processor as a child process `./pycppp` in
 * <project>/prepro with "2\n0\n1\n" on stdin (pyCATHY/cathy_tools.py:378-389); it reads hap.in + dtm_13.val and
 * writes the rasters dem, lakes_map, zone, dtm_* and the cell order qoi_a that the processor reads when ISIMGR=2
 * (SRC/datin.f:325-372).  PROGRAM CPPP: PRE/cppp.f90:20-86, PRE = examples/SSHydro/weill_exemple/prepro/src.
 * There is no FFI in the reference; this is the interface a binding would target.  Text I/O (hap.in parsing and its
 * rewrite, raster formats of PRE/mrbb_sr.f90) stays on the host side (pycathy_wrapper_b200/preprocessor.py);
 * everything between -- CSORT, DEPIT, CCA, SMEAN, DSF, HG -- runs on the device.
 *
 * Cell numbering is the reference's i_basin = (i-1)*M + j, i = 1..N west to east, j = 1..M south to north
 * (PRE/wbb_sr.f90:9-17); arrays crossing this boundary are 0-based copies of it: element [i_basin-1].
 */
#ifndef CATHY_PREPRO_H
#define CATHY_PREPRO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CATHY_PREPRO_ABI_VERSION 1

/* hap.in (PRE/mpar.f90:28-68).  The reference rewrites hap.in after WBB (PRE/cppp.f90:27, mpar.f90:400-541) and CCA
 * re-reads the ROUNDED values (PRE/cca.f90:31), so the stages before CCA use the values as the user wrote them
 * (suffix 0) and all later stages the re-read ones. */
typedef struct CathyPreproParams {
    int32_t abi_version;
    int32_t N, M;              /* DEM rectangle: columns (x), rows (y)                                   */
    int32_t imethod;           /* 1 = D8-LAD, 2 = D8-LTD (deviations; PRE/dsf.f90:250-259)               */
    int32_t ndcf;              /* non-dispersive channel flow                                            */
    int32_t nchc;              /* channel initiation: 1 = A, 2 = A*S**k (3 = normalised divergence: refused) */
    int32_t p_outflow_vo;      /* drainage direction of the outlet cell if none can be derived           */
    int32_t bcc;               /* boundary channel construction (PRE/wbb_sr.f90:95-160)                  */
    double delta_x0, pt0;      /* as written by the user: DEPIT's eps = pt0*delta_x0 (PRE/depit.f90:39)  */
    float cqm0, cqg0;          /* boundary-channel coefficients (used by WBB only)                       */
    double delta_x, delta_y;   /* re-read values                                                         */
    double lambda;             /* upstream deviation memory factor                                       */
    double A_threshold;
    float CC_threshold, ASk_threshold, kas, _pad0;
    double dr;
    double As_rf, As_cf;
    float Qsf_rf, w_rf, Wsf_rf, b1_rf, b2_rf, kSsf_rf, y1_rf, y2_rf;
    float Qsf_cf, w_cf, Wsf_cf, b1_cf, b2_cf, kSsf_cf, y1_cf, y2_cf;
} CathyPreproParams;

/* Caller-owned host arrays of N*M elements each (order: n_cells elements are filled). */
typedef struct CathyPreproOut {
    double *quota;             /* cell elevations after boundary channel + DEPIT (file `dem`)            */
    double *A_inflow;          /* upstream drainage area (dtm_A_inflow)                                  */
    float *w_1, *w_2;          /* weights of the cardinal / diagonal direction                           */
    float *local_slope_1, *local_slope_2, *epl_1, *epl_2;
    float *Ws1_sf_1, *Ws1_sf_2, *b1_sf, *kSs1_sf_1, *kSs1_sf_2, *y1_sf, *nrc;
    int32_t *p_outflow_1, *p_outflow_2, *hcID, *dmID;
    int32_t *order;            /* qoi: i_basin (1-based) in descending elevation, the reference's quicksort tie order */
    int32_t n_cells;           /* N_celle                                                                */
    int32_t n_modifications;   /* DEPIT's total                                                          */
    int32_t n_waves;           /* dependency wavefronts the DSF sweep needed                             */
    int32_t n_launches;        /* kernels launched                                                       */
    double mean_s_max;         /* SMEAN                                                                  */
    double device_ms;          /* CUDA-event time of the whole device part                               */
    double stage_ms[8];        /* csort, pit check, DEPIT, 2nd csort, window analysis + SMEAN, DSF sweep,
                                  outlet + HG (ms); [7] = DEPIT sweeps                                   */
} CathyPreproOut;

/* CSORT + DEPIT + CSORT + CCA + SMEAN + DSF + HG on cuda device `device`.
 * quota_in[N*M]: elevations, present[N*M]: 1 = catchment cell (dtm_13.val value > -9999, PRE/wbb_sr.f90:76).
 * Returns 0, or a negative code with a text in cathy_prepro_last_error():
 *  -1 bad arguments / unsupported option (nchc = 3, a raster one cell wide: both undefined in the reference), -2 "catchment with more than one outlet cell!" (PRE/depit.f90:56-61),
 *  -4 non-positive elevation inside the catchment (the reference uses 0 and negative values as "no cell" marks,
 *     PRE/dsf.f90:87-99), -5 boundary-channel check of PRE/wbb_sr.f90:147-158, -100 CUDA error. */
int32_t cathy_prepro_run(const CathyPreproParams *p, const double *quota_in, const uint8_t *present, int32_t device,
                         CathyPreproOut *out);

const char *cathy_prepro_last_error(void);

/* Host-side text of the raster files (RBB, PRE/mrbb_sr.f90:445-470): nrows records of ncols fields in Fortran Ew.d
 * (kind 0) or Fw.d (kind 1), resp. Iw, each record ending in a newline, formatted row-parallel on `nthreads` host
 * threads (0 = up to 32).  out holds nrows * (ncols * w + 1) bytes; returns the bytes written, -1 if a value needs a
 * form these routines do not write (NaN, infinity, three-digit exponent). */
int64_t cathy_prepro_format_real(const double *v, int64_t nrows, int64_t ncols, int32_t w, int32_t d, int32_t kind, char *out,
                                 int32_t nthreads);
int64_t cathy_prepro_format_int(const int32_t *v, int64_t nrows, int64_t ncols, int32_t w, char *out, int32_t nthreads);

#ifdef __cplusplus
}
#endif
#endif
