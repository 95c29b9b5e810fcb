/*
 * mc_oracle.h -- CPU oracle for the Monte Carlo pricing hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: the
 * only legal callers are tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs.  The product (montecarlocuda_b200/)
 * never links, imports or executes this code and has no CPU fallback.
 *
 * What it restates (plain C, libm, one thread), each function citing the
 * reference lines it follows (DP/ = /root/reference/double_precision/,
 * SP/ = /root/reference/single_precision/):
 *   - the three per-path estimators in their *device* form
 *       vanilla   DP/MonteCarloKernel.cu:67-71
 *       basket    DP/MonteCarloKernel.cu:74-101
 *       CVA       DP/MonteCarloKernel.cu:104-129, 234-262
 *     (+ switches reproducing the reference HOST variants: lagged spot,
 *      DP/MonteCarloHost.c:254-261, and the DP host basket vol bug,
 *      DP/MonteCarloHost.c:176-183),
 *   - the closing formulas         DP/MonteCarloKernel.cu:412-423, 459-469
 *   - Chol / cnd / host_bsCall     DP/MonteCarloHost.c:90-105, 124-143
 * on the Philox4x32-10 stream the GPU engine uses (the reference's XORWOW /
 * rand() streams cannot be shared between CPU and GPU).
 *
 * Parity pinning: the reference ships no golden vectors.  The
 * oracle is pinned against (i) outputs of the reference itself compiled here
 * into oracle/_ref/ (tests/golden/make_golden.py -> tests/golden/*.json):
 * host_bsCall and Chol bit-for-bit, the three host estimators statistically,
 * (ii) the Random123 Philox4x32-10 known-answer vectors, (iii) closed forms.
 */
#ifndef MC_ORACLE_H_
#define MC_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- stream definition (shared with the GPU engine) ---- */
enum { ORC_TAG_VANILLA = 1, ORC_TAG_BASKET = 2, ORC_TAG_CVA = 3 };
enum { ORC_THREADS = 256, ORC_LANES = 5 };

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* four words -> four normals (two Box-Muller pairs on bit-stuffed uniforms), either precision */
void orc_normals_f64(const uint32_t w[4], double z[4]);
void orc_normals_f32(const uint32_t w[4], float z[6]); /* three pairs per block in single precision */
/* the exact uniform stage, for bit-exact comparison with the device */
void orc_uniforms_f64(const uint32_t w[4], double f[4]); /* {radius f in [1,2), angle in turns} x 2 */
void orc_uniforms_f32(const uint32_t w[4], float f[6]); /* {radius, angle} stuffed floats in [1,2) x 3 */

/* ---- closed forms and helpers ---- */
double orc_cnd_hastings_f64(double d);
float  orc_cnd_hastings_f32(float d);
double orc_bs_call_hastings_f64(double s, double k, double r, double v, double t);
float  orc_bs_call_hastings_f32(float s, float k, float r, float v, float t);
double orc_bs_call_exact(double s, double k, double r, double v, double t);
void   orc_chol_f64(int n, const double *c, double *a);
void   orc_chol_f32(int n, const float *c, float *a);
/* time grid by repeated subtraction: tau[j-1] = remaining time
 * at date j, keep[j-1] = 1 when the reference evaluates the exposure there. */
void orc_cva_grid_f64(double T, int n, double *tau, int *keep);
void orc_cva_grid_f32(float T, int n, float *tau, int *keep);
/* E[CVA] of the device estimator / of the host (lagged) estimator in closed form */
double orc_cva_closed_form(double s, double k, double r, double v, double T, double lambda,
                           double lgd, int n, const int *keep, int lagged);

/* ---- per-path values on the Philox stream (undiscounted payoff / path CVA) ---- */
void orc_vanilla_payoffs_f64(double s, double k, double r, double v, double t, uint64_t seed,
                             uint64_t first_path, uint64_t n_paths, double *out);
void orc_vanilla_payoffs_f32(float s, float k, float r, float v, float t, uint64_t seed,
                             uint64_t first_path, uint64_t n_paths, float *out);
/* p = row-major n x n Cholesky factor (the caller factorises, DP/basketOpt.cu:96-99) */
void orc_basket_payoffs_f64(int n, const double *s, const double *v, const double *p,
                            const double *d, const double *w, double k, double t, double r,
                            int ref_dp_host_bug, uint64_t seed, uint64_t first_path,
                            uint64_t n_paths, double *out);
void orc_basket_payoffs_f32(int n, const float *s, const float *v, const float *p,
                            const float *d, const float *w, float k, float t, float r,
                            int ref_dp_host_bug, uint64_t seed, uint64_t first_path,
                            uint64_t n_paths, float *out);
void orc_cva_path_values_f64(double s, double k, double r, double v, double t, double lambda,
                             double lgd, int n_grid, int lagged_spot, uint64_t seed,
                             uint64_t first_path, uint64_t n_paths, double *out);
void orc_cva_path_values_f32(float s, float k, float r, float v, float t, float lambda,
                             float lgd, int n_grid, int lagged_spot, uint64_t seed,
                             uint64_t first_path, uint64_t n_paths, float *out);

/* ---- closing (DP/MonteCarloKernel.cu:412-423 pricing, :459-469 CVA) ---- */
void orc_closing(double sum, double sumsq, uint64_t n, double r, double t, int discount,
                 double *expected, double *confidence);

/* ---- the order-free combine: chunked partials -> exact integer lanes ---- */
/* units per thread per chunk for a job of total_units draw units */
int  orc_chunk_rounds(uint64_t total_units);
/* split a non-negative double, scaled by 2^scale_exp, into 5 x 32-bit limbs added to lanes */
int  orc_lanes_add(double value, int scale_exp, uint64_t lanes[ORC_LANES]);
/* lanes (after any integer summation) -> double, undoing the scale */
double orc_lanes_to_double(const uint64_t lanes[ORC_LANES], int scale_exp);
/* reduce per-path values of ONE chunk exactly as a 256-thread block does:
 * thread tid owns units k*256+tid (k < rounds) in order, then the xor-butterfly
 * over lanes of a warp, then warps 0..7 in order.  values_f64 holds the chunk's
 * paths (chunk-local index), n_valid the number of valid paths in it. */
void orc_chunk_reduce(const double *values, uint64_t n_valid, int unit_paths, int rounds,
                      int accumulate_in_float, double *sum, double *sumsq);

#ifdef __cplusplus
}
#endif
#endif
