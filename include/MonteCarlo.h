/*
 * MonteCarlo.h -- drop-in data model for the mcb200 Monte Carlo pricing engine.
 *
 * Same struct names, field names, field order and field types as the reference header
 * (/root/reference/double_precision/MonteCarlo.h:32-73 and single_precision/MonteCarlo.h:34-74),
 * so a caller compiled against the reference header links against libmcb200_dp / libmcb200_sp
 * unchanged.  Differences, all additive:
 *   - one header serves both precisions: define MCB200_SINGLE before including it to get the
 *     single_precision layout (the reference keeps two copies that differ only in the type);
 *   - N, the basket width, is wrapped in #ifndef so -DN=10 works (the reference's bare
 *     `#define N 3`, MonteCarlo.h:16, silently overrides a command-line N);
 *   - CudaCheck only exists when the CUDA runtime header was included first.
 */
#ifndef MONTECARLO_H_
#define MONTECARLO_H_

#include <stdlib.h>
#include <stdio.h>
#include <math.h>
#include <time.h>

#ifndef N
#define N 3
#endif

#ifdef MCB200_SINGLE
typedef float mc_real;
#else
typedef double mc_real;
#endif

#if !defined(CudaCheck) && defined(__CUDA_RUNTIME_H__)
#define CudaCheck(value)                                                                  \
    {                                                                                     \
        cudaError_t _m_cudaStat = value;                                                  \
        if (_m_cudaStat != cudaSuccess) {                                                 \
            fprintf(stderr, "Error %s at line %d in file %s\n",                           \
                    cudaGetErrorString(_m_cudaStat), __LINE__, __FILE__);                 \
            exit(1);                                                                      \
        }                                                                                 \
    }
#endif

/* European option: spot, strike, risk-free rate, volatility, maturity */
typedef struct {
    mc_real s;
    mc_real k;
    mc_real r;
    mc_real v;
    mc_real t;
} OptionData;

/* Basket of N underlyings.  p is row-major and must hold the lower-triangular Cholesky factor
 * of the correlation matrix when it reaches dev_basketOpt / host_basketOpt (the reference
 * driver overwrites it, basketOpt.cu:96-99). */
typedef struct {
    mc_real s[N];
    mc_real v[N];
    mc_real p[N][N];
    mc_real d[N];
    mc_real w[N];
    mc_real k;
    mc_real t;
    mc_real r;
} MultiOptionData;

/* Expected = price (discounted) or CVA; Confidence = 1.96 * stdev / sqrt(n) of the
 * UNdiscounted per-path value, exactly as the reference computes it
 * (MonteCarloKernel.cu:420-423). */
typedef struct {
    mc_real Expected;
    mc_real Confidence;
} OptionValue;

/* CVA of one call: flat default intensity, loss given default, ns (unused by the reference),
 * the option, n = number of exposure dates. */
typedef struct {
    mc_real defInt, lgd;
    int ns;
    OptionData option;
    int n;
} CVA;

typedef struct {
    OptionValue callValue;
    MultiOptionData mopt;
    OptionData sopt;
    int numOpt, path;
} MonteCarloData;

#ifdef __cplusplus
extern "C" {
#endif

/* ---- the drop-in boundary: GPU estimators (reference MonteCarloKernel.cu:483,500,517) ----
 * numBlocks / numThreads were the reference's launch shape; they no longer choose the
 * geometry, but n = numBlocks * (sims / numBlocks) paths are simulated exactly as before.
 * Any CUDA failure prints a message and exit(1)s, as CudaCheck does in the reference. */
OptionValue dev_vanillaOpt(OptionData *opt, int numBlocks, int numThreads, int sims);
OptionValue dev_basketOpt(MultiOptionData *option, int numBlocks, int numThreads, int sims);
OptionValue dev_cvaEquityOption(CVA *cva, int numBlocks, int numThreads, int sims);

/* ---- host helpers (reference MonteCarloHost.c:20-143, 282-311), libmcb200_hostapi_{dp,sp}[_nN].so ----
 * CPU-side companions the reference drivers import next to the GPU entry points.  They are NOT on the
 * GPU pricing path and the dev_* functions never fall back to them (see csrc/hostapi.c). */
void printVect(mc_real *mat, int c);
void printMat(mc_real *mat, int r, int c);
void printOption(OptionData o);
void printMultiOpt(MultiOptionData *o);
void prodMat(mc_real *first, mc_real *second, mc_real *result, int f_rows, int f_cols, int s_cols);
void Chol(mc_real c[N][N], mc_real a[N][N]);
mc_real randMinMax(mc_real min, mc_real max);
mc_real host_bsCall(OptionData option);
OptionValue host_vanillaOpt(OptionData option, int path);
OptionValue host_basketOpt(MultiOptionData *option, int path);
OptionValue host_cvaEquityOption(CVA *cva, int path);
/* runtime-width Cholesky that REPORTS a non-positive pivot (returns its 1-based index, 0 = success)
 * instead of silently zeroing the column as Chol does (MonteCarloHost.c:100-101) */
int mcb200_chol(int n, const double *c, double *a);

#ifdef __cplusplus
}
#endif

#endif /* MONTECARLO_H_ */
