/*
 * hostapi.c -- the reference's CPU helper API as first-class exports.
 *
 * The reference drivers import, besides the three dev_* GPU entry points, a set of host helpers from
 * MonteCarloHost.c: Chol, host_bsCall, printOption, printMultiOpt, randMinMax, prodMat and the
 * single-threaded CPU estimators host_vanillaOpt / host_basketOpt / host_cvaEquityOption that the
 * drivers time next to the GPU to print a speed-up (double_precision/vanillaOpt.cu:61-72,
 * basketOpt.cu:96-106).  This file provides them, written from scratch, in a library of its own
 * (libmcb200_hostapi_{dp,sp}[_nN].so) so that the reference drivers link with no reference object
 * at all.  NOTHING here is on the GPU pricing path: libmcb200*.so never calls into this library
 * and there is no fallback from the dev_* entry points to these CPU estimators.
 *
 * Same signatures and struct layouts as the reference (MonteCarloHost.c:20-143, 282-311).
 * Deliberate differences, all documented in DESIGN.md 9:
 *   - the estimators follow the reference's DEVICE formulas: the basket keeps the volatility in the
 *     diffusion (the DP host drops it, MonteCarloHost.c:180) and the CVA exposure uses the new spot
 *     (the host uses the previous one, :254-261);
 *   - random numbers: the engine's Philox4x32-10 stream instead of rand() seeded by the wall clock
 *     (:111-121,189), so a CPU estimate is reproducible and walks the SAME paths as the GPU estimate
 *     for the same seed (MCB200_SEED, default as in dropin.cpp);
 *   - sums are accumulated in double also in the single-precision build (:188 saturates at 2^25 paths);
 *   - mcb200_chol() reports a non-positive pivot instead of silently zeroing the column (:100-101);
 *     Chol() keeps the reference's silent behaviour for drop-in compatibility.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/MonteCarlo.h"

#ifdef MCB200_SINGLE
#define R_EXP expf
#define R_LOG logf
#define R_SQRT sqrtf
#define R_FABS fabsf
#else
#define R_EXP exp
#define R_LOG log
#define R_SQRT sqrt
#define R_FABS fabs
#endif
#ifdef MCB200_SINGLE
#define NORMALS_PER_BLOCK 6 /* one Philox block = three single-precision Box-Muller pairs (23-bit radius, 19-bit angle) */
#else
#define NORMALS_PER_BLOCK 4 /* ... or two double-precision pairs (44-bit radius, 20-bit angle) */
#endif

/* ---- printing (MonteCarloHost.c:20-65): same fields, same order ---- */
void printVect(mc_real *mat, int c)
{
    printf("\n!\t");
    for (int j = 0; j < c; j++)
        printf("\t%f\t", (double)mat[j]);
    printf("\t!\n\n");
}

void printMat(mc_real *mat, int r, int c)
{
    for (int i = 0; i < r; i++) {
        printf("\n!\t");
        for (int j = 0; j < c; j++)
            printf("\t%f\t", (double)mat[j + i * c]);
        printf("\t!");
    }
    printf("\n\n");
}

void printOption(OptionData o)
{
    printf("\n-\tOption data\t-\n\n");
    printf("Underlying asset price:\t \xe2\x82\xac %.2f\n", (double)o.s);
    printf("Strike price:\t\t \xe2\x82\xac %.2f\n", (double)o.k);
    printf("Risk free interest rate: %.2f %%\n", (double)o.r * 100);
    printf("Volatility:\t\t %.2f %%\n", (double)o.v * 100);
    printf("Time to maturity:\t %.2f %s\n", (double)o.t, (o.t > 1) ? "years" : "year");
}

void printMultiOpt(MultiOptionData *o)
{
    printf("\n-\tBasket Option data\t-\n\n");
    printf("Number of assets: %d\n", N);
    printf("Underlying assets prices:\n");
    printVect(o->s, N);
    printf("Volatility:\n");
    printVect(o->v, N);
    printf("Weights:");
    printVect(o->w, N);
    printf("Correlation matrix:\n");
    printMat(&o->p[0][0], N, N);
    printf("Strike price:\t\t \xe2\x82\xac %.2f\n", (double)o->k);
    printf("Risk free interest rate: %.2f \n", (double)o->r);
    printf("Time to maturity:\t %.2f %s\n", (double)o->t, (o->t > 1) ? "years" : "year");
}

/* ---- small linear algebra (MonteCarloHost.c:67-105) ---- */
void prodMat(mc_real *first, mc_real *second, mc_real *result, int f_rows, int f_cols, int s_cols)
{
    for (int i = 0; i < f_rows; i++)
        for (int j = 0; j < s_cols; j++) {
            mc_real acc = 0;
            for (int k = 0; k < f_cols; k++)
                acc += first[k + i * f_cols] * second[j + k * s_cols];
            result[j + i * s_cols] = acc;
        }
}

/* Column Cholesky, c = a a^T, a lower triangular.  Returns 0 on success, j + 1 when pivot j is not
 * positive (c not positive definite); the factor is then valid for the leading j x j block only. */
int mcb200_chol(int n, const double *c, double *a)
{
    int bad = 0;
    for (int i = 0; i < n * n; i++)
        a[i] = 0;
    for (int j = 0; j < n; j++) {
        double pivot = c[j * n + j];
        for (int k = 0; k < j; k++)
            pivot -= a[j * n + k] * a[j * n + k];
        if (!(pivot > 0)) {
            if (!bad)
                bad = j + 1;
            continue; /* leave the column zero, like the reference */
        }
        const double root = sqrt(pivot);
        a[j * n + j] = root;
        for (int i = j + 1; i < n; i++) {
            double v = c[i * n + j];
            for (int k = 0; k < j; k++)
                v -= a[j * n + k] * a[i * n + k];
            a[i * n + j] = v / root;
        }
    }
    return bad;
}

/* The reference entry point: same arithmetic order as MonteCarloHost.c:90-105 (the tests compare it
 * bit for bit with the reference's own output), silent on a non-positive pivot. */
void Chol(mc_real c[N][N], mc_real a[N][N])
{
    mc_real v[N];
    for (int j = 0; j < N; j++)
        for (int i = 0; i < N; i++) {
            a[i][j] = 0;
            if (i >= j) {
                v[i] = c[i][j];
                for (int k = 0; k < j; k++)
                    v[i] -= a[j][k] * a[i][k];
                if (v[j] > 0)
                    a[i][j] = v[i] / R_SQRT(v[j]);
            }
        }
}

/* ---- finance helpers (MonteCarloHost.c:111-143) ---- */
mc_real randMinMax(mc_real min, mc_real max)
{
    mc_real x = (mc_real)rand() / (mc_real)(RAND_MAX);
    return max * x + ((mc_real)1.0 - x) * min;
}

/* Hastings / Abramowitz-Stegun 26.2.17, the reference's normal CDF */
static mc_real cnd(mc_real d)
{
    const mc_real b1 = (mc_real)0.31938153, b2 = (mc_real)-0.356563782, b3 = (mc_real)1.781477937;
    const mc_real b4 = (mc_real)-1.821255978, b5 = (mc_real)1.330274429;
    const mc_real inv_sqrt_2pi = (mc_real)0.39894228040143267793994605993438;
    const mc_real k = (mc_real)1.0 / ((mc_real)1.0 + (mc_real)0.2316419 * R_FABS(d));
    mc_real tail = inv_sqrt_2pi * R_EXP((mc_real)-0.5 * d * d) * (k * (b1 + k * (b2 + k * (b3 + k * (b4 + k * b5)))));
    return d > 0 ? (mc_real)1.0 - tail : tail;
}

mc_real host_bsCall(OptionData option)
{
    const mc_real vol_t = option.v * R_SQRT(option.t);
    const mc_real d1 = (R_LOG(option.s / option.k) + (option.r + (mc_real)0.5 * option.v * option.v) * option.t) / vol_t;
    const mc_real d2 = d1 - vol_t;
    return option.s * cnd(d1) - option.k * R_EXP(-option.r * option.t) * cnd(d2);
}

/* ---- the engine's random stream on the host: Philox4x32-10, bit-stuffed uniforms, Box-Muller ---- */
static void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint64_t seed, uint32_t out[4])
{
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int i = 0; i < 10; i++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static void normals(const uint32_t w[4], mc_real z[NORMALS_PER_BLOCK])
{
#ifdef MCB200_SINGLE
    /* the block as one 128-bit string w0:w1:w2:w3; pair i = 23 radius bits, then 19 angle bits (csrc/device_math.cuh) */
    const uint32_t field[6] = {
        w[0] >> 9,
        ((w[0] << 14) | (w[1] >> 18)) & 0x007ffff0u,
        ((w[1] << 1) | (w[2] >> 31)) & 0x007fffffu,
        (w[2] >> 8) & 0x007ffff0u,
        ((w[2] << 11) | (w[3] >> 21)) & 0x007fffffu,
        (w[3] << 2) & 0x007ffff0u,
    };
    for (int i = 0; i < 3; i++) {
        uint32_t b0 = 0x3f800000u | field[2 * i], b1 = 0x3f800000u | field[2 * i + 1];
        float f0, f1;
        memcpy(&f0, &b0, 4);
        memcpy(&f1, &b1, 4);
        const float rad = sqrtf(-2.0f * logf(2.0f - f0));
        const float ang = fmaf(f1, 6.283185307179586f, -9.42477796076938f);
        z[2 * i] = rad * cosf(ang);
        z[2 * i + 1] = rad * sinf(ang);
    }
#else
    for (int i = 0; i < 2; i++) {
        const uint32_t wa = w[2 * i], wb = w[2 * i + 1];
        const uint64_t bits = ((uint64_t)(0x3ff00000u | (wa >> 12)) << 32) | (uint32_t)((wa << 20) | ((wb >> 12) & 0x000fff00u));
        double f;
        memcpy(&f, &bits, 8);
        const double rad = sqrt(-2.0 * log(2.0 - f));
        const double ang = 6.283185307179586476925286766559 * ((double)(wb & 0x000fffffu) * 0x1p-20);
        z[2 * i] = rad * cos(ang);
        z[2 * i + 1] = rad * sin(ang);
    }
#endif
}

static uint64_t stream_seed(void)
{
    const char *s = getenv("MCB200_SEED");
    return s ? strtoull(s, NULL, 0) : 0x6d63623230300001ull;
}

static OptionValue closing(double sum, double sumsq, int n, double discount)
{
    /* MonteCarloHost.c:220-228 / MonteCarloKernel.cu:420-423: discounted mean, half-width of the
     * undiscounted value */
    OptionValue v;
    const double nn = (double)n;
    const double var = (nn * sumsq - sum * sum) / (nn * (nn - 1.0));
    v.Expected = (mc_real)(discount * sum / nn);
    v.Confidence = (mc_real)(1.96 * sqrt(var > 0 ? var : 0) / sqrt(nn));
    return v;
}

/* ---- CPU estimators (MonteCarloHost.c:185-311), single thread ---- */
OptionValue host_vanillaOpt(OptionData option, int path)
{
    const uint64_t seed = stream_seed();
    const mc_real drift = (option.r - (mc_real)0.5 * option.v * option.v) * option.t;
    const mc_real vol = option.v * R_SQRT(option.t);
    double sum = 0, sumsq = 0;
    for (int first = 0; first < path; first += NORMALS_PER_BLOCK) {
        uint32_t w[4];
        mc_real z[NORMALS_PER_BLOCK];
        const uint64_t unit = (uint64_t)first / NORMALS_PER_BLOCK;
        philox((uint32_t)unit, (uint32_t)(unit >> 32), 0u, 1u, seed, w);
        normals(w, z);
        for (int q = 0; q < NORMALS_PER_BLOCK && first + q < path; q++) {
            const mc_real pay = option.s * R_EXP(drift + vol * z[q]) - option.k;
            const double p = pay > 0 ? (double)pay : 0.0;
            sum += p;
            sumsq += p * p;
        }
    }
    return closing(sum, sumsq, path, exp(-(double)option.r * (double)option.t));
}

OptionValue host_basketOpt(MultiOptionData *option, int path)
{
    const uint64_t seed = stream_seed();
    const mc_real sqrt_t = R_SQRT(option->t);
    double sum = 0, sumsq = 0;
    for (int i = 0; i < path; i++) {
        mc_real g[N + NORMALS_PER_BLOCK];
        for (int jb = 0; jb * NORMALS_PER_BLOCK < N; jb++) {
            uint32_t w[4];
            philox((uint32_t)i, 0u, (uint32_t)jb, 2u, seed, w);
            normals(w, g + jb * NORMALS_PER_BLOCK);
        }
        mc_real basket = 0;
        for (int a = 0; a < N; a++) {
            mc_real bt = option->d[a];
            for (int b = 0; b < N; b++)
                bt += option->p[a][b] * g[b];
            basket += option->w[a] * option->s[a] *
                      R_EXP((option->r - (mc_real)0.5 * option->v[a] * option->v[a]) * option->t + option->v[a] * bt * sqrt_t);
        }
        const double p = basket > option->k ? (double)(basket - option->k) : 0.0;
        sum += p;
        sumsq += p * p;
    }
    return closing(sum, sumsq, path, exp(-(double)option->r * (double)option->t));
}

OptionValue host_cvaEquityOption(CVA *cva, int path)
{
    const uint64_t seed = stream_seed();
    const OptionData o = cva->option;
    const int n = cva->n;
    const mc_real dt = o.t / n;
    const mc_real drift = (o.r - (mc_real)0.5 * o.v * o.v) * dt, vol = o.v * R_SQRT(dt);
    double sum = 0, sumsq = 0;
    for (int i = 0; i < path; i++) {
        OptionData cur = o;
        mc_real acc = 0;
        mc_real z[NORMALS_PER_BLOCK];
        for (int j = 1; j <= n; j++) {
            if ((j - 1) % NORMALS_PER_BLOCK == 0) {
                uint32_t w[4];
                philox((uint32_t)i, 0u, (uint32_t)((j - 1) / NORMALS_PER_BLOCK), 3u, seed, w);
                normals(w, z);
            }
            const mc_real dp = R_EXP(-(dt * (j - 1)) * cva->defInt) - R_EXP(-(dt * j) * cva->defInt);
            mc_real ee = 0;
            if ((cur.t -= dt) >= 0) { /* the reference's time grid: repeated subtraction decides the last date */
                cur.s = cur.s * R_EXP(drift + vol * z[(j - 1) % NORMALS_PER_BLOCK]);
                ee = host_bsCall(cur);
            }
            acc += dp * ee;
        }
        const double p = (double)(acc * cva->lgd);
        sum += p;
        sumsq += p * p;
    }
    return closing(sum, sumsq, path, 1.0);
}
