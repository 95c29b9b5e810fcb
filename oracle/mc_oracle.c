/*
 * mc_oracle.c -- CPU oracle for the Monte Carlo pricing hot path (plain C, libm, one thread).
 * TEST INFRASTRUCTURE ONLY -- see mc_oracle.h for who may call it and what it restates.
 */
#include "mc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al., SC'11; Random123 convention, the one cuRAND's
 * curand_philox4x32_x.h uses: multipliers M0/M1, Weyl key increments W0/W1).
 * Pinned by the Random123 known-answer vectors in tests/golden/philox_kat.json.
 * ---------------------------------------------------------------------------------------- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; round++) {
        uint64_t p0 = (uint64_t)M0 * c0;
        uint64_t p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* ------------------------------------------------------------------------------------------
 * words -> uniforms -> normals.  The reference takes its normals from curand_normal()
 * (float Box-Muller on XORWOW words, DP/MonteCarloKernel.cu:68,78,250) or from rand()
 * (DP/MonteCarloHost.c:111-121); here the same Box-Muller map is applied to Philox words.
 * Uniforms are built by stuffing random bits under a fixed exponent: f in [1,2) exactly,
 * u1 = 2 - f in (0,1] for the radius, f - 1 (a turn fraction) for the angle.
 * ---------------------------------------------------------------------------------------- */
/* fp64: a block serves two pairs; pair i = (w[2i], w[2i+1]).  f[2i] = the radius uniform's
 * stuffed double in [1,2) (44 random bits: w[2i] and the top 12 bits of w[2i+1]), f[2i+1] = the
 * angle as a turn fraction k / 2^20 (the low 20 bits of w[2i+1]). */
void orc_uniforms_f64(const uint32_t w[4], double f[4])
{
    for (int i = 0; i < 2; i++) {
        uint32_t wa = w[2 * i], wb = w[2 * i + 1];
        uint64_t bits = ((uint64_t)(0x3ff00000u | (wa >> 12)) << 32) | (uint32_t)((wa << 20) | ((wb >> 12) & 0x000fff00u));
        memcpy(&f[2 * i], &bits, 8);
        f[2 * i + 1] = (double)(wb & 0x000fffffu) * 0x1p-20;
    }
}

/* fp32: a block serves THREE pairs.  The block is read as one 128-bit string w0:w1:w2:w3 (w0 most
 * significant); pair i takes 42 consecutive bits: 23 for the radius uniform (the whole mantissa of a float in
 * [1,2)), then 19 for the angle (mantissa bits [22:4] of a float in [1,2)); the last 2 bits are unused.
 * f[2i] = radius uniform, f[2i+1] = angle uniform, both stuffed floats in [1,2). */
void orc_uniforms_f32(const uint32_t w[4], float f[6])
{
    /* bit k of the string, k = 0 the most significant */
    for (int i = 0; i < 6; i++) {
        int off = 42 * (i / 2) + ((i & 1) ? 23 : 0);
        int len = (i & 1) ? 19 : 23;
        uint32_t field = 0;
        for (int b = 0; b < len; b++) {
            int k = off + b;
            uint32_t bit = (w[k / 32] >> (31 - k % 32)) & 1u;
            field = (field << 1) | bit;
        }
        uint32_t bits = 0x3f800000u | (field << (23 - len));
        memcpy(&f[i], &bits, 4);
    }
}

void orc_normals_f64(const uint32_t w[4], double z[4])
{
    double f[4];
    orc_uniforms_f64(w, f);
    for (int i = 0; i < 2; i++) {
        double rad = sqrt(-2.0 * log(2.0 - f[2 * i]));
        double ang = 6.283185307179586476925286766559 * f[2 * i + 1];
        z[2 * i] = rad * cos(ang);
        z[2 * i + 1] = rad * sin(ang);
    }
}

void orc_normals_f32(const uint32_t w[4], float z[6])
{
    float f[6];
    orc_uniforms_f32(w, f);
    for (int i = 0; i < 3; i++) {
        float rad = sqrtf(-2.0f * logf(2.0f - f[2 * i]));
        /* angle in [-pi, pi): 2*pi*(f - 1.5) as one fused multiply-add, like the device */
        float ang = fmaf(f[2 * i + 1], 6.283185307179586f, -9.42477796076938f);
        z[2 * i] = rad * cosf(ang);
        z[2 * i + 1] = rad * sinf(ang);
    }
}

/* ---- the two precision instances of the estimators ---- */
#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)

#define REAL double
#define FN(name) CAT(name, f64)
#define NPB 4
#define R_EXP exp
#define R_LOG log
#define R_SQRT sqrt
#define R_FABS fabs
#include "mc_oracle_real.inc"
#undef REAL
#undef FN
#undef NPB
#undef R_EXP
#undef R_LOG
#undef R_SQRT
#undef R_FABS

#define REAL float
#define FN(name) CAT(name, f32)
#define NPB 6
#define R_EXP expf
#define R_LOG logf
#define R_SQRT sqrtf
#define R_FABS fabsf
#include "mc_oracle_real.inc"
#undef REAL
#undef FN
#undef NPB
#undef R_EXP
#undef R_LOG
#undef R_SQRT
#undef R_FABS

/* Exact Black-Scholes call (erfc), the comparator for "within 3 SE of closed form". */
double orc_bs_call_exact(double s, double k, double r, double v, double t)
{
    double d1 = (log(s / k) + (r + 0.5 * v * v) * t) / (v * sqrt(t));
    double d2 = d1 - v * sqrt(t);
    double n1 = 0.5 * erfc(-d1 * 0.70710678118654752440);
    double n2 = 0.5 * erfc(-d2 * 0.70710678118654752440);
    return s * n1 - k * exp(-r * t) * n2;
}

/* Expectation of the CVA estimators: E[BS(S_t, T-t)] = C0(T) e^{rt}, so
 * device form : LGD * sum_{kept j} dp_j * C(S0,T)    * e^{r t_j}
 * host form   : LGD * sum_{kept j} dp_j * C(S0,T-dt) * e^{r t_{j-1}}   (lagged spot). */
double orc_cva_closed_form(double s, double k, double r, double v, double T, double lambda,
                           double lgd, int n, const int *keep, int lagged)
{
    double dt = T / n, acc = 0;
    double c0 = lagged ? orc_bs_call_exact(s, k, r, v, T - dt) : orc_bs_call_exact(s, k, r, v, T);
    for (int j = 1; j <= n; j++) {
        if (!keep[j - 1])
            continue;
        double dp = exp(-(dt * (j - 1)) * lambda) - exp(-(dt * j) * lambda);
        double tj = lagged ? dt * (j - 1) : dt * j;
        acc += dp * c0 * exp(r * tj);
    }
    return lgd * acc;
}

/* Closing: DP/MonteCarloKernel.cu:412-423 (pricing: discounted mean, half-width on the
 * UNdiscounted payoff) and :459-469 (CVA: no discount). */
void orc_closing(double sum, double sumsq, uint64_t n, double r, double t, int discount,
                 double *expected, double *confidence)
{
    double nn = (double)n;
    double price = sum / nn;
    if (discount)
        price = exp(-r * t) * price;
    double empstd = sqrt((nn * sumsq - sum * sum) / (nn * (nn - 1.0)));
    *confidence = 1.96 * empstd / sqrt(nn);
    *expected = price;
}

/* ------------------------------------------------------------------------------------------
 * Order-free combine.  The reference sums block partials sequentially on the host
 * (DP/MonteCarloKernel.cu:416-419): deterministic only for one launch shape.  The engine
 * instead turns every chunk partial into exact integer limbs, whose sum is independent of
 * grid shape, schedule and GPU count.  This is the CPU restatement used by the tests.
 * ---------------------------------------------------------------------------------------- */
int orc_chunk_rounds(uint64_t total_units)
{
    uint64_t q = total_units >> 24;
    int r = 1;
    while (r < 64 && (uint64_t)(r * 2) <= q)
        r *= 2;
    return r;
}

int orc_lanes_add(double value, int scale_exp, uint64_t lanes[ORC_LANES])
{
    double t = ldexp(value, scale_exp);
    double th = t * 0x1p-96;
    if (!(value >= 0.0) || !(th < 0x1p63))
        return 1; /* negative, NaN or outside the 160-bit window */
    uint64_t hi = (uint64_t)th;
    double rem = t - (double)hi * 0x1p96;
    uint64_t mid = (uint64_t)(rem * 0x1p-32);
    double lo_d = rem - (double)mid * 0x1p32;
    uint64_t lo = (uint64_t)lo_d;
    lanes[0] += lo;
    lanes[1] += mid & 0xffffffffu;
    lanes[2] += mid >> 32;
    lanes[3] += hi & 0xffffffffu;
    lanes[4] += hi >> 32;
    return 0;
}

double orc_lanes_to_double(const uint64_t lanes[ORC_LANES], int scale_exp)
{
    uint64_t limb[ORC_LANES + 1], carry = 0;
    for (int i = 0; i < ORC_LANES; i++) {
        uint64_t x = lanes[i] + carry;
        limb[i] = x & 0xffffffffu;
        carry = x >> 32;
    }
    limb[ORC_LANES] = carry;
    long double acc = 0;
    for (int i = ORC_LANES; i >= 0; i--)
        acc = acc * 4294967296.0L + (long double)limb[i];
    return (double)ldexpl(acc, -scale_exp);
}

void orc_chunk_reduce(const double *values, uint64_t n_valid, int unit_paths, int rounds,
                      int accumulate_in_float, double *sum, double *sumsq)
{
    double ts[ORC_THREADS], ts2[ORC_THREADS];
    for (int tid = 0; tid < ORC_THREADS; tid++) {
        double s = 0, s2 = 0;
        /* fp32 with an even number of paths per draw unit: the even and the odd paths of every unit have running sums
         * of their own, joined by one float addition at the end of the chunk (the engine accumulates two paths per
         * packed instruction); otherwise one running sum */
        float fs[2] = {0, 0}, fs2[2] = {0, 0};
        int interleaved = (unit_paths & 1) == 0;
        for (int k = 0; k < rounds; k++) {
            uint64_t unit = (uint64_t)k * ORC_THREADS + (uint64_t)tid;
            for (int q = 0; q < unit_paths; q++) {
                uint64_t idx = unit * (uint64_t)unit_paths + (uint64_t)q;
                if (idx >= n_valid)
                    continue;
                if (accumulate_in_float) {
                    float x = (float)values[idx];
                    int which = interleaved ? (q & 1) : 0;
                    fs[which] += x;
                    fs2[which] = fmaf(x, x, fs2[which]);
                } else {
                    double x = values[idx];
                    s += x;
                    s2 = fma(x, x, s2);
                }
            }
        }
        ts[tid] = accumulate_in_float ? (double)(float)(fs[0] + fs[1]) : s;
        ts2[tid] = accumulate_in_float ? (double)(float)(fs2[0] + fs2[1]) : s2;
    }
    /* xor butterfly inside each warp (offsets 16, 8, 4, 2, 1): every lane ends with the
     * same value, so only lane 0 is tracked */
    double ws[ORC_THREADS / 32], ws2[ORC_THREADS / 32];
    for (int wi = 0; wi < ORC_THREADS / 32; wi++) {
        double a[32], b[32], na[32], nb[32];
        memcpy(a, ts + wi * 32, sizeof a);
        memcpy(b, ts2 + wi * 32, sizeof b);
        for (int off = 16; off >= 1; off >>= 1) {
            for (int l = 0; l < 32; l++) {
                na[l] = a[l] + a[l ^ off];
                nb[l] = b[l] + b[l ^ off];
            }
            memcpy(a, na, sizeof a);
            memcpy(b, nb, sizeof b);
        }
        ws[wi] = a[0];
        ws2[wi] = b[0];
    }
    double S = ws[0], S2 = ws2[0];
    for (int wi = 1; wi < ORC_THREADS / 32; wi++) {
        S += ws[wi];
        S2 += ws2[wi];
    }
    *sum = S;
    *sumsq = S2;
}
