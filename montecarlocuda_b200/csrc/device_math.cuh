// device_math.cuh -- in-register uniforms, Box-Muller normals and the scalar special functions
// of the pricing kernels, in both working precisions (sm_100a).
//
// Replaces curand_normal() (float Box-Muller on XORWOW words even in the reference's DP tree,
// DP/MonteCarloKernel.cu:68,78,250) and the libm calls of callPayoff / basketPayoff /
// geomBrownian / cnd / device_bsCall (DP/MonteCarloKernel.cu:67-129).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "device_math64.cuh"

namespace mcb {

// Per-CTA shared-memory state of a workload: nothing for fp32 (every special function is a MUFU
// op), the log / exp / angle tables for fp64.  load() is called by all threads of the CTA before the
// first barrier of the kernel.
struct NoShared {
    __device__ __forceinline__ void load() {}
};
struct SharedTables64 {
    Tables64 t;
    __device__ __forceinline__ void load()
    {
        for (int i = threadIdx.x; i < 256; i += blockDim.x)
            t.fill(i);
    }
};
// the same tables in the bank-conflict-free layout (device_math64.cuh), for the pricing kernels
struct SharedTables64Rep {
    Tables64Rep t;
    __device__ __forceinline__ void load()
    {
        for (int i = threadIdx.x; i < kLogEntries * 8; i += blockDim.x) {
            const int j = i >> 3, rep = i & 7;
            t.log_rep[j][rep][0] = bias_log_recip(kLogTable[j][0]);
            t.log_rep[j][rep][1] = kLogTable[j][1];
        }
        for (int i = threadIdx.x; i < 256 * 16; i += blockDim.x) {
            const int j = i >> 4, rep = i & 15;
            t.rows[j].exp[rep] = bias_exp_entry(kExpTable[j], j);
            t.rows[j].turn_lo[rep >> 1][rep & 1] = kTurnLoTable[j][rep & 1];
        }
        for (int i = threadIdx.x; i < 4096; i += blockDim.x)
            *reinterpret_cast<double2 *>(t.turn_hi[i]) = *reinterpret_cast<const double2 *>(kTurnHiTable[i]);
    }
};
template <typename Real> struct SharedFor;
template <> struct SharedFor<float> { using type = NoShared; };
template <> struct SharedFor<double> { using type = SharedTables64; };
// what mc_accumulate_kernel instantiates a workload with
template <typename Real> struct SharedAccumFor { using type = typename SharedFor<Real>::type; };
template <> struct SharedAccumFor<double> { using type = SharedTables64Rep; };

// Per-JOB shared-memory state, one per sub-block of 256 threads (the sub-blocks of a sweep launch may be working on
// different jobs): nothing for fp32; for fp64 the exponent table of the job's scaled logarithm (LogScale64).
// W::prepare(P, job_state, tid) fills it; the kernel puts a sub-block barrier behind it.
struct NoJobState {
    __device__ __forceinline__ void fill(int, float) {}
};
template <typename Real> struct JobStateFor;
template <> struct JobStateFor<float> { using type = NoJobState; };
template <> struct JobStateFor<double> { using type = LogScale64; };

// ---- fp32: every transcendental is ONE MUFU op (no denormal fix-up, no range-reduction code) ----
__device__ __forceinline__ float mufu_lg2(float x)
{
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float mufu_ex2(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float mufu_sqrt(float x)
{
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float mufu_rcp(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float mufu_sin(float x)
{
    float y;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float mufu_cos(float x)
{
    float y;
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Packed fp32 arithmetic (sm_100: add / mul / fma .f32x2 -- FADD2, FMUL2, FFMA2): two independent IEEE operations in one
// issue slot.  The fp32 kernels are bound by issue slots around their MUFU chain, so wherever two
// scalar operations of the same kind sit side by side with no shared operand to duplicate, they go out as one.
// Results are bit-identical to the scalar instructions.
#ifndef MCB_PACKED_F32
#define MCB_PACKED_F32 1
#endif
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi)
{
    unsigned long long d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ void unpack_f32x2(unsigned long long v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// (d0, d1) = (a0, a1) * (b, b) + (c, c)
__device__ __forceinline__ void fma_f32x2(float a0, float a1, float b, float c, float &d0, float &d1)
{
#if MCB_PACKED_F32
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pack_f32x2(a0, a1)), "l"(pack_f32x2(b, b)), "l"(pack_f32x2(c, c)));
    unpack_f32x2(d, d0, d1);
#else
    d0 = fmaf(a0, b, c);
    d1 = fmaf(a1, b, c);
#endif
}
// (d0, d1) = (a0, a1) * (b, b)
__device__ __forceinline__ void mul_f32x2(float a0, float a1, float b, float &d0, float &d1)
{
#if MCB_PACKED_F32
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pack_f32x2(a0, a1)), "l"(pack_f32x2(b, b)));
    unpack_f32x2(d, d0, d1);
#else
    d0 = a0 * b;
    d1 = a1 * b;
#endif
}
// (d0, d1) = (a0, a1) - (b0, b1)
__device__ __forceinline__ void sub_f32x2(float a0, float a1, float b0, float b1, float &d0, float &d1)
{
#if MCB_PACKED_F32
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pack_f32x2(a0, a1)), "l"(pack_f32x2(b0, b1)));
    unpack_f32x2(d, d0, d1);
#else
    d0 = a0 - b0;
    d1 = a1 - b1;
#endif
}
// (d0, d1) = (a0, a1) + (c, c)
__device__ __forceinline__ void add_f32x2(float a0, float a1, float c, float &d0, float &d1)
{
#if MCB_PACKED_F32
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pack_f32x2(a0, a1)), "l"(pack_f32x2(c, c)));
    unpack_f32x2(d, d0, d1);
#else
    d0 = a0 + c;
    d1 = a1 + c;
#endif
}

// fp32: one Philox block (128 bits) serves THREE Box-Muller pairs, i.e. six normals.  Pair i takes 42
// consecutive bits of the block read as one 128-bit string w0:w1:w2:w3 (w0 most significant): 23 for the
// radius uniform, then 19 for the angle; the last 2 bits are unused.
//   radius: the 23 bits become the mantissa of f in [1, 2) exactly; u = 2 - f in (0, 1] (tails to 5.65 sigma)
//   angle:  the 19 bits become mantissa bits [22:4] of f in [1, 2): 2^19 directions (for anything smooth in the
//           pair the lattice error is that of a trapezoid rule on a periodic function, i.e. below fp32 rounding)
// The first version gave every uniform the top 23 bits of its own word (two pairs per block): a third more
// IMAD.WIDE work -- the instruction that binds the fp32 kernels -- for bits no estimate can see.
// Field extraction is a funnel shift + one LOP3 ((x & mask) | exponent) per uniform, 11 ALU ops per block.
__device__ __forceinline__ float stuffed_unit_f32(uint32_t w)  // 23 random bits from the top of a word (one LEA.HI)
{
    return __uint_as_float(0x3f800000u | (w >> 9));
}
__device__ __forceinline__ float stuff_f32(uint32_t window, uint32_t mask)
{
    return __uint_as_float((window & mask) | 0x3f800000u);
}
__device__ __forceinline__ void uniforms_f32(const uint32_t (&w)[4], float (&f)[6])
{
    constexpr uint32_t kRadius = 0x007fffffu, kAngle = 0x007ffff0u;
    f[0] = stuffed_unit_f32(w[0]);                                    // bits   0.. 22: w0[31:9]
    f[1] = stuff_f32(__funnelshift_l(w[1], w[0], 14), kAngle);        // bits  23.. 41: w0[8:0] w1[31:22]
    f[2] = stuff_f32(__funnelshift_l(w[2], w[1], 1), kRadius);        // bits  42.. 64: w1[21:0] w2[31]
    f[3] = stuff_f32(w[2] >> 8, kAngle);                              // bits  65.. 83: w2[30:12]
    f[4] = stuff_f32(__funnelshift_l(w[3], w[2], 11), kRadius);       // bits  84..106: w2[11:0] w3[31:21]
    f[5] = stuff_f32(w[3] << 2, kAngle);                              // bits 107..125: w3[20:2]
}

// One Box-Muller pair from its two stuffed uniforms: radius u = 2 - f in (0, 1], r = sqrt(-2 ln u) = sqrt(-2 ln2 lg2 u);
// angle 2 pi (f - 1.5) in [-pi, pi) as one FFMA, the range where MUFU.SIN/COS are most accurate.  4 MUFU per pair (LG2,
// SQRT, SIN, COS).
// (r, cos, sin) of a block's three pairs: r = sqrt(c lg2(u)) -- c = -2 ln 2 for a standard normal pair, times b^2 for
// a caller that wants b r (polar_from_words below).  Pairs 0 and 1 go side by side on packed instructions.
__device__ __forceinline__ void polar_f32(const uint32_t (&w)[4], float c, float (&r)[3], float (&cs)[3], float (&sn)[3])
{
    float f[6];
    uniforms_f32(w, f);
    float u[3], ang[3], l[3];
    fma_f32x2(f[0], f[2], -1.0f, 2.0f, u[0], u[1]);                    // 2 - f exactly, as the scalar subtraction
    u[2] = 2.0f - f[4];
    fma_f32x2(f[1], f[3], 6.283185307179586f, -9.42477796076938f, ang[0], ang[1]);
    ang[2] = fmaf(f[5], 6.283185307179586f, -9.42477796076938f);
    mul_f32x2(mufu_lg2(u[0]), mufu_lg2(u[1]), c, l[0], l[1]);
    l[2] = mufu_lg2(u[2]) * c;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        r[i] = mufu_sqrt(l[i]);
        cs[i] = mufu_cos(ang[i]);
        sn[i] = mufu_sin(ang[i]);
    }
}

__device__ __forceinline__ void normals_from_words(const uint32_t (&w)[4], float (&z)[6], const NoShared &)
{
    float r[3], cs[3], sn[3];
    polar_f32(w, -1.3862943611198906f, r, cs, sn);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        z[2 * i] = r[i] * cs[i];
        z[2 * i + 1] = r[i] * sn[i];
    }
}

// ---- fp64 ----
// One Philox block (128 bits) serves TWO fp64 Box-Muller pairs, i.e. four normals.
// Each pair takes 64 bits (wa, wb): the radius uniform gets 44 of them (wa and the top 12 bits of wb,
// stuffed under the exponent of 1.0: f in [1,2), u = 2 - f in (0,1], tails to 7.8 sigma), the angle
// the other 20 (a turn fraction k / 2^20: 2^20 directions; for anything smooth in the pair the
// lattice error is that of a trapezoid rule on a periodic function, i.e. far below fp64 rounding).
// The first version spent a whole block per pair (52-bit radius and angle): twice the IMAD.WIDE
// work for bits no estimate can see.  All arithmetic after the bits is fp64: 1 (2 - f) + 9 (log) + 5 or 7 (sqrt) +
// 4 (cos/sin from the two-level table) [+ 2 for z = r cos, r sin] fp64 instructions per pair (libdevice: 68+).
// No |.| around the logarithm: u = 2 - f is a multiple of 2^-44, so k ln u is either the tiny positive offset (u == 1)
// or at least 2^-43 |k| / 2 -- four orders of magnitude above the function's rounding error (checked on the host over
// every u = 1 - j 2^-44, j <= 2e6: min 1.1366e-13, error 2.7e-17).  With a 52-bit radius it could come out as -1e-17,
// and the fabs cost an fp64 instruction per pair (MUFU.RSQ64H cannot take an operand modifier).
__device__ __forceinline__ double radius_uniform_f64(uint32_t wa, uint32_t wb)
{
    const double f = __hiloint2double((int)(0x3ff00000u | (wa >> 12)), (int)((wa << 20) | ((wb >> 12) & 0x000fff00u)));
    return 2.0 - f;
}
template <class Tab>
__device__ __forceinline__ void box_muller_f64(uint32_t wa, uint32_t wb, double &z0, double &z1, const Tab &T, const LogScale64 &S)
{
    const double r = sqrt_pos<true>(neg2log_unit(radius_uniform_f64(wa, wb), T, S));
    double cs, sn;
    sincos_turn20(wb, cs, sn, T);
    z0 = r * cs;
    z1 = r * sn;
}

// S: LogScale64 filled with -2 ln 2 (polar_scale(1))
template <class Sh>
__device__ __forceinline__ void normals_from_words(const uint32_t (&w)[4], double (&z)[4], const Sh &sh, const LogScale64 &S)
{
    box_muller_f64(w[0], w[1], z[0], z[1], sh.t, S);
    box_muller_f64(w[2], w[3], z[2], z[3], sh.t, S);
}
__device__ __forceinline__ void normals_from_words(const uint32_t (&w)[4], float (&z)[6], const NoShared &sh, const NoJobState &)
{
    normals_from_words(w, z, sh);
}

// ---- scaled polar form of a block's Box-Muller pairs ---------------------------------------------------
// A caller that only needs b z (a diffusion step, an exponent), never z itself, takes the pairs as
// (b r, cos, sin): b z0 = (b r) cos, b z1 = (b r) sin fold into the FMA that consumes them, and the scale b
// costs nothing because it rides on constants that are there anyway -- fp32: b r = sqrt(lg2(u) * c) with
// c = -2 ln2 b^2; fp64: b^2 (-2 ln u) = scaled_log_unit(u, c, S) with c = -2 b^2 and the job's exponent table S
// filled with c ln 2.  Saves the two multiplies r cos, r sin of every pair.
template <typename Real> struct PolarScale {
    Real c, c_ln2;
};
template <typename Real> __host__ __device__ inline PolarScale<Real> polar_scale(double b)
{
    PolarScale<Real> s;
    if (sizeof(Real) == 4) {
        s.c = (Real)(-1.3862943611198906188 * b * b);
        s.c_ln2 = 0;
    } else {
        s.c = (Real)(-2.0 * b * b);
        s.c_ln2 = (Real)(-2.0 * b * b * 0.69314718055994530942);
    }
    return s;
}
// the per-job state of a workload whose normals are scaled by S (W::prepare)
__device__ __forceinline__ void prepare_polar(const PolarScale<float> &, NoJobState &, int) {}
__device__ __forceinline__ void prepare_polar(const PolarScale<double> &S, LogScale64 &job, int tid)
{
    if (tid < 64)
        job.fill(tid, S.c_ln2);
}
template <bool kShortSqrt = true>
__device__ __forceinline__ void polar_from_words(const uint32_t (&w)[4], float (&br)[3], float (&cs)[3], float (&sn)[3],
                                                 const NoShared &, const PolarScale<float> &S, const NoJobState &)
{
    polar_f32(w, S.c, br, cs, sn);
}
template <bool kShortSqrt = true, class Sh>
__device__ __forceinline__ void polar_from_words(const uint32_t (&w)[4], double (&br)[2], double (&cs)[2], double (&sn)[2],
                                                 const Sh &sh, const PolarScale<double> &S, const LogScale64 &job)
{
#pragma unroll
    for (int i = 0; i < 2; i++) {
        br[i] = sqrt_pos<kShortSqrt>(scaled_log_unit(radius_uniform_f64(w[2 * i], w[2 * i + 1]), sh.t, S.c, job));
        sincos_turn20(w[2 * i + 1], cs[i], sn[i], sh.t);
    }
}

// normals per Philox block
template <typename Real> struct NormalsPerBlock;
template <> struct NormalsPerBlock<float> { static constexpr int value = 6; };
template <> struct NormalsPerBlock<double> { static constexpr int value = 4; };

// max(x, 0) for fp32: one FMNMX (the fp64 kernels leave the clamp to add_value, device_common.cuh)
__device__ __forceinline__ float positive_part(float x) { return fmaxf(x, 0.0f); }

// precision-generic wrappers used by the workload policies.  exp_scaled takes its argument in the units the
// exponential is cheapest in -- log2 units for fp32 (MUFU.EX2), units of ln2/256 for fp64 (exp_units) -- and
// kExpUnit<Real> is the factor that takes a natural-log quantity there (the host folds it into the job's constants).
template <typename Real> struct ExpUnit;
template <> struct ExpUnit<float> { static constexpr double value = 1.4426950408889634074; };           // 1 / ln 2
template <> struct ExpUnit<double> { static constexpr double value = 369.32993046757463228; };           // 256 / ln 2
template <bool kLateTable = false, class Sh> __device__ __forceinline__ float exp_scaled(float x, const Sh &) { return mufu_ex2(x); }
template <bool kLateTable = false, class Sh> __device__ __forceinline__ double exp_scaled(double y, const Sh &sh) { return exp_units<kLateTable>(y, sh.t); }
__device__ __forceinline__ float rcp_real(float x) { return mufu_rcp(x); }
__device__ __forceinline__ double rcp_real(double x) { return rcp_newton(x); }

}  // namespace mcb
