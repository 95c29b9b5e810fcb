// device_math64.cuh -- the fp64 special functions of the pricing kernels, hand-built for the
// sm_100a pipe mix: the fp64 pipe retires 64 thread-instr/clk/SM and there is no fp64 MUFU, so
// libdevice's log / sincospi / exp / division cost 30 / 27 / 18 / 8+ fp64 instructions plus as many
// integer ones (measured: 162 warp-instructions per vanilla path, profiles/r01a_*).  Here:
//
//   neg2log_unit   -2 ln(u), u in (0,1]   table of 256 reciprocals (shared memory) + degree-6 log1p   9 fp64
//   sqrt_pos       sqrt(x)                MUFU.RSQ64H seed + 2 coupled Newton steps (or 1 + a correction) 7 / 5 fp64
//   sincos_turn    cos/sin(2 pi k/2^52)   octant taken from the integer bits (exact reduction),
//                                         fdlibm kernel polynomials on [0, pi/4]                     18 fp64
//   exp_tab        e^x                    n = rint(256 x / ln2) by magic add, Cody-Waite, table of
//                                         2^(j/256), degree-4 expm1, exponent added as an integer     9 fp64
//   rcp_newton     1/x                    MUFU.RCP64H seed + 1 cubic step                             3 fp64
//
// Accuracy (checked on the CPU against libm by tests/test_device_math64.py through the host build
// of this very header, and on the GPU against the oracle): <= 2 ulp for exp/sqrt/rcp/sincos,
// <= 2.5e-16 absolute on -2 ln u.
//
// The header compiles both for the device and, with MCB_HOST_MATH defined, for the host (MUFU
// seeds emulated at their documented precision), which is how the accuracy tests run without a GPU.
#pragma once

#include <cstdint>
#include <cstring>

#ifdef MCB_HOST_MATH
#include <cmath>
#define MCB_FN static inline
#define MCB_MEMBER inline
#define MCB_TABLE static const
#else
#include <cuda_runtime.h>
#define MCB_FN __device__ __forceinline__
#define MCB_MEMBER __device__ __forceinline__
#define MCB_TABLE __device__ const
#endif

namespace mcb {
#ifdef MCB_HOST_MATH
namespace hostmath {
#endif

#include "tables64.inc"

// Polynomial and reduction constants that do not fit an instruction's immediate field (an fp64 immediate is the
// high word only).  Written as literals, ptxas re-materialises each use with two MOVs -- in the register-capped CVA
// kernel that was 38 of the 167 instructions of a path-step (profiles/r01k_cva50_f64_2p26.txt: UMOV 20, IMAD.MOV 12,
// LDC 6 per step).  From the constant bank they are plain DFMA operands (CVA 20.4 -> 19.8 ms).  Kernels with
// registers to spare keep the literals: there ptxas holds them in registers across the loop and the constant-bank
// operands are slower (vanilla fp64 9.72 -> 10.36 ms when forced), so the choice rides on the table type
// (Tab::kConstBank) the kernel instantiates the functions with.  Since the CVA kernel runs 2 sub-blocks of 256 threads
// (128 registers) it keeps the literals too (16.92 vs 17.01 ms); the constant-bank route stays selectable
// (MCB_CVA_BANK=1 together with MCB_CVA_SUBBLOCKS=3, kernels_cva.cu).
struct MathConsts64 {
    double log_magic;          // 2^52 + 1023
    double log_c6, log_c5, log_c3;   // -1/6, 1/5, 1/3   (-1/4, -1/2 are immediates)
    double neg2ln2;            // -2 ln 2
    double tiny;               // 1e-300
    double exp_scale;          // 256 / ln 2
    double exp_ln2_hi, exp_ln2_lo;   // -ln2/256 split (Cody-Waite)
    double exp_c4, exp_c3;     // 1/24, 1/6
    double inv_sqrt_2pi;
    double hast_k, hast_a1, hast_a2, hast_a3, hast_a4, hast_a5;
};
#define MCB_MATH_CONSTS_INIT                                                                                             \
    {4503599627371519.0, -1.0 / 6.0, 0.2, 1.0 / 3.0, -2.0 * 0x1.62e42fefa39efp-1, 1e-300, 0x1.71547652b82fep+8,           \
     -0x1.62e42fee00000p-9, -0x1.a39ef35793c76p-41, 1.0 / 24.0, 1.0 / 6.0, 0.39894228040143267793994605993438, 0.2316419,   \
     0.31938153, -0.356563782, 1.781477937, -1.821255978, 1.330274429}
#ifdef MCB_HOST_MATH
static const MathConsts64 kMathConsts64 = MCB_MATH_CONSTS_INIT;
#else
static __constant__ MathConsts64 kMathConsts64 = MCB_MATH_CONSTS_INIT;
#endif

// ---- bit access and the two MUFU seeds -----------------------------------------------------------
#ifdef MCB_HOST_MATH
MCB_FN int hi_word(double x) { uint64_t b; std::memcpy(&b, &x, 8); return (int)(b >> 32); }
MCB_FN int lo_word(double x) { uint64_t b; std::memcpy(&b, &x, 8); return (int)(uint32_t)b; }
MCB_FN double make_double(int hi, int lo)
{
    uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double x;
    std::memcpy(&x, &b, 8);
    return x;
}
MCB_FN double fma_(double a, double b, double c) { return std::fma(a, b, c); }
// MUFU.RSQ64H / MUFU.RCP64H work on the high word and deliver ~2^-22 relative error (PTX ISA:
// rsqrt.approx.ftz.f64, rcp.approx.ftz.f64); emulate by truncating the input to its high word
// and the result to its high word (20 mantissa bits: a little worse than the hardware)
MCB_FN double seed_trunc(double y) { return make_double(hi_word(y), 0); }
MCB_FN double rsqrt_seed(double x) { return seed_trunc(1.0 / std::sqrt(make_double(hi_word(x), 0))); }
MCB_FN double rcp_seed(double x) { return seed_trunc(1.0 / make_double(hi_word(x), 0)); }
#else
MCB_FN int hi_word(double x) { return __double2hiint(x); }
MCB_FN int lo_word(double x) { return __double2loint(x); }
MCB_FN double make_double(int hi, int lo) { return __hiloint2double(hi, lo); }
MCB_FN double fma_(double a, double b, double c) { return fma(a, b, c); }
MCB_FN double rsqrt_seed(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}
MCB_FN double rcp_seed(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}
#endif

// The tables live in shared memory inside the kernels (random per-thread indices: a constant-bank
// read would serialise); this is the view the functions take.  Plain layout: host build, instrumentation
// and per-path kernels.
#define MCB_K(Tab, field, literal) (Tab::kConstBank ? kMathConsts64.field : (literal))
struct Tables64 {
    static constexpr bool kConstBank = false;
    double log_tab[256][2];       // { c_i, -ln c_i }
    double exp_tab[256];          // 2^(j/256)
    double turn_hi[1024][2];      // { cos, sin } of 2 pi i / 1024
    double turn_lo[1024][2];      // { cos, sin } of 2 pi j / 2^20
    MCB_MEMBER void log_entry(int i, double &c, double &l) const { c = log_tab[i][0]; l = log_tab[i][1]; }
    MCB_MEMBER double exp_entry(int j) const { return exp_tab[j]; }
    MCB_MEMBER void turn_hi_entry(uint32_t i, double &c, double &s) const { c = turn_hi[i][0]; s = turn_hi[i][1]; }
    MCB_MEMBER void turn_lo_entry(uint32_t j, double &c, double &s) const { c = turn_lo[j][0]; s = turn_lo[j][1]; }
};

#ifndef MCB_HOST_MATH
// Bank-conflict-free layout for the pricing kernels.  The indices are random per thread, so in the plain layout
// the 8 threads of a quarter-warp (16-byte loads) or the 16 of a half-warp (8-byte loads) collide in the 32
// shared-memory banks: ~2.7 wavefronts where 1 would do.  ncu on the plain layout (profiles/r01j_vanilla_f64_2p32.txt):
// 62 % of all shared-memory wavefronts were bank conflicts and the shared-memory pipe was 97 % busy -- the
// European-call fp64 kernel was bound by it, not by the fp64 pipe.  Here the two small tables are replicated once
// per bank group: thread t reads replica t % 8 (16-byte entries: 128 bytes per index, one bank group per replica)
// or t % 16 (8-byte entries), so threads that are served together can never share a bank.  Same values, same
// arithmetic, bit-identical results; 64 KB instead of 6 KB, shared by all warps of a (large) CTA.
// The two angle tables (16 KB each) stay single copies: 2.6 wavefronts per quarter-warp instead of 1 (22 of the 29
// wavefronts of a Box-Muller pair).  Four copies each (thread t reads replica t % 4: 1.9 wavefronts per quarter-warp,
// 192 KB in all) were measured and are SLOWER -- European call 10.15 vs 9.73 ms, basket-10 7.26 vs 7.21 ms
// (profiles/r01p_ab_experiments.txt): once log and exp were conflict-free the shared-memory pipe stopped being what
// binds, and the wider index arithmetic costs more than the wavefronts give back.
struct Tables64Rep {
    static constexpr bool kConstBank = false;
    double log_rep[256][8][2];    // [index][replica]{ c_i, -ln c_i }
    double exp_rep[256][16];      // [index][replica] 2^(j/256)
    double turn_hi[1024][2];
    double turn_lo[1024][2];
    MCB_MEMBER void log_entry(int i, double &c, double &l) const
    {
        const double2 v = *reinterpret_cast<const double2 *>(&log_rep[i][threadIdx.x & 7][0]);
        c = v.x;
        l = v.y;
    }
    MCB_MEMBER double exp_entry(int j) const { return exp_rep[j][threadIdx.x & 15]; }
    MCB_MEMBER void turn_hi_entry(uint32_t i, double &c, double &s) const { c = turn_hi[i][0]; s = turn_hi[i][1]; }
    MCB_MEMBER void turn_lo_entry(uint32_t j, double &c, double &s) const { c = turn_lo[j][0]; s = turn_lo[j][1]; }
};
struct Tables64RepBank : Tables64Rep {   // same layout; the functions take their constants from the constant bank
    static constexpr bool kConstBank = true;
};
#endif

// ---- cos and sin of 2 pi k / 2^20 for a 20-bit integer k (the Box-Muller angle of the kernels) ---
// Two-level table: k = 1024 i + j, angle = coarse_i + fine_j, and the addition theorems give the
// result from four correctly rounded table values with 2 multiplies + 2 FMAs (abs error < 2 ulp of
// 1).  Two 16-byte shared-memory loads replace 19 fp64 and ~20 integer instructions of the
// polynomial version below (sincos_turn), which stays as the reference implementation in the tests.
template <class Tab> MCB_FN void sincos_turn20(uint32_t k, double &cs, double &sn, const Tab &T)
{
    const uint32_t i = (k >> 10) & 1023u, j = k & 1023u;
    double ch, sh, cl, sl;
    T.turn_hi_entry(i, ch, sh);
    T.turn_lo_entry(j, cl, sl);
    cs = fma_(-sh, sl, ch * cl);
    sn = fma_(ch, sl, sh * cl);
}

// ---- -2 ln(u) for u in (0, 1] --------------------------------------------------------------------
// u = 2^e m, m in [1,2); i = top 8 mantissa bits; r = m c_i - 1 in [0, 2^-8);
// ln u = e ln2 + (-ln c_i) + log1p(r), log1p by its degree-6 Taylor polynomial (|error| < 2^-59).
// The result can come out as -1e-17 instead of +0 when u is one ulp below 1; callers take |.|.
template <class Tab> MCB_FN double neg2log_unit(double u, const Tab &T)
{
    const int hi = hi_word(u);
    const int idx = (hi >> 12) & 0xff;
    const double m = make_double((hi & 0x000fffff) | 0x3ff00000, lo_word(u));
    // exponent as a double without a conversion instruction: 2^52 + biased exponent, minus (2^52 + 1023)
    const double e = make_double(0x43300000, (int)((unsigned)hi >> 20)) - MCB_K(Tab, log_magic, 4503599627371519.0);
    double c, l;
    T.log_entry(idx, c, l);
    const double r = fma_(m, c, -1.0);
    double q = fma_(r, MCB_K(Tab, log_c6, -1.0 / 6.0), MCB_K(Tab, log_c5, 0.2));
    q = fma_(r, q, -0.25);
    q = fma_(r, q, MCB_K(Tab, log_c3, 1.0 / 3.0));
    q = fma_(r, q, -0.5);
    const double p = fma_(r * r, q, r);                    // log1p(r)
    // -2 (e ln2 + l) + 1e-300: the tiny offset (free: it rides in an FMA) keeps the result away from
    // an exact 0 at u == 1, so the square root below needs no zero guard
    const double t = fma_(e, MCB_K(Tab, neg2ln2, -2.0 * 0x1.62e42fefa39efp-1), fma_(l, -2.0, MCB_K(Tab, tiny, 1e-300)));
    return fma_(p, -2.0, t);
}

// k * ln(u) for a caller-chosen k (k_ln2 = k ln 2): the scale rides on the three constants of the final FMAs, so
// e.g. b^2 (-2 ln u) -- the squared radius of a Box-Muller pair already multiplied by a diffusion scale b --
// costs the same 12 instructions as -2 ln u.  neg2log_unit(u) == scaled_log_unit(u, -2, -2 ln 2).
template <class Tab> MCB_FN double scaled_log_unit(double u, const Tab &T, double k, double k_ln2)
{
    const int hi = hi_word(u);
    const int idx = (hi >> 12) & 0xff;
    const double m = make_double((hi & 0x000fffff) | 0x3ff00000, lo_word(u));
    const double e = make_double(0x43300000, (int)((unsigned)hi >> 20)) - MCB_K(Tab, log_magic, 4503599627371519.0);
    double c, l;
    T.log_entry(idx, c, l);
    const double r = fma_(m, c, -1.0);
    double q = fma_(r, MCB_K(Tab, log_c6, -1.0 / 6.0), MCB_K(Tab, log_c5, 0.2));
    q = fma_(r, q, -0.25);
    q = fma_(r, q, MCB_K(Tab, log_c3, 1.0 / 3.0));
    q = fma_(r, q, -0.5);
    const double p = fma_(r * r, q, r);                    // log1p(r)
    const double t = fma_(e, k_ln2, fma_(l, k, MCB_K(Tab, tiny, 1e-300)));
    return fma_(p, k, t);
}

// ---- sqrt(x), x > 0 finite and normal ------------------------------------------------------------
// y ~ 1/sqrt(x) to 2^-22 (MUFU.RSQ64H); g = x y, h = y/2.
//   kShort = false: two coupled Newton steps r = 1/2 - h g; g += g r; h += h r (7 fp64 instructions; g and h
//                   update side by side).
//   kShort = true:  one coupled step brings g to 1.5 * 2^-44; the Newton correction g += (x - g^2) h only needs h
//                   to the seed's accuracy (error 2^-44 * 2^-22), so h is never refined and y/2 is an exponent
//                   decrement on the integer pipe: 5 fp64 instructions.
// Both are within 1 ulp (tests/test_device_math64.py).  Which one is faster depends on the kernel around it
// (measured, profiles/r01j_tune_vanilla.txt): the European call runs 3 % FASTER with the longer version (its two
// independent update chains schedule better between the table look-ups), the basket and CVA kernels 1.5-2 % faster
// with the short one.  No zero guard: the only callers feed |k ln u| >= 1e-300.
template <bool kShort = false> MCB_FN double sqrt_pos(double x)
{
    const double y = rsqrt_seed(x);
    double g = x * y;
    if (kShort) {
        const double h = make_double(hi_word(y) - 0x00100000, lo_word(y));   // y / 2
        const double r = fma_(-h, g, 0.5);
        g = fma_(g, r, g);
        const double d = fma_(-g, g, x);
        return fma_(d, h, g);
    }
    double h = 0.5 * y;
    double r = fma_(-h, g, 0.5);
    g = fma_(g, r, g);
    h = fma_(h, r, h);
    r = fma_(-h, g, 0.5);
    return fma_(g, r, g);
}

// ---- 1/x, x finite and not tiny ------------------------------------------------------------------
// One cubically convergent step from the 2^-22 seed: with e = 1 - x y, 1/x = y (1 + e + e^2 + ...), and
// y (1 + e + e^2) is off by e^3 < 2^-60 -- three fp64 instructions where two Newton steps take four.
MCB_FN double rcp_newton(double x)
{
    const double y = rcp_seed(x);
    const double e = fma_(-x, y, 1.0);
    return fma_(y, fma_(e, e, e), y);
}

// ---- cos and sin of 2 pi k / 2^52 for a 52-bit integer k = (k_hi[19:0] : k_lo) --------------------
// octant q = top 3 bits, frac = the other 49 bits (exact); odd octants are reflected (1 - frac,
// exact), phi = frac pi/4 in [0, pi/4]; fdlibm __kernel_sin / __kernel_cos polynomials; the
// octant's swap and signs are integer selects / sign-bit XORs.
MCB_FN void sincos_turn(uint32_t k_hi, uint32_t k_lo, double &cs, double &sn)
{
    const uint32_t q = (k_hi >> 17) & 7u;
    // d in [1,2): 49 fraction bits moved up by 3
    const uint32_t fh = ((k_hi << 3) | (k_lo >> 29)) & 0x000fffffu;
    const double d = make_double((int)(fh | 0x3ff00000u), (int)(k_lo << 3));
    const bool odd = (q & 1u) != 0u;
    // y = odd ? 2 - d : d - 1   (exact): flip d's sign bit for odd octants, add 2 or -1
    const double ds = make_double(hi_word(d) ^ (int)(odd ? 0x80000000u : 0u), lo_word(d));
    const double y = ds + make_double(odd ? 0x40000000 : (int)0xbff00000u, 0);
    const double x = y * 0x1.921fb54442d18p-1;  // pi/4
    const double z = x * x;
    double s = fma_(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    s = fma_(z, s, 2.75573137070700676789e-06);
    s = fma_(z, s, -1.98412698298579493134e-04);
    s = fma_(z, s, 8.33333333332248946124e-03);
    s = fma_(z, s, -1.66666666666666324348e-01);
    const double sin_phi = fma_(x * z, s, x);
    double c = fma_(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    c = fma_(z, c, -2.75573143513906633035e-07);
    c = fma_(z, c, 2.48015872894767294178e-05);
    c = fma_(z, c, -1.38888888888741095749e-03);
    c = fma_(z, c, 4.16666666666666019037e-02);
    const double cos_phi = fma_(z * z, c, fma_(z, -0.5, 1.0));
    // octant table: swap for q in {1,2,5,6}; cos negative for q in {2,3,4,5}; sin negative for q >= 4
    const bool swap = ((q + 1u) & 2u) != 0u;
    const double a = swap ? sin_phi : cos_phi;
    const double b = swap ? cos_phi : sin_phi;
    const int cos_sign = (int)(((q + 2u) & 4u) << 29);
    const int sin_sign = (int)((q & 4u) << 29);
    cs = make_double(hi_word(a) ^ cos_sign, lo_word(a));
    sn = make_double(hi_word(b) ^ sin_sign, lo_word(b));
}

// ---- e^x -----------------------------------------------------------------------------------------
// n = rint(x 256/ln2) (magic-number add), r = x - n ln2/256 (Cody-Waite, |r| <= ln2/512),
// e^x = 2^(n>>8) * T[n & 255] * (1 + r + r^2/2 + r^3/6 + r^4/24).  The power of two is added to the
// exponent field as an integer, so the argument must satisfy |x| <= 700: the host validates every
// job's reachable exponent range (engine.cu: make_*_job) and the CVA kernel floors -d^2/2 at -700.
template <class Tab> MCB_FN double exp_tab(double x, const Tab &T)
{
    const double magic = 6755399441055744.0;  // 1.5 * 2^52 (an immediate: low word zero)
    const double t = fma_(x, MCB_K(Tab, exp_scale, 0x1.71547652b82fep+8), magic);
    const int n = lo_word(t);
    const double nd = t - magic;
    double r = fma_(nd, MCB_K(Tab, exp_ln2_hi, -0x1.62e42fee00000p-9), x);
    r = fma_(nd, MCB_K(Tab, exp_ln2_lo, -0x1.a39ef35793c76p-41), r);
    const double tj = T.exp_entry(n & 255);
    double p = fma_(r, MCB_K(Tab, exp_c4, 1.0 / 24.0), MCB_K(Tab, exp_c3, 1.0 / 6.0));
    p = fma_(r, p, 0.5);
    p = fma_(r * r, p, r);          // e^r - 1
    const double v = fma_(tj, p, tj);
    // (the exponent insertion as mask + IMAD -- one instruction fewer than shift, mask, add -- was measured: European
    // call 2 % slower, basket-10 1.3 % faster, CVA 1 % slower, profiles/r01p_ab_experiments.txt; not used)
    return make_double(hi_word(v) + ((n >> 8) << 20), lo_word(v));
}

// ---- max(x, 0) on the integer pipe: clear every bit when the sign bit is set ----------------------
MCB_FN double relu64(double x)
{
    const int hi = hi_word(x);
    const int keep = ~(hi >> 31);
    return make_double(hi & keep, lo_word(x) & keep);
}

#ifdef MCB_HOST_MATH
}  // namespace hostmath
#endif
}  // namespace mcb
