// device_math64.cuh -- the fp64 special functions of the pricing kernels, hand-built for the
// sm_100a pipe mix: the fp64 pipe retires 64 thread-instr/clk/SM and there is no fp64 MUFU, so
// libdevice's log / sincospi / exp / division cost 30 / 27 / 18 / 8+ fp64 instructions plus as many
// integer ones (measured: 162 warp-instructions per vanilla path, profiles/r01a_*).  Here:
//
//   scaled_log_unit  k ln(u), u in (0,1]  table of 512 reciprocals (shared memory) + degree-5 log1p; the exponent's
//                                          share comes from a 64-entry table per job (LogScale64)          9 fp64
//   sqrt_pos         sqrt(x)               MUFU.RSQ64H seed + 2 coupled Newton steps (or 1 + a correction) 7 / 5 fp64
//   sincos_turn20    cos/sin(2 pi k/2^20)  two-level table (12 + 8 bits) and the addition theorems         4 fp64
//   sincos_turn      cos/sin(2 pi k/2^52)  octant taken from the integer bits (exact reduction),
//                                          fdlibm kernel polynomials on [0, pi/4] (tests only)             18 fp64
//   exp_units        2^(y/256)             n = rint(y) by magic add, r = y - n exactly, table of 2^(j/256) with
//                                          the exponent added as an integer, degree-4 expm1               8 fp64
//   exp_tab          e^x                   the same behind a Cody-Waite reduction                         10 fp64
//   rcp_newton       1/x                   MUFU.RCP64H seed + 1 cubic step                                 3 fp64
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
MCB_FN uint32_t and_or(uint32_t a, uint32_t mask, uint32_t c) { return (a & mask) | c; }
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
// (a & mask) | c as ONE LOP3: a table offset "index field of a word, plus this thread's replica" costs a shift and
// this, where base + replica + (index << s) written as pointer arithmetic compiled to shift, mask, add
MCB_FN uint32_t and_or(uint32_t a, uint32_t mask, uint32_t c)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(mask), "r"(c));
    return d;
}
#endif

// ---- the tables as the kernels hold them in shared memory ------------------------------------------
// (random per-thread indices: a constant-bank read would serialise).  Two of the generated tables are kept in a
// form that saves integer work per look-up:
//   log:  c_i carries the exponent bias of 1.0 on top of its own exponent field (high word + 0x3ff00000), so that
//         subtracting u's exponent field from it is c_i 2^-e -- the reciprocal for u itself, not for its mantissa
//         taken out and re-biased (one LOP3 + one integer add where building m = u 2^-e took three instructions);
//   exp:  2^(j/256) has j << 12 subtracted from its high word, so that adding n << 12 for n = 256 k + j gives
//         2^k 2^(j/256) = 2^(n/256) in one integer instruction (shift, mask and add on the result before).
// Neither is a valid double on its own; both changes are exact, the values the functions return did not change.
MCB_FN double bias_log_recip(double c) { return make_double(hi_word(c) + 0x3ff00000, lo_word(c)); }
MCB_FN double bias_exp_entry(double t, int j) { return make_double(hi_word(t) - (j << 12), lo_word(t)); }

// index bits of the logarithm's table: 512 entries and a degree-5 log1p (256 and degree 6 until round 2: one fp64
// instruction more per Box-Muller pair, profiles/r02k_ab_experiments.txt 3)
constexpr int kLogBits = 9;
constexpr int kLogEntries = 1 << kLogBits;

// Plain layout: host build, instrumentation and per-path kernels (their coarse angle table stays where the
// generated one lies -- global memory on the device: 64 KB do not fit a static shared-memory allocation).
struct Tables64 {
    double log_tab[kLogEntries][2];       // { c_i (biased), -ln c_i }
    double exp_tab[256];          // 2^(j/256) (biased)
    double turn_lo[256][2];       // { cos, sin } of 2 pi j / 2^20
    // hi_u = high word of u: the entry of its top kLogBits mantissa bits
    MCB_MEMBER void log_entry(int hi_u, double &c, double &l) const
    {
        const int i = (hi_u >> (20 - kLogBits)) & (kLogEntries - 1);
        c = log_tab[i][0];
        l = log_tab[i][1];
    }
    MCB_MEMBER double exp_entry(int n) const { return exp_tab[n & 255]; }
    // k = a 20-bit turn fraction (higher bits ignored): coarse entry of its top 12 bits, fine entry of its low 8
    MCB_MEMBER void turn_hi_entry(uint32_t k, double &c, double &s) const { c = kTurnHiTable[(k >> 8) & 4095u][0]; s = kTurnHiTable[(k >> 8) & 4095u][1]; }
    MCB_MEMBER void turn_lo_entry(uint32_t k, double &c, double &s) const { c = turn_lo[k & 255u][0]; s = turn_lo[k & 255u][1]; }
    MCB_MEMBER void fill(int i)   // for every i < 256
    {
        for (int j = i; j < kLogEntries; j += 256) {
            log_tab[j][0] = bias_log_recip(kLogTable[j][0]);
            log_tab[j][1] = kLogTable[j][1];
        }
        exp_tab[i] = bias_exp_entry(kExpTable[i], i);
        turn_lo[i][0] = kTurnLoTable[i][0];
        turn_lo[i][1] = kTurnLoTable[i][1];
    }
};

// The plain layout with the coarse angle table in shared memory as well (the two-pass wide-basket kernel).
struct Tables64Wide : Tables64 {
    double turn_hi[4096][2];
    MCB_MEMBER void turn_hi_entry(uint32_t k, double &c, double &s) const { c = turn_hi[(k >> 8) & 4095u][0]; s = turn_hi[(k >> 8) & 4095u][1]; }
    MCB_MEMBER void fill_wide(int i)   // for every i < 4096
    {
        if (i < 256)
            fill(i);
        turn_hi[i][0] = kTurnHiTable[i][0];
        turn_hi[i][1] = kTurnHiTable[i][1];
    }
};

#ifndef MCB_HOST_MATH
// Bank-conflict-free layout for the pricing kernels.  The indices are random per thread, so in the plain layout
// the 8 threads of a quarter-warp (16-byte loads) or the 16 of a half-warp (8-byte loads) collide in the 32
// shared-memory banks: ~2.6 wavefronts where 1 would do.  ncu on the plain layout (profiles/r01j_vanilla_f64_2p32.txt):
// 62 % of all shared-memory wavefronts were bank conflicts and the shared-memory pipe was 97 % busy -- the
// European-call fp64 kernel was bound by it, not by the fp64 pipe.  Here the small tables are replicated once
// per bank group: thread t reads replica t % 8 (16-byte entries: 128 bytes per index, one bank group per replica)
// or t % 16 (8-byte entries), so threads that are served together can never share a bank.  Same values, same
// arithmetic, bit-identical results.
// The angle is split 12 + 8 bits so that its FINE table (256 entries) is small enough to be replicated as well; the
// coarse one (4096 entries, 64 KB) stays a single copy at 2.6 wavefronts per quarter-warp.  With the 10 + 10 split of
// round 1 (two single 16 KB tables) the angle look-ups were 22 of the 29 wavefronts of a Box-Muller pair and the
// shared-memory pipe the busiest unit of the fp64 call (68 %) and the 10-asset basket (79 %); four copies of each
// (192 KB) had been measured slower (profiles/r01p_ab_experiments.txt) -- with pointer arithmetic, not the one-
// instruction offsets used here.  192 KB in all, one table set per SM shared by the CTA's sub-blocks.
struct Tables64Rep {
    double log_rep[kLogEntries][8][2];   // [index][replica]{ c_i (biased), -ln c_i }                      64 KB
    // the two 256-entry tables share rows of 256 bytes: a row's byte offset is then the index byte moved up by one
    // byte, and "index byte of a word, plus this thread's replica" is ONE byte permute (PRMT)
    struct Row {
        double exp[16];          // [replica] 2^(j/256) (biased)
        double turn_lo[8][2];    // [replica]{ cos, sin } of 2 pi j / 2^20
    } rows[256];                 //                                                                        64 KB
    double turn_hi[4096][2];     // { cos, sin } of 2 pi i / 4096                                          64 KB
    MCB_MEMBER void log_entry(int hi_u, double &c, double &l) const
    {
        const uint32_t off = and_or((uint32_t)hi_u >> (13 - kLogBits), (uint32_t)(kLogEntries - 1) << 7, (threadIdx.x & 7u) << 4);
        const double2 v = *reinterpret_cast<const double2 *>(reinterpret_cast<const char *>(log_rep) + off);
        c = v.x;
        l = v.y;
    }
    // { 0, 0, low byte of index, offset inside the row }
    static MCB_MEMBER uint32_t row_offset(uint32_t index, uint32_t inside) { return __byte_perm(index, inside, 0x6504); }
    MCB_MEMBER double exp_entry(int n) const
    {
        return *reinterpret_cast<const double *>(reinterpret_cast<const char *>(rows) + row_offset((uint32_t)n, (threadIdx.x & 15u) << 3));
    }
    MCB_MEMBER void turn_hi_entry(uint32_t k, double &c, double &s) const
    {
        const double2 v = *reinterpret_cast<const double2 *>(reinterpret_cast<const char *>(turn_hi) + ((k >> 4) & 0xfff0u));
        c = v.x;
        s = v.y;
    }
    MCB_MEMBER void turn_lo_entry(uint32_t k, double &c, double &s) const
    {
        const double2 v = *reinterpret_cast<const double2 *>(reinterpret_cast<const char *>(rows) +
                                                             row_offset(k, 128u + ((threadIdx.x & 7u) << 4)));
        c = v.x;
        s = v.y;
    }
};
#endif

// The exponent part of a scaled logarithm, per job: entry (E & 63) = (E - 1023) k ln 2 + 1e-300 for the biased
// exponent E of u in [2^-63, 1] (the kernels' radius uniform is >= 2^-44).  One 8-byte load replaces the conversion
// of the exponent (an fp64 subtract of a magic number) and its FMA.  The tiny offset keeps k ln u away from an exact
// 0 at u == 1, so the square root after it needs no zero guard.  64 entries: a sub-block's threads 0..63 fill it.
struct LogScale64 {
    double e[64];
    MCB_MEMBER void fill(int i, double k_ln2) { e[i] = fma_((double)(i - 63), k_ln2, 1e-300); }
    // ebits = u's exponent field in place (high word & 0x7ff00000)
    MCB_MEMBER double entry(int ebits) const
    {
        return *reinterpret_cast<const double *>(reinterpret_cast<const char *>(e) + (((uint32_t)ebits >> 17) - (960u << 3)));
    }
};

// ---- cos and sin of 2 pi k / 2^20 for a 20-bit integer k (the Box-Muller angle of the kernels) ---
// Two-level table: k = 256 i + j, angle = coarse_i + fine_j, and the addition theorems give the
// result from four correctly rounded table values with 2 multiplies + 2 FMAs (abs error < 2 ulp of
// 1).  Two 16-byte shared-memory loads replace 19 fp64 and ~20 integer instructions of the
// polynomial version below (sincos_turn), which stays as the reference implementation in the tests.
// Bits of k above the 20th are ignored.
template <class Tab> MCB_FN void sincos_turn20(uint32_t k, double &cs, double &sn, const Tab &T)
{
    double ch, sh, cl, sl;
    T.turn_hi_entry(k, ch, sh);
    T.turn_lo_entry(k, cl, sl);
    cs = fma_(ch, cl, -(sh * sl));
    sn = fma_(sh, cl, ch * sl);
}

// ---- k ln(u) for u in [2^-63, 1] and a caller-chosen k ---------------------------------------------
// u = 2^e m, m in [1,2); i = top 9 mantissa bits; r = m c_i - 1 = u (c_i 2^-e) - 1 in [0, 2^-9);
// ln u = e ln2 + (-ln c_i) + log1p(r), log1p by its degree-5 Taylor polynomial (|error| < 2^-56).
// The scale k rides on the constants of the last two FMAs and on the job's exponent table S (LogScale64, filled
// with k ln 2), so e.g. b^2 (-2 ln u) -- the squared radius of a Box-Muller pair already multiplied by a diffusion
// scale b -- costs the same 9 fp64 instructions as -2 ln u (k = -2).
// With a 52-bit u one ulp below 1 the result can come out as -1e-17 instead of +0; the kernels' 44-bit uniforms
// cannot get there (device_math.cuh).
template <class Tab> MCB_FN double scaled_log_unit(double u, const Tab &T, double k, const LogScale64 &S)
{
    const int hi = hi_word(u);
    const int ebits = hi & 0x7ff00000;
    double c, l;
    T.log_entry(hi, c, l);
    const double r = fma_(u, make_double(hi_word(c) - ebits, lo_word(c)), -1.0);
    double q = fma_(r, 0.2, -0.25);                        // r < 2^-9: r^6/6 < 2^-56
    q = fma_(r, q, 1.0 / 3.0);
    q = fma_(r, q, -0.5);
    const double p = fma_(r * r, q, r);                    // log1p(r)
    return fma_(p, k, fma_(l, k, S.entry(ebits)));
}
// -2 ln(u): S filled with -2 ln 2
template <class Tab> MCB_FN double neg2log_unit(double u, const Tab &T, const LogScale64 &S) { return scaled_log_unit(u, T, -2.0, S); }

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
// exp_units(y) = 2^(y/256) = e^(y ln2/256): the argument in units of the table step.  n = rint(y) by a magic-number
// add, r = y - n EXACTLY (|r| <= 1/2), 2^(n/256) from the table with the power of two added to its exponent field
// (bias_exp_entry), e^(r ln2/256) - 1 by its degree-4 Taylor polynomial with the powers of ln2/256 folded into the
// coefficients (truncation (ln2/512)^5/120 < 2^-54).  8 fp64 instructions and no Cody-Waite reduction: a caller
// whose exponent is a sum of products (a + b z; a GBM step) scales its constants by 256/ln2 on the host and pays
// nothing for it.  The relative rounding error of y is the same 2^-53 that the natural-units argument would carry,
// so nothing is lost.  |y| <= 700 * 256/ln2: the host validates every job's reachable exponent range
// (engine.cu: make_*_job) and the CVA kernel floors its density exponent.
// kLateTable: which of the two equivalent last steps -- fma(ts r, p, ts) has the table entry ts in a multiply beside the
// polynomial, fma(ts, r p, ts) needs it only in the final FMA.  Measured per kernel (profiles/r02k_ab_experiments.txt,
// 10): the European call is 0.3 % faster with the first, the basket 1.3 % and the CVA 2.5 % faster with the second.
template <bool kLateTable = false, class Tab> MCB_FN double exp_units(double y, const Tab &T)
{
    const double magic = 6755399441055744.0;  // 1.5 * 2^52 (an immediate: low word zero)
    const double t = y + magic;
    const int n = lo_word(t);
    const double r = y - (t - magic);
    const double tj = T.exp_entry(n);
    const double ts = make_double(hi_word(tj) + (int)((uint32_t)n << 12), lo_word(tj));   // 2^(n/256)
    double p = fma_(r, 0x1.3b2ab6fba4e77p-39 /* h^4/24 */, 0x1.c6b08d704a0c0p-29 /* h^3/6 */);
    p = fma_(r, p, 0x1.ebfbdff82c58fp-19 /* h^2/2 */);
    p = fma_(r, p, 0x1.62e42fefa39efp-9 /* h = ln2/256 */);
    return kLateTable ? fma_(ts, r * p, ts) : fma_(ts * r, p, ts);
}

// e^x for an argument in natural units: the same with a Cody-Waite reduction in front (10 fp64 instructions).
// |x| <= 700.
template <class Tab> MCB_FN double exp_tab(double x, const Tab &T)
{
    const double magic = 6755399441055744.0;
    const double t = fma_(x, 0x1.71547652b82fep+8, magic);
    const int n = lo_word(t);
    const double nd = t - magic;
    double r = fma_(nd, -0x1.62e42fee00000p-9, x);
    r = fma_(nd, -0x1.a39ef35793c76p-41, r);
    const double tj = T.exp_entry(n);
    const double ts = make_double(hi_word(tj) + (int)((uint32_t)n << 12), lo_word(tj));
    double p = fma_(r, 1.0 / 24.0, 1.0 / 6.0);
    p = fma_(r, p, 0.5);
    p = fma_(r * r, p, r);          // e^r - 1
    return fma_(ts, p, ts);
}

#ifdef MCB_HOST_MATH
}  // namespace hostmath
#endif
}  // namespace mcb
