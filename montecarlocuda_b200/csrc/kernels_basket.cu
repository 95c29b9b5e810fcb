// kernels_basket.cu -- basket call on N Cholesky-correlated underlyings, fp32 and fp64 (sm_100a).
//
// Replaces brownianVect + basketPayoff + basketOptMonteCarlo (DP/MonteCarloKernel.cu:74-101,
// :133-177).  One draw unit = one path; draw block j of the path's sub-stream gives normals
// kNpb * j .. kNpb * j + kNpb - 1 (kNpb = 6 in fp32, 4 in fp64).
//   x_i    = a_i + sum_j F_ij z_j      F_ij = v_i sqrt(T) L_ij,  a_i = (r - v_i^2/2) T + v_i sqrt(T) d_i
//   payoff = max(sum_i m_i e^{x_i} - K, 0),  m_i = w_i s_i      (x in the units of exp_scaled: log2 units for MUFU.EX2
//                                                                in fp32, units of ln2/256 for exp_units in fp64)
// The mat-vec is a column sweep kept in registers: normal j is produced, applied to the N - j
// accumulators at or below the diagonal and discarded, so N accumulators + one normal are live
// and every factor entry reaches the FMA pipe as a constant-bank operand.  The reference indexes
// g[], bt[], s[] with runtime bounds, which puts them in local memory (LDL/STL),
// and multiplies the zero upper triangle too.  kFull keeps that behaviour for a caller whose p
// is not triangular.
#include <atomic>
#include <cstdlib>
#include <type_traits>
#include <utility>
#include <vector>

#include "device_math.cuh"
#include "launch.h"
#include "table_lock.h"

// The table is addressed from inline PTX by name, hence C linkage and global scope.
extern "C" {
__constant__ __align__(16) unsigned char mcb_basket_table[36 * 1024];
}

namespace mcb {

constexpr int kBasketTableBytes = 36 * 1024;
static TableLock g_basket_lock;

// One table entry as a constant-space load at a compile-time byte offset.  The load is volatile
// inline PTX on purpose: written as ordinary C++ (table[i]), ptxas treats the ~2000 entries of a
// 64-asset factor as loop invariants of the path loop, hoists them and spills them to local memory
// (9-23 KB of stack per thread, measured); pinned like this each entry stays where it is used and
// ptxas feeds it to the FMA through a uniform register (LDCU.128 serves four FMAs).
template <typename Real, int kByteOffset> __device__ __forceinline__ Real table_entry();
template <int kByteOffset> __device__ __forceinline__ float table_entry_f32()
{
    float t;
    asm volatile("ld.const.f32 %0, [mcb_basket_table+%1];" : "=f"(t) : "n"(kByteOffset));
    return t;
}
template <int kByteOffset> __device__ __forceinline__ double table_entry_f64()
{
    double t;
    asm volatile("ld.const.f64 %0, [mcb_basket_table+%1];" : "=d"(t) : "n"(kByteOffset));
    return t;
}
template <typename Real, int kByteOffset> __device__ __forceinline__ Real table_entry()
{
    if constexpr (sizeof(Real) == 4)
        return table_entry_f32<kByteOffset>();
    else
        return table_entry_f64<kByteOffset>();
}

// x2 = {lo, hi} += table pair at kByteOffset * {z, z}: one FFMA2 (fma.rn.f32x2, new in sm_100).
// The packed FMA keeps the FP32 pipe at 128 FMA/clk/SM with HALF the issue slots, and the freed
// slots are where the Philox LOP3s, the uniform loads and the MUFU ops of the next normals issue
// (profiles/r01_pipe_ffma2.txt: FFMA2 x8 + LOP3 x4 costs the same 16.2 cycles as FFMA2 x8 alone).
template <int kByteOffset>
__device__ __forceinline__ void packed_fma_entry(unsigned long long &x2, unsigned long long zz)
{
    asm volatile("{\n\t.reg .b64 f;\n\tld.const.b64 f, [mcb_basket_table+%2];\n\tfma.rn.f32x2 %0, f, %1, %0;\n\t}"
                 : "+l"(x2)
                 : "l"(zz), "n"(kByteOffset));
}
template <int kByteOffset>
__device__ __forceinline__ unsigned long long table_pair()
{
    unsigned long long t;
    asm volatile("ld.const.b64 %0, [mcb_basket_table+%1];" : "=l"(t) : "n"(kByteOffset));
    return t;
}
__device__ __forceinline__ unsigned long long pack2(float lo, float hi)
{
    return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32);
}

// Table layout.  Column-major; column `col` holds rows first_row(col) .. N-1.  fp32 with an even
// width runs on packed FMAs (FFMA2, two rows per instruction), so its columns start on an even row
// (the extra entry above the diagonal of an odd column is an exact 0) and every pair is 8-byte
// aligned; fp64 and odd widths keep the plain packed triangle.
template <typename Real, int N, bool kFull>
struct BasketTable {
    static constexpr bool kPaired = sizeof(Real) == 4 && N % 2 == 0;
    // wide fp32 factors are read from a shared-memory copy with 16-byte loads (see Basket::column):
    // their columns start on 16-byte boundaries
    static constexpr bool kSharedFactor = kPaired && N >= 32;
    static __host__ __device__ constexpr int first_row(int col) { return kFull ? 0 : (kPaired ? (col & ~1) : col); }
    static __host__ __device__ constexpr int column_start(int col)
    {
        int off = 0;
        for (int c = 0; c < col; c++) {
            off += N - first_row(c);
            if (kSharedFactor)
                off = (off + 3) & ~3;
        }
        return off;
    }
    static constexpr int kFactor = column_start(N);
    Real factor[kFactor];
    Real a[N];
    Real m[N];
    Real k;
    static __host__ __device__ constexpr int index(int col, int row) { return column_start(col) + (row - first_row(col)); }
};

constexpr int basket_min_blocks(int n, int real_bytes)
{
    // registers: the accumulators + ~44 (fp32) / ~60 (fp64) for generator, normals and loop state
    const int regs = n * real_bytes / 4 + (real_bytes == 8 ? 60 : 44);
    int blocks = 65536 / (kThreads * regs);
    // fp64 from 8 assets up: at 3 CTAs per SM (80 registers) the sweep spills 56-192 bytes per thread and the
    // local-memory round trips cost more than the third CTA hides (N=10, 2^28 paths: 9.58 ms -> 8.69 ms at 2 CTAs)
    if (real_bytes == 8 && n >= 8 && blocks > 2)
        blocks = 2;
    return blocks < 1 ? 1 : (blocks > 4 ? 4 : blocks);
}

// Shared-memory copy of a wide fp32 factor.  Read through the constant bank, a 64-asset factor
// (8.4 KB) misses the small first-level constant cache on every access: the ncu source page of
// profiles/r01d_basket64_f32_ffma2.txt charges a third of all warp-stall samples to FFMA2s waiting on
// their LDCU.128 (short scoreboard).  Shared memory answers a broadcast 16-byte load in ~30 cycles.
template <int kFloats>
struct SharedFactor : NoShared {
    __align__(16) float factor[kFloats];
    __device__ __forceinline__ void load()
    {
        const float *src = reinterpret_cast<const float *>(mcb_basket_table);
        for (int i = threadIdx.x; i < kFloats; i += blockDim.x)
            factor[i] = src[i];
    }
};

// The fp64 counterpart, on top of the (replicated) math tables.  Through the constant bank ptxas hoists the
// loop-invariant factor loads of a 10-asset basket out of the path loop into uniform registers, runs out of
// them and shuffles the overflow between uniform and vector registers once per path (52 R2UR + ~40 MOV per path,
// profiles/r01k_basket10_f64_2p28.txt).  From shared memory two entries arrive per (broadcast, conflict-free)
// 16-byte load right where their DFMAs are.
template <int kDoubles>
struct SharedFactor64 : SharedTables64Rep {
    __align__(16) double factor[kDoubles];
    __device__ __forceinline__ void load()
    {
        SharedTables64Rep::load();
        const double *src = reinterpret_cast<const double *>(mcb_basket_table);
        for (int i = threadIdx.x; i < kDoubles; i += blockDim.x)
            factor[i] = src[i];
    }
};
// Wide fp64 baskets (two-pass sweep, see Basket::eval_two_pass): the plain math tables with the coarse angle table
// next to them, plus one slot per thread and sub-block for each normal of the first half ([normal][thread]: consecutive
// threads, conflict-free 8-byte accesses).  14 KB + 64 KB + 2 x 64 KB; the replicated tables (192 KB) would not fit
// next to the slots.
template <int kHalf>
struct SharedTwoPass64 {
    Tables64Wide t;
    double stash[2][kHalf][kThreads];  // [sub-block]: Basket::kSubBlocks == 2 for the two-pass kernel (static_assert there)
    __device__ __forceinline__ void load()
    {
        for (int i = threadIdx.x; i < 4096; i += blockDim.x)
            t.fill_wide(i);
    }
};
// xa, xb += {2 consecutive factor entries at smem address base + kByteOffset} * z
template <int kByteOffset>
__device__ __forceinline__ void fma_pair_shared_f64(double &xa, double &xb, double z, uint32_t base)
{
    asm volatile(
        "{\n\t.reg .f64 fa, fb;\n\tld.shared.v2.f64 {fa, fb}, [%3+%4];\n\t"
        "fma.rn.f64 %0, fa, %2, %0;\n\tfma.rn.f64 %1, fb, %2, %1;\n\t}"
        : "+d"(xa), "+d"(xb)
        : "d"(z), "r"(base), "n"(kByteOffset));
}
template <int kByteOffset>
__device__ __forceinline__ void fma_one_shared_f64(double &xa, double z, uint32_t base)
{
    asm volatile("{\n\t.reg .f64 fa;\n\tld.shared.f64 fa, [%2+%3];\n\tfma.rn.f64 %0, fa, %1, %0;\n\t}"
                 : "+d"(xa)
                 : "d"(z), "r"(base), "n"(kByteOffset));
}

// x2a, x2b += {4 consecutive factor entries at smem address base + kByteOffset} * {z, z}
template <int kByteOffset>
__device__ __forceinline__ void packed_fma_quad(unsigned long long &x2a, unsigned long long &x2b, unsigned long long zz,
                                                uint32_t base)
{
    asm volatile(
        "{\n\t.reg .b64 fa, fb;\n\tld.shared.v2.b64 {fa, fb}, [%3+%4];\n\t"
        "fma.rn.f32x2 %0, fa, %2, %0;\n\tfma.rn.f32x2 %1, fb, %2, %1;\n\t}"
        : "+l"(x2a), "+l"(x2b)
        : "l"(zz), "r"(base), "n"(kByteOffset));
}
template <int kByteOffset>
__device__ __forceinline__ void packed_fma_pair(unsigned long long &x2, unsigned long long zz, uint32_t base)
{
    asm volatile("{\n\t.reg .b64 fa;\n\tld.shared.b64 fa, [%2+%3];\n\tfma.rn.f32x2 %0, fa, %1, %0;\n\t}"
                 : "+l"(x2)
                 : "l"(zz), "r"(base), "n"(kByteOffset));
}

template <typename RealT, int N, bool kFull, bool kAccumLayout = false>
struct Basket {
    using Real = RealT;
    using Table = BasketTable<Real, N, kFull>;
    static_assert(sizeof(Table) <= kBasketTableBytes, "basket table exceeds its constant buffer");
    static constexpr int kUnitPaths = 1;
    static constexpr int kUnroll = 1;
    // fp64: the CTAs of 256 threads an SM holds become ONE CTA of that many sub-blocks around one (replicated) table set
    // fp64, 64 assets: 64 accumulators + generator need 255 registers, i.e. 8 warps per SM, and the kernel sat at 27 % of
    // the fp64 pipe (758 ms for 2^30 paths, profiles/r01p_bench_all_precisions.json).  Two passes of 32 accumulators
    // (assets 0..31, then 32..63 with the first 32 normals parked in shared memory) fit 128 registers: twice the warps,
    // the same FMAs in the same order per accumulator and the same summation order -- bit-identical values.
    static constexpr bool kTwoPass = kAccumLayout && sizeof(RealT) == 8 && N == 64 && !kFull;
    static constexpr int kHalf = N / 2;
    static constexpr int kSubBlocks = kTwoPass ? 2 : (kAccumLayout && sizeof(RealT) == 8) ? basket_min_blocks(N, 8) : 1;
    static constexpr int kMinBlocks = kSubBlocks > 1 ? 1 : basket_min_blocks(N, (int)sizeof(Real));
    static constexpr int kNpb = NormalsPerBlock<RealT>::value;
    struct Params {
        PhiloxKeys keys;
    };
    static constexpr bool kSharedFactor = Table::kSharedFactor;
    // fp64 pricing kernels from 8 assets up read the factor from shared memory too (see SharedFactor64)
#ifndef MCB_BASKET_SHARED_FACTOR64
#define MCB_BASKET_SHARED_FACTOR64 1
#endif
    static constexpr bool kSharedFactor64 = MCB_BASKET_SHARED_FACTOR64 && kAccumLayout && sizeof(RealT) == 8 && N >= 8 && N <= 16;  // wider: ptxas spills more than it saves
    using Shared = std::conditional_t<
        kTwoPass, SharedTwoPass64<kHalf>,
        std::conditional_t<
        kSharedFactor, SharedFactor<(kSharedFactor ? Table::kFactor : 1)>,
        std::conditional_t<kSharedFactor64, SharedFactor64<(kSharedFactor64 ? Table::kFactor : 1)>,
                           std::conditional_t<kAccumLayout, typename SharedAccumFor<Real>::type, typename SharedFor<Real>::type>>>>;
    using JobState = typename JobStateFor<Real>::type;   // fp64: the exponent table of -2 ln u
    static constexpr bool kClampAtZero = sizeof(Real) == 8;   // fp64: add_value clamps the payoff (device_common.cuh)
    static __device__ __forceinline__ void prepare(const Params &, JobState &job, int tid) { prepare_polar(polar_scale<Real>(1.0), job, tid); }
    // exponents are kept in the units of exp_scaled (fill_table): log2 units for fp32, units of ln2/256 for fp64
    template <class Sh> static __device__ __forceinline__ Real grow(Real x, const Sh &sh) { return exp_scaled<true>(x, sh); }
    static __device__ __forceinline__ Real clamp(Real payoff)
    {
        if constexpr (kClampAtZero)
            return payoff;
        else
            return positive_part(payoff);
    }
    static constexpr int kBlocks = (N + kNpb - 1) / kNpb;
    static constexpr int kFactorBase = 0;
    static constexpr int kABase = Table::kFactor * (int)sizeof(Real);
    static constexpr int kMBase = kABase + N * (int)sizeof(Real);
    static constexpr int kKBase = kMBase + N * (int)sizeof(Real);

    static constexpr bool kPaired = Table::kPaired;
    // accumulators: N scalars, or N/2 packed pairs for the FFMA2 path
    struct State {
        Real x[kPaired ? 1 : N];
        unsigned long long x2[kPaired ? N / 2 : 1];
    };

    // column J of the sweep: x[row] += F[row][J] * z for row = first_row(J) .. N-1
    template <int J, int kFirst, int... kQuad>
    static __device__ __forceinline__ void column_shared(State &st, unsigned long long zz, uint32_t base,
                                                         std::integer_sequence<int, kQuad...>)
    {
        (packed_fma_quad<Table::index(J, kFirst + 4 * kQuad) * 4>(st.x2[kFirst / 2 + 2 * kQuad], st.x2[kFirst / 2 + 2 * kQuad + 1], zz,
                                                                  base),
         ...);
    }
    // fp64 column from shared memory: rows first .. N-1 of column J; 16-byte loads where the entry index is even
    template <int J, int kRow>
    static __device__ __forceinline__ void column_shared_f64(State &st, double z, uint32_t base)
    {
        if constexpr (kRow < N) {
            constexpr int idx = Table::index(J, kRow);
            if constexpr (idx % 2 == 0 && kRow + 1 < N) {
                fma_pair_shared_f64<idx * 8>(st.x[kRow], st.x[kRow + 1], z, base);
                column_shared_f64<J, kRow + 2>(st, z, base);
            } else {
                fma_one_shared_f64<idx * 8>(st.x[kRow], z, base);
                column_shared_f64<J, kRow + 1>(st, z, base);
            }
        }
    }
    template <int J, int... kRow>
    static __device__ __forceinline__ void column(State &st, Real z, const Shared &sh, std::integer_sequence<int, kRow...>)
    {
        constexpr int first = Table::first_row(J);
        if constexpr (kSharedFactor64) {
            column_shared_f64<J, first>(st, z, (uint32_t)__cvta_generic_to_shared(sh.factor));
        } else if constexpr (kSharedFactor) {
            constexpr int pairs = sizeof...(kRow);
            const unsigned long long zz = pack2(z, z);
            const uint32_t base = (uint32_t)__cvta_generic_to_shared(sh.factor);
            column_shared<J, first>(st, zz, base, std::make_integer_sequence<int, pairs / 2>{});
            if constexpr (pairs % 2 == 1)
                packed_fma_pair<Table::index(J, first + 2 * (pairs - 1)) * 4>(st.x2[first / 2 + pairs - 1], zz, base);
        } else if constexpr (kPaired) {
            // (a deeper software pipeline of the table loads was tried: ptxas re-sinks every LDCU.128
            // next to its two FFMA2s whatever the source order, it keeps two uniform quads in flight)
            const unsigned long long zz = pack2(z, z);
            (packed_fma_entry<kFactorBase + Table::index(J, first + 2 * kRow) * 4>(st.x2[first / 2 + kRow], zz), ...);
        } else {
            ((st.x[first + kRow] =
                  fma(table_entry<Real, kFactorBase + Table::index(J, first + kRow) * (int)sizeof(Real)>(), z,
                      st.x[first + kRow])),
             ...);
        }
    }
    template <int J>
    static __device__ __forceinline__ void column_if(State &st, Real z, const Shared &sh)
    {
        if constexpr (J < N) {
            constexpr int rows = N - Table::first_row(J);
            column<J>(st, z, sh, std::make_integer_sequence<int, (kPaired ? rows / 2 : rows)>{});
        }
    }
    // draw block JB: one Philox block -> kNpb normals -> kNpb columns
    template <int JB>
    static __device__ __forceinline__ void draw_block(const Params &P, uint32_t path_lo, uint32_t path_hi, State &st,
                                                      const Shared &sh, const JobState &job)
    {
        uint32_t w[4];
        philox4x32_10(path_lo, path_hi, (uint32_t)JB, kTagBasket, P.keys, w);
        Real z[kNpb];
        normals_from_words(w, z, sh, job);
        columns<JB>(st, z, sh, std::make_integer_sequence<int, kNpb>{});
    }
    template <int JB, int... kQ>
    static __device__ __forceinline__ void columns(State &st, const Real (&z)[kNpb], const Shared &sh,
                                                   std::integer_sequence<int, kQ...>)
    {
        (column_if<JB * kNpb + kQ>(st, z[kQ], sh), ...);
    }
    template <int... kJB>
    static __device__ __forceinline__ void sweep(const Params &P, uint32_t path_lo, uint32_t path_hi, State &st,
                                                 const Shared &sh, const JobState &job, std::integer_sequence<int, kJB...>)
    {
        (draw_block<kJB>(P, path_lo, path_hi, st, sh, job), ...);
    }
    template <int... kI>
    static __device__ __forceinline__ void init(State &st, std::integer_sequence<int, kI...>)
    {
        if constexpr (kPaired)
            ((st.x2[kI] = table_pair<kABase + kI * 8>()), ...);
        else
            ((st.x[kI] = table_entry<Real, kABase + kI * (int)sizeof(Real)>()), ...);
    }
    template <int I>
    static __device__ __forceinline__ Real exponent(const State &st)
    {
        if constexpr (kPaired)
            return __uint_as_float((uint32_t)(st.x2[I / 2] >> (32 * (I % 2))));
        else
            return st.x[I];
    }
    template <int... kI>
    static __device__ __forceinline__ Real payoff(const State &st, const Shared &sh, std::integer_sequence<int, kI...>)
    {
        Real sum = -table_entry<Real, kKBase>();
        ((sum = fma(table_entry<Real, kMBase + kI * (int)sizeof(Real)>(), grow(exponent<kI>(st), sh), sum)), ...);
        return clamp(sum);
    }
    // ---- two-pass sweep (kTwoPass) ----
    // rows kRow0 + kRow... of column J into the accumulators x[row - kOff]
    template <int J, int kRow0, int kOff, int... kRow>
    static __device__ __forceinline__ void column_rows(Real (&x)[kHalf], Real z, std::integer_sequence<int, kRow...>)
    {
        ((x[kRow0 + kRow - kOff] =
              fma(table_entry<Real, kFactorBase + Table::index(J, kRow0 + kRow) * (int)sizeof(Real)>(), z, x[kRow0 + kRow - kOff])),
         ...);
    }
    // draw block JB of the first half: normals 4 JB .. 4 JB + 3 -> their slots and the rows J .. kHalf-1
    template <int JB, int... kQ>
    static __device__ __forceinline__ void first_half_block(const Params &P, uint32_t path_lo, uint32_t path_hi, Real (&x)[kHalf],
                                                            const Shared &sh, const JobState &job, Real (*slots)[kThreads],
                                                            std::integer_sequence<int, kQ...>)
    {
        uint32_t w[4];
        philox4x32_10(path_lo, path_hi, (uint32_t)JB, kTagBasket, P.keys, w);
        Real z[kNpb];
        normals_from_words(w, z, sh, job);
        ((slots[JB * kNpb + kQ][0] = z[kQ]), ...);
        (column_rows<JB * kNpb + kQ, JB * kNpb + kQ, 0>(x, z[kQ], std::make_integer_sequence<int, kHalf - (JB * kNpb + kQ)>{}), ...);
    }
    // draw block JB of the second half: rows J .. N-1 (all in the upper half)
    template <int JB, int... kQ>
    static __device__ __forceinline__ void second_half_block(const Params &P, uint32_t path_lo, uint32_t path_hi, Real (&x)[kHalf],
                                                             const Shared &sh, const JobState &job, std::integer_sequence<int, kQ...>)
    {
        uint32_t w[4];
        philox4x32_10(path_lo, path_hi, (uint32_t)JB, kTagBasket, P.keys, w);
        Real z[kNpb];
        normals_from_words(w, z, sh, job);
        (column_rows<JB * kNpb + kQ, JB * kNpb + kQ, kHalf>(x, z[kQ], std::make_integer_sequence<int, N - (JB * kNpb + kQ)>{}), ...);
    }
    template <int... kJB>
    static __device__ __forceinline__ void first_half(const Params &P, uint32_t path_lo, uint32_t path_hi, Real (&x)[kHalf],
                                                      const Shared &sh, const JobState &job, Real (*slots)[kThreads],
                                                      std::integer_sequence<int, kJB...>)
    {
        (first_half_block<kJB>(P, path_lo, path_hi, x, sh, job, slots, std::make_integer_sequence<int, kNpb>{}), ...);
    }
    template <int... kJB>
    static __device__ __forceinline__ void second_half(const Params &P, uint32_t path_lo, uint32_t path_hi, Real (&x)[kHalf],
                                                       const Shared &sh, const JobState &job, std::integer_sequence<int, kJB...>)
    {
        (second_half_block<kHalf / kNpb + kJB>(P, path_lo, path_hi, x, sh, job, std::make_integer_sequence<int, kNpb>{}), ...);
    }
    // the dense block: parked normal J times rows kHalf .. N-1
    template <int... kJ>
    static __device__ __forceinline__ void parked_columns(Real (&x)[kHalf], Real (*slots)[kThreads], std::integer_sequence<int, kJ...>)
    {
        (column_rows<kJ, kHalf, kHalf>(x, slots[kJ][0], std::make_integer_sequence<int, kHalf>{}), ...);
    }
    template <int kFirst, int... kI>
    static __device__ __forceinline__ void half_init(Real (&x)[kHalf], std::integer_sequence<int, kI...>)
    {
        ((x[kI] = table_entry<Real, kABase + (kFirst + kI) * (int)sizeof(Real)>()), ...);
    }
    template <int kFirst, int... kI>
    static __device__ __forceinline__ Real half_value(const Real (&x)[kHalf], Real sum, const Shared &sh, std::integer_sequence<int, kI...>)
    {
        ((sum = fma(table_entry<Real, kMBase + (kFirst + kI) * (int)sizeof(Real)>(), grow(x[kI], sh), sum)), ...);
        return sum;
    }
    static __device__ __forceinline__ void eval_two_pass(const Params &P, uint32_t path_lo, uint32_t path_hi, Real (&v)[1],
                                                         const Shared &sh, const JobState &job)
    {
        static_assert(!kTwoPass || (kHalf % kNpb == 0 && N % kNpb == 0), "halves must fall on draw-block boundaries");
        static_assert(!kTwoPass || kSubBlocks == 2, "SharedTwoPass64 holds the slots of two sub-blocks");
        // this thread's column of slots: [normal][thread]
        Real (*slots)[kThreads] =
            reinterpret_cast<Real (*)[kThreads]>(const_cast<Real *>(&sh.stash[threadIdx.x / kThreads][0][threadIdx.x % kThreads]));
        using Half = std::make_integer_sequence<int, kHalf>;
        Real x[kHalf];
        half_init<0>(x, Half{});
        first_half(P, path_lo, path_hi, x, sh, job, slots, std::make_integer_sequence<int, kHalf / kNpb>{});
        Real sum = half_value<0>(x, -table_entry<Real, kKBase>(), sh, Half{});
        half_init<kHalf>(x, Half{});
        parked_columns(x, slots, Half{});
        second_half(P, path_lo, path_hi, x, sh, job, std::make_integer_sequence<int, (N - kHalf) / kNpb>{});
        v[0] = clamp(half_value<kHalf>(x, sum, sh, Half{}));
    }
    static __device__ __forceinline__ void eval(const Params &P, uint32_t path_lo, uint32_t path_hi, Real (&v)[1],
                                                const Shared &sh, const JobState &job)
    {
        if constexpr (kTwoPass) {
            eval_two_pass(P, path_lo, path_hi, v, sh, job);
            return;
        }
        State st;
        init(st, std::make_integer_sequence<int, (kPaired ? N / 2 : N)>{});
        sweep(P, path_lo, path_hi, st, sh, job, std::make_integer_sequence<int, kBlocks>{});
        v[0] = payoff(st, sh, std::make_integer_sequence<int, N>{});
    }
};

// ---- baskets wider than the widest register template (64 < n <= kBasketWideMax) ----
// The reference's N is a compile-time macro with no upper bound (DP/MonteCarlo.h:16); its kernel keeps g[], bt[], s[] in
// local memory whatever N is.  Here only baskets beyond 64 assets take a generic route: the path's normals
// go to a local-memory array once, and the column sweep runs over row blocks of kBasketWideRows accumulators in registers
// -- per block: z_j from local memory, kBasketWideRows factor entries per column from a device-memory table laid out in
// exactly that order (uniform 16-byte loads).  Same summation order as the register templates (columns 0, 1, 2, ... into
// each exponent, assets 0, 1, 2, ... into the payoff), so a narrow basket forced through this kernel gives their bits.
constexpr int kBasketWideMax = 256;
constexpr int kBasketWideRows = 16;

template <typename RealT, bool kAccumLayout = false>
struct BasketWide {
    using Real = RealT;
    static constexpr int kUnitPaths = 1;
    static constexpr int kUnroll = 1;
    static constexpr int kSubBlocks = (kAccumLayout && sizeof(RealT) == 8) ? 2 : 1;
    static constexpr int kMinBlocks = kSubBlocks > 1 ? 1 : 2;
    static constexpr int kNpb = NormalsPerBlock<RealT>::value;
    static constexpr int kRows = kBasketWideRows;
    struct Params {
        PhiloxKeys keys;
        int n_blocks;        // row blocks: ceil(n / kRows)
        int full;            // the factor has entries above the diagonal: every block sweeps all columns
        const Real *table;   // device memory, see fill_wide_table
        Real k;
    };
    using Shared = std::conditional_t<kAccumLayout, typename SharedAccumFor<Real>::type, typename SharedFor<Real>::type>;
    using JobState = typename JobStateFor<Real>::type;
    static constexpr bool kClampAtZero = sizeof(Real) == 8;
    static __device__ __forceinline__ void prepare(const Params &, JobState &job, int tid) { prepare_polar(polar_scale<Real>(1.0), job, tid); }
    static __device__ __forceinline__ void eval(const Params &P, uint32_t path_lo, uint32_t path_hi, Real (&v)[1], const Shared &sh,
                                                const JobState &job)
    {
        const int n_pad = P.n_blocks * kRows;
        Real z[kBasketWideMax + kNpb];          // local memory: indexed at run time
        for (int jb = 0; jb * kNpb < n_pad; jb++) {
            uint32_t w[4];
            philox4x32_10(path_lo, path_hi, (uint32_t)jb, kTagBasket, P.keys, w);
            Real zz[kNpb];
            normals_from_words(w, zz, sh, job);
#pragma unroll
            for (int q = 0; q < kNpb; q++)
                z[jb * kNpb + q] = zz[q];
        }
        const Real *f = P.table;                // [block][column][kRows] factor entries, then a[n_pad], then m[n_pad]
        const Real *a = P.table + wide_factor_entries(P.n_blocks, P.full != 0);
        const Real *m = a + n_pad;
        Real sum = -P.k;
        for (int rb = 0; rb < P.n_blocks; rb++) {
            Real x[kRows];
#pragma unroll
            for (int r = 0; r < kRows; r++)
                x[r] = __ldg(a + rb * kRows + r);
            const int cols = P.full ? n_pad : (rb + 1) * kRows;
            for (int j = 0; j < cols; j++) {
                const Real zj = z[j];
#pragma unroll
                for (int r = 0; r < kRows; r++)
                    x[r] = fma(__ldg(f + r), zj, x[r]);
                f += kRows;
            }
#pragma unroll
            for (int r = 0; r < kRows; r++)
                sum = fma(__ldg(m + rb * kRows + r), exp_scaled<true>(x[r], sh), sum);
        }
        if constexpr (kClampAtZero)
            v[0] = sum;
        else
            v[0] = positive_part(sum);
    }
    static __host__ __device__ constexpr size_t wide_factor_entries(int n_blocks, bool full)
    {
        // triangular: block rb sweeps columns 0 .. (rb + 1) kRows - 1
        return full ? (size_t)n_blocks * (size_t)(n_blocks * kRows) * kRows
                    : (size_t)kRows * kRows * ((size_t)n_blocks * (size_t)(n_blocks + 1) / 2);
    }
};

}  // namespace mcb

#include "basket_tc.cuh"

namespace mcb {

// Host: narrow the fp64 job into the kernel's table.  Assets beyond job.n (padding up to the
// template width) get weight 0 and a zero factor row/column: they add exactly 0 to the payoff.
template <typename Real, int N, bool kFull>
static void fill_table(const BasketJob &job, BasketTable<Real, N, kFull> &T)
{
    using Table = BasketTable<Real, N, kFull>;
    const double unit = ExpUnit<Real>::value;  // exponents in the units of exp_scaled: 1/ln2 (MUFU.EX2) or 256/ln2 (exp_units)
    for (int i = 0; i < Table::kFactor; i++)
        T.factor[i] = 0;
    for (int col = 0; col < N; col++)
        for (int row = Table::first_row(col); row < N; row++) {
            const double f = (row < job.n && col < job.n) ? job.factor[row * job.n + col] : 0.0;
            T.factor[Table::index(col, row)] = (Real)(f * unit);
        }
    for (int i = 0; i < N; i++) {
        T.a[i] = (Real)(i < job.n ? job.a[i] * unit : 0.0);
        T.m[i] = (Real)(i < job.n ? job.m[i] : 0.0);
    }
    T.k = (Real)job.k;
}

template <typename Real, int N, bool kFull>
static cudaError_t launch_t(const BasketJob &job, const Geometry *geom, int grid, unsigned long long *d_acc,
                            unsigned long long first_unit, unsigned long long n_units, void *d_out,
                            cudaStream_t stream, const LaunchOptions &opt)
{
    using W = Basket<Real, N, kFull>;
    using WA = Basket<Real, N, kFull, true>;
    std::vector<unsigned char> staging(sizeof(typename W::Table));
    fill_table(job, *reinterpret_cast<typename W::Table *>(staging.data()));
    typename W::Params p;
    p.keys = job.keys;
    typename WA::Params pa;
    pa.keys = job.keys;
    TableUse use(g_basket_lock, stream, staging.data(), staging.size());
    if (use.status() != cudaSuccess)
        return use.status();
    if (use.needs_upload()) {
        cudaError_t e = cudaMemcpyToSymbolAsync(mcb_basket_table, staging.data(), staging.size(), 0,
                                                cudaMemcpyHostToDevice, stream);
        if (e != cudaSuccess) {
            use.invalidate();
            return e;
        }
        use.uploaded();
    }
    cudaError_t e;
    if (geom)
        e = accumulate_launch<WA>(grid, pa, *geom, d_acc, stream, opt);
    else {
        const unsigned long long blocks = (n_units + kThreads - 1) / kThreads;
        mc_paths_kernel<W><<<(int)(blocks < 65535ull ? blocks : 65535ull), kThreads, 0, stream>>>(p, first_unit, n_units,
                                                                                                   (Real *)d_out);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess)
        use.invalidate();
    return e;
}

template <typename Real, int N, bool kFull>
static int occupancy_t()
{
    return accumulate_blocks_per_sm<Basket<Real, N, kFull, true>>();
}

// ---- tensor-core engine (basket_tc.cuh): fp32, 32 < n <= 64 ----
static std::atomic<int> g_basket_engine{-1};  // -1: not set (environment decides), 0: auto, 1: FFMA only

int basket_engine_get()
{
    int e = g_basket_engine.load();
    if (e < 0) {
        const char *env = std::getenv("MCB200_BASKET_ENGINE");
        e = (env && (env[0] == '1' || env[0] == 'f' || env[0] == 'F')) ? 1 : 0;
        g_basket_engine.store(e);
    }
    return e;
}
void basket_engine_set(int engine) { g_basket_engine.store(engine == 1 ? 1 : 0); }

static bool basket_takes_wide_route(int n);
bool basket_uses_tensor_cores(int precision, int n)
{
    return precision == 0 && n > 32 && n <= kTcWidth && basket_engine_get() == 0 && !basket_takes_wide_route(n);
}

static void fill_tc_table(const BasketJob &job, BasketTcTable &T)
{
    const double unit = 1.4426950408889634074;  // log2(e): the exponents feed MUFU.EX2
    for (int i = 0; i < kTcWidth; i++) {
        for (int k = 0; k < kTcWidth; k++)
            T.f[i * kTcWidth + k] = (float)((i < job.n && k < job.n) ? job.factor[i * job.n + k] * unit : 0.0);
        T.a[i] = (float)(i < job.n ? job.a[i] * unit : 0.0);
        T.m[i] = (float)(i < job.n ? job.m[i] : 0.0);
    }
    T.k = (float)job.k;
}

template <bool kFull>
static cudaError_t launch_tc(const BasketJob &job, const Geometry *geom, int grid, unsigned long long *d_acc,
                             unsigned long long first_unit, unsigned long long n_units, void *d_out, cudaStream_t stream,
                             const LaunchOptions &opt)
{
    static_assert(sizeof(BasketTcTable) <= kBasketTableBytes, "tensor-core basket table exceeds its constant buffer");
    std::vector<unsigned char> staging(sizeof(BasketTcTable) + 1);
    fill_tc_table(job, *reinterpret_cast<BasketTcTable *>(staging.data()));
    staging[sizeof(BasketTcTable)] = 0x7c;  // layout tag: never equal to an FFMA-engine image of the same size
    BasketTcParams p;
    p.keys = job.keys;
    TableUse use(g_basket_lock, stream, staging.data(), staging.size());
    if (use.status() != cudaSuccess)
        return use.status();
    if (use.needs_upload()) {
        cudaError_t e = cudaMemcpyToSymbolAsync(mcb_basket_table, staging.data(), sizeof(BasketTcTable), 0,
                                                cudaMemcpyHostToDevice, stream);
        if (e != cudaSuccess) {
            use.invalidate();
            return e;
        }
        use.uploaded();
    }
    cudaError_t e;
    if (geom) {
        e = launch_kernel(basket_tc_accumulate_kernel<kFull>, grid, kTcThreads, 0, stream, opt.overlap, p, *geom, d_acc);
    } else {
        const unsigned long long blocks = (n_units + kThreads - 1) / kThreads;
        basket_tc_paths_kernel<kFull><<<(int)((blocks + 1) / 2 < 148ull ? (blocks + 1) / 2 : 148ull), kTcThreads, 0, stream>>>(p, first_unit, n_units,
                                                                                                       (float *)d_out);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess)
        use.invalidate();
    return e;
}

// ---- host side of the wide route: the table image in the kernel's order, one growable device buffer per device ----
static TableLock g_basket_wide_lock;
static void *g_basket_wide_buffer[TableLock::kMaxDevices] = {};
static size_t g_basket_wide_capacity[TableLock::kMaxDevices] = {};

template <typename Real>
static std::vector<Real> fill_wide_table(const BasketJob &job, int n_blocks)
{
    using W = BasketWide<Real>;
    const double unit = ExpUnit<Real>::value;
    const int rows = W::kRows, n_pad = n_blocks * rows;
    std::vector<Real> t;
    t.reserve(W::wide_factor_entries(n_blocks, job.full) + 2 * (size_t)n_pad);
    for (int rb = 0; rb < n_blocks; rb++) {
        const int cols = job.full ? n_pad : (rb + 1) * rows;
        for (int j = 0; j < cols; j++)
            for (int r = 0; r < rows; r++) {
                const int i = rb * rows + r;
                const bool inside = i < job.n && j < job.n && (job.full || j <= i);
                t.push_back((Real)(inside ? job.factor[(size_t)i * job.n + j] * unit : 0.0));
            }
    }
    for (int i = 0; i < n_pad; i++)
        t.push_back((Real)(i < job.n ? job.a[i] * unit : 0.0));
    for (int i = 0; i < n_pad; i++)
        t.push_back((Real)(i < job.n ? job.m[i] : 0.0));      // padding assets: weight 0, exactly 0 contribution
    return t;
}

template <typename Real>
static cudaError_t launch_wide_t(const BasketJob &job, const Geometry *geom, int grid, unsigned long long *d_acc,
                                 unsigned long long first_unit, unsigned long long n_units, void *d_out, cudaStream_t stream,
                                 const LaunchOptions &opt)
{
    const int n_blocks = (job.n + kBasketWideRows - 1) / kBasketWideRows;
    const std::vector<Real> staging = fill_wide_table<Real>(job, n_blocks);
    const size_t bytes = staging.size() * sizeof(Real);
    TableUse use(g_basket_wide_lock, stream, staging.data(), bytes);
    if (use.status() != cudaSuccess)
        return use.status();
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess)
        return e;
    if (g_basket_wide_capacity[device] < bytes) {
        use.invalidate();   // (cudaFree waits for everything that may still read the old buffer)
        if (g_basket_wide_buffer[device])
            cudaFree(g_basket_wide_buffer[device]);
        g_basket_wide_buffer[device] = nullptr;
        g_basket_wide_capacity[device] = 0;
        e = cudaMalloc(&g_basket_wide_buffer[device], bytes);
        if (e != cudaSuccess)
            return e;
        g_basket_wide_capacity[device] = bytes;
    }
    if (use.needs_upload()) {
        e = cudaMemcpyAsync(g_basket_wide_buffer[device], staging.data(), bytes, cudaMemcpyHostToDevice, stream);
        if (e != cudaSuccess) {
            use.invalidate();
            return e;
        }
        use.uploaded();
    }
    auto params = [&](auto tag) {
        typename decltype(tag)::Params p;
        p.keys = job.keys;
        p.n_blocks = n_blocks;
        p.full = job.full ? 1 : 0;
        p.table = static_cast<const Real *>(g_basket_wide_buffer[device]);
        p.k = (Real)job.k;
        return p;
    };
    if (geom) {
        using W = BasketWide<Real, true>;
        e = accumulate_launch<W>(grid, params(W{}), *geom, d_acc, stream, opt);
    } else {
        using W = BasketWide<Real, false>;
        const unsigned long long blocks = (n_units + kThreads - 1) / kThreads;
        mc_paths_kernel<W><<<(int)(blocks < 65535ull ? blocks : 65535ull), kThreads, 0, stream>>>(params(W{}), first_unit, n_units, (Real *)d_out);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess)
        use.invalidate();
    return e;
}

// MCB200_BASKET_WIDE=1 sends every basket through the wide route (tests: its bits against the register templates')
static bool basket_takes_wide_route(int n)
{
    static const bool forced = [] {
        const char *env = std::getenv("MCB200_BASKET_WIDE");
        return env && env[0] == '1';
    }();
    return n > 64 || forced;
}

int basket_max_width() { return kBasketWideMax; }

int basket_padded_width(int n)
{
    static const int widths[] = {3, 4, 8, 10, 16, 32, 64};
    if (n < 1)
        return 0;
    for (int w : widths)
        if (n <= w)
            return w;
    return 0;
}

// (precision, width, full) -> template instance
#define MCB_BASKET_DISPATCH(CALL)                                                          \
    switch (basket_padded_width(n) * 4 + (precision ? 2 : 0) + (full ? 1 : 0)) {           \
        case 3 * 4 + 0: return CALL(float, 3, false);                                      \
        case 3 * 4 + 1: return CALL(float, 3, true);                                       \
        case 3 * 4 + 2: return CALL(double, 3, false);                                     \
        case 3 * 4 + 3: return CALL(double, 3, true);                                      \
        case 4 * 4 + 0: return CALL(float, 4, false);                                      \
        case 4 * 4 + 1: return CALL(float, 4, true);                                       \
        case 4 * 4 + 2: return CALL(double, 4, false);                                     \
        case 4 * 4 + 3: return CALL(double, 4, true);                                      \
        case 8 * 4 + 0: return CALL(float, 8, false);                                      \
        case 8 * 4 + 1: return CALL(float, 8, true);                                       \
        case 8 * 4 + 2: return CALL(double, 8, false);                                     \
        case 8 * 4 + 3: return CALL(double, 8, true);                                      \
        case 10 * 4 + 0: return CALL(float, 10, false);                                    \
        case 10 * 4 + 1: return CALL(float, 10, true);                                     \
        case 10 * 4 + 2: return CALL(double, 10, false);                                   \
        case 10 * 4 + 3: return CALL(double, 10, true);                                    \
        case 16 * 4 + 0: return CALL(float, 16, false);                                    \
        case 16 * 4 + 1: return CALL(float, 16, true);                                     \
        case 16 * 4 + 2: return CALL(double, 16, false);                                   \
        case 16 * 4 + 3: return CALL(double, 16, true);                                    \
        case 32 * 4 + 0: return CALL(float, 32, false);                                    \
        case 32 * 4 + 1: return CALL(float, 32, true);                                     \
        case 32 * 4 + 2: return CALL(double, 32, false);                                   \
        case 32 * 4 + 3: return CALL(double, 32, true);                                    \
        case 64 * 4 + 0: return CALL(float, 64, false);                                    \
        case 64 * 4 + 1: return CALL(float, 64, true);                                     \
        case 64 * 4 + 2: return CALL(double, 64, false);                                   \
        case 64 * 4 + 3: return CALL(double, 64, true);                                    \
        default: break;                                                                    \
    }

int basket_blocks_per_sm(int precision, int n, bool full)
{
    if (basket_takes_wide_route(n))
        return precision ? accumulate_blocks_per_sm<BasketWide<double, true>>() : accumulate_blocks_per_sm<BasketWide<float, true>>();
    if (basket_uses_tensor_cores(precision, n))
        return 1;  // one CTA per SM: four tiles = all 512 tensor-memory columns

#define MCB_OCC(R, W, F) occupancy_t<R, W, F>()
    MCB_BASKET_DISPATCH(MCB_OCC)
#undef MCB_OCC
    return 0;
}

cudaError_t basket_launch(int precision, const BasketJob &job, const Geometry &geom, int grid,
                          unsigned long long *d_acc, cudaStream_t stream, const LaunchOptions &opt)
{
    const int n = job.n;
    const bool full = job.full;
    if (basket_takes_wide_route(n))
        return precision ? launch_wide_t<double>(job, &geom, grid, d_acc, 0ull, 0ull, nullptr, stream, opt)
                         : launch_wide_t<float>(job, &geom, grid, d_acc, 0ull, 0ull, nullptr, stream, opt);
    if (basket_uses_tensor_cores(precision, n))
        return full ? launch_tc<true>(job, &geom, grid, d_acc, 0ull, 0ull, nullptr, stream, opt)
                    : launch_tc<false>(job, &geom, grid, d_acc, 0ull, 0ull, nullptr, stream, opt);
#define MCB_LAUNCH(R, W, F) launch_t<R, W, F>(job, &geom, grid, d_acc, 0ull, 0ull, nullptr, stream, opt)
    MCB_BASKET_DISPATCH(MCB_LAUNCH)
#undef MCB_LAUNCH
    return cudaErrorInvalidValue;
}

cudaError_t basket_paths(int precision, const BasketJob &job, unsigned long long first_unit,
                         unsigned long long n_units, void *d_out, cudaStream_t stream)
{
    const int n = job.n;
    const bool full = job.full;
    if (basket_takes_wide_route(n))
        return precision ? launch_wide_t<double>(job, nullptr, 0, nullptr, first_unit, n_units, d_out, stream, LaunchOptions())
                         : launch_wide_t<float>(job, nullptr, 0, nullptr, first_unit, n_units, d_out, stream, LaunchOptions());
    if (basket_uses_tensor_cores(precision, n))
        return full ? launch_tc<true>(job, nullptr, 0, nullptr, first_unit, n_units, d_out, stream, LaunchOptions())
                    : launch_tc<false>(job, nullptr, 0, nullptr, first_unit, n_units, d_out, stream, LaunchOptions());
#define MCB_PATHS(R, W, F) launch_t<R, W, F>(job, nullptr, 0, nullptr, first_unit, n_units, d_out, stream, LaunchOptions())
    MCB_BASKET_DISPATCH(MCB_PATHS)
#undef MCB_PATHS
    return cudaErrorInvalidValue;
}

}  // namespace mcb
