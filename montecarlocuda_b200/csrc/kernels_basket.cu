// kernels_basket.cu -- basket call on N Cholesky-correlated underlyings, fp32 and fp64 (sm_100a).
//
// Replaces brownianVect + basketPayoff + basketOptMonteCarlo (DP/MonteCarloKernel.cu:74-101,
// :133-177).  One draw unit = one path; draw block j of the path's sub-stream gives normals
// 2j, 2j+1 (fp64) or 4j .. 4j+3 (fp32).
//   x_i    = a_i + sum_j F_ij z_j      F_ij = v_i sqrt(T) L_ij,  a_i = (r - v_i^2/2) T + v_i sqrt(T) d_i
//   payoff = max(sum_i m_i e^{x_i} - K, 0),  m_i = w_i s_i      (fp32: x in log2 units, 2^x by MUFU.EX2)
// The mat-vec is a column sweep kept in registers: normal j is produced, applied to the N - j
// accumulators at or below the diagonal and discarded, so N accumulators + one normal are live
// and every factor entry reaches the FMA pipe as a constant-bank operand.  The reference indexes
// g[], bt[], s[] with runtime bounds, which puts them in local memory (LDL/STL),
// and multiplies the zero upper triangle too.  kFull keeps that behaviour for a caller whose p
// is not triangular.
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

template <typename Real> struct NormalsPerBlock;
template <> struct NormalsPerBlock<float> { static constexpr int value = 4; };
template <> struct NormalsPerBlock<double> { static constexpr int value = 2; };

template <typename Real, int N, bool kFull>
struct BasketTable {
    static constexpr int kFactor = kFull ? N * N : N * (N + 1) / 2;
    Real factor[kFactor];  // column-major; packed lower triangle unless kFull
    Real a[N];
    Real m[N];
    Real k;
    static __host__ __device__ constexpr int index(int col, int row)
    {
        return kFull ? col * N + row : col * N - col * (col - 1) / 2 + (row - col);
    }
};

constexpr int basket_min_blocks(int n, int real_bytes)
{
    // registers: the accumulators + ~44 for generator, normals and loop state
    const int regs = n * real_bytes / 4 + 44;
    const int blocks = 65536 / (kThreads * regs);
    return blocks < 1 ? 1 : (blocks > 4 ? 4 : blocks);
}

template <typename RealT, int N, bool kFull>
struct Basket {
    using Real = RealT;
    using Table = BasketTable<Real, N, kFull>;
    static_assert(sizeof(Table) <= kBasketTableBytes, "basket table exceeds its constant buffer");
    static constexpr int kUnitPaths = 1;
    static constexpr int kUnroll = 1;
    static constexpr int kMinBlocks = basket_min_blocks(N, (int)sizeof(Real));
    static constexpr int kNpb = NormalsPerBlock<Real>::value;
    struct Params {
        PhiloxKeys keys;
    };
    using Shared = typename SharedFor<Real>::type;
    static __device__ __forceinline__ float grow(float x, const NoShared &) { return mufu_ex2(x); }
    static __device__ __forceinline__ double grow(double x, const SharedTables64 &sh) { return exp_tab(x, sh.t); }
    static constexpr int kBlocks = (N + kNpb - 1) / kNpb;
    static constexpr int kFactorBase = 0;
    static constexpr int kABase = Table::kFactor * (int)sizeof(Real);
    static constexpr int kMBase = kABase + N * (int)sizeof(Real);
    static constexpr int kKBase = kMBase + N * (int)sizeof(Real);

    // column J of the sweep: x[row] += F[row][J] * z for row = first .. N-1
    template <int J, int... kRow>
    static __device__ __forceinline__ void column(Real (&x)[N], Real z, std::integer_sequence<int, kRow...>)
    {
        constexpr int first = kFull ? 0 : J;
        ((x[first + kRow] =
              fma(table_entry<Real, kFactorBase + Table::index(J, first + kRow) * (int)sizeof(Real)>(), z,
                  x[first + kRow])),
         ...);
    }
    template <int J>
    static __device__ __forceinline__ void column_if(Real (&x)[N], Real z)
    {
        if constexpr (J < N)
            column<J>(x, z, std::make_integer_sequence<int, (kFull ? N : N - J)>{});
    }
    // draw block JB: one Philox block -> kNpb normals -> kNpb columns
    template <int JB>
    static __device__ __forceinline__ void draw_block(const Params &P, unsigned long long path, Real (&x)[N],
                                                      const Shared &sh)
    {
        uint32_t w[4];
        philox4x32_10((uint32_t)path, (uint32_t)(path >> 32), (uint32_t)JB, kTagBasket, P.keys, w);
        Real z[kNpb];
        normals_from_words(w, z, sh);
        column_if<JB * kNpb + 0>(x, z[0]);
        column_if<JB * kNpb + 1>(x, z[1]);
        if constexpr (kNpb == 4) {
            column_if<JB * kNpb + 2>(x, z[2]);
            column_if<JB * kNpb + 3>(x, z[3]);
        }
    }
    template <int... kJB>
    static __device__ __forceinline__ void sweep(const Params &P, unsigned long long path, Real (&x)[N],
                                                 const Shared &sh, std::integer_sequence<int, kJB...>)
    {
        (draw_block<kJB>(P, path, x, sh), ...);
    }
    template <int... kI>
    static __device__ __forceinline__ void init(Real (&x)[N], std::integer_sequence<int, kI...>)
    {
        ((x[kI] = table_entry<Real, kABase + kI * (int)sizeof(Real)>()), ...);
    }
    template <int... kI>
    static __device__ __forceinline__ Real payoff(const Real (&x)[N], const Shared &sh,
                                                  std::integer_sequence<int, kI...>)
    {
        Real sum = -table_entry<Real, kKBase>();
        ((sum = fma(table_entry<Real, kMBase + kI * (int)sizeof(Real)>(), grow(x[kI], sh), sum)), ...);
        return positive_part(sum);
    }
    static __device__ __forceinline__ void eval(const Params &P, unsigned long long path, Real (&v)[1],
                                                const Shared &sh)
    {
        Real x[N];
        init(x, std::make_integer_sequence<int, N>{});
        sweep(P, path, x, sh, std::make_integer_sequence<int, kBlocks>{});
        v[0] = payoff(x, sh, std::make_integer_sequence<int, N>{});
    }
};

// Host: narrow the fp64 job into the kernel's table.  Assets beyond job.n (padding up to the
// template width) get weight 0 and a zero factor row/column: they add exactly 0 to the payoff.
template <typename Real, int N, bool kFull>
static void fill_table(const BasketJob &job, BasketTable<Real, N, kFull> &T)
{
    using Table = BasketTable<Real, N, kFull>;
    const double unit = sizeof(Real) == 4 ? 1.4426950408889634074 : 1.0;  // log2(e) for the MUFU.EX2 path
    for (int i = 0; i < Table::kFactor; i++)
        T.factor[i] = 0;
    for (int col = 0; col < N; col++)
        for (int row = (kFull ? 0 : col); row < N; row++) {
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
                            cudaStream_t stream)
{
    using W = Basket<Real, N, kFull>;
    std::vector<unsigned char> staging(sizeof(typename W::Table));
    fill_table(job, *reinterpret_cast<typename W::Table *>(staging.data()));
    typename W::Params p;
    p.keys = job.keys;
    TableUse use(g_basket_lock, stream);
    if (use.status() != cudaSuccess)
        return use.status();
    cudaError_t e = cudaMemcpyToSymbolAsync(mcb_basket_table, staging.data(), staging.size(), 0,
                                            cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess)
        return e;
    if (geom) {
        mc_accumulate_kernel<W><<<grid, kThreads, 0, stream>>>(p, *geom, d_acc);
    } else {
        const unsigned long long blocks = (n_units + kThreads - 1) / kThreads;
        mc_paths_kernel<W><<<(int)(blocks < 65535ull ? blocks : 65535ull), kThreads, 0, stream>>>(
            p, first_unit, n_units, (Real *)d_out);
    }
    return cudaGetLastError();
}

template <typename Real, int N, bool kFull>
static int occupancy_t()
{
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, mc_accumulate_kernel<Basket<Real, N, kFull>>, kThreads,
                                                      0) != cudaSuccess)
        return 0;
    return n;
}

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
#define MCB_OCC(R, W, F) occupancy_t<R, W, F>()
    MCB_BASKET_DISPATCH(MCB_OCC)
#undef MCB_OCC
    return 0;
}

cudaError_t basket_launch(int precision, const BasketJob &job, const Geometry &geom, int grid,
                          unsigned long long *d_acc, cudaStream_t stream)
{
    const int n = job.n;
    const bool full = job.full;
#define MCB_LAUNCH(R, W, F) launch_t<R, W, F>(job, &geom, grid, d_acc, 0ull, 0ull, nullptr, stream)
    MCB_BASKET_DISPATCH(MCB_LAUNCH)
#undef MCB_LAUNCH
    return cudaErrorInvalidValue;
}

cudaError_t basket_paths(int precision, const BasketJob &job, unsigned long long first_unit,
                         unsigned long long n_units, void *d_out, cudaStream_t stream)
{
    const int n = job.n;
    const bool full = job.full;
#define MCB_PATHS(R, W, F) launch_t<R, W, F>(job, nullptr, 0, nullptr, first_unit, n_units, d_out, stream)
    MCB_BASKET_DISPATCH(MCB_PATHS)
#undef MCB_PATHS
    return cudaErrorInvalidValue;
}

}  // namespace mcb
