// kernels_cva.cu -- CVA of one European call under Black-Scholes, fp32 and fp64 (sm_100a).
//
// Replaces geomBrownian + cnd + device_bsCall + cvaCallOptMC (DP/MonteCarloKernel.cu:104-129,
// :222-283).  One draw unit = one path; exposure date j uses normal j of the path's sub-stream.
// The state is the log-moneyness y = ln(S/K).  Per kept date j (host-built table, engine.cu):
//   y  += mu_dt + sig_dt z                      exact GBM step
//   s   = K e^y
//   d1  = y inv_j + c1_j,  d2 = d1 - sig_j      inv_j = 1/(v sqrt(tau_j)), c1_j = (r + v^2/2) tau_j inv_j
//   A   = s phi(d1) = K/sqrt(2 pi) e^{y - d1^2/2}  one exponential; kd_j phi(d2) = A as well (kd_j = K e^{-r tau_j})
//   q_i = cnd-tail(|d_i|) / phi(d_i)             Hastings polynomial in 1 / (1 + 0.2316419 |d_i|), like the reference
//   ee  = s cnd(d1) - kd_j cnd(d2)
//       = s [d1 > 0] - kd_j [d2 > 0] - A (sgn(d1) q1 - sgn(d2) q2)
//   cva += w_j ee                               w_j = LGD (e^{-lambda t_{j-1}} - e^{-lambda t_j})
// so a path-step costs one normal, two exponentials and two reciprocals; the reference spends six
// exponentials, a log, three square roots and four divisions on it.  Dates the reference drops
// (remaining time rounded below zero) are simply absent from the table.
#include <type_traits>
#include <vector>

#include "device_math.cuh"
#include "launch.h"
#include "table_lock.h"

#ifndef MCB_CVA_SUBBLOCKS
// sub-blocks of 256 threads around one table set.  2: 128 registers per thread, 17.41 ms; 3: 80 registers, 17.89 ms
// (profiles/r01p_ab_experiments.txt; tools/build_variant.sh builds the other one for A/B timing)
#define MCB_CVA_SUBBLOCKS 2
#endif
#ifndef MCB_CVA_BANK
#define MCB_CVA_BANK 0  // fp64 math constants as constant-bank operands (1) or literals (0): at 128 registers literals win, 16.92 vs 17.01 ms
#endif

namespace mcb {

template <typename Real>
struct alignas(16) CvaDate {  // three 16-byte constant loads per date
    Real w, inv, c1, sig, kd, pad;
};

__constant__ __align__(16) unsigned char c_cva_table[kCvaMaxDates * sizeof(CvaDate<double>)];
static TableLock g_cva_lock;

// q(d) = cnd-tail(|d|) / phi(d) = k polynomial(k), k = 1 / (1 + 0.2316419 |d|)
// (Abramowitz-Stegun 26.2.17, the constants of DP/MonteCarloKernel.cu:111-116)
template <typename Real, bool kBank = false>
__device__ __forceinline__ Real hastings_ratio(Real d)
{
    if constexpr (kBank) {
        const MathConsts64 &K = kMathConsts64;  // constant-bank operands instead of re-materialised immediates
        const double k = rcp_real(fma(K.hast_k, fabs(d), 1.0));
        double poly = fma(k, K.hast_a5, K.hast_a4);
        poly = fma(k, poly, K.hast_a3);
        poly = fma(k, poly, K.hast_a2);
        poly = fma(k, poly, K.hast_a1);
        return k * poly;
    } else {
        const Real k = rcp_real(fma((Real)0.2316419, fabs(d), (Real)1.0));
        Real poly = fma(k, (Real)1.330274429, (Real)-1.821255978);
        poly = fma(k, poly, (Real)1.781477937);
        poly = fma(k, poly, (Real)-0.356563782);
        poly = fma(k, poly, (Real)0.31938153);
        return k * poly;
    }
}

// sgn(d) q for q > 0 and x [d > 0], by the sign bit of d on the integer pipe (the fp64 comparisons are
// DSETPs on the pipe that binds).  d = +0 counts as positive where the reference's `d > 0` does not:
// cnd(0) is 1/2 from either branch.
__device__ __forceinline__ double with_sign_of(double q, double d)
{
    return __hiloint2double(__double2hiint(q) ^ (__double2hiint(d) & (int)0x80000000), __double2loint(q));
}
__device__ __forceinline__ float with_sign_of(float q, float d)
{
    return __int_as_float(__float_as_int(q) ^ (__float_as_int(d) & (int)0x80000000));
}
__device__ __forceinline__ double keep_if_positive(double x, double d)
{
    const int keep = ~(__double2hiint(d) >> 31);
    return __hiloint2double(__double2hiint(x) & keep, __double2loint(x) & keep);
}
__device__ __forceinline__ float keep_if_positive(float x, float d)
{
    return __int_as_float(__float_as_int(x) & ~(__float_as_int(d) >> 31));
}

// max(a, -700) without touching the fp64 pipe: negative doubles order like their high words taken as
// unsigned integers and every non-negative one lies below them, so one integer min on the high word does
// it (-inf included).
__device__ __forceinline__ double floor_at_minus_700(double a)
{
    const unsigned hi = min((unsigned)__double2hiint(a), 0xC085E000u);  // high word of -700.0
    return __hiloint2double((int)hi, hi == 0xC085E000u ? 0 : __double2loint(a));
}
__device__ __forceinline__ float floor_at_minus_700(float a) { return a; }  // MUFU.EX2(-inf) = 0

template <typename RealT, bool kAccumLayout = false>
struct Cva {
    using Real = RealT;
    static constexpr int kUnitPaths = 1;
    static constexpr int kUnroll = 1;
    static constexpr int kSubBlocks = (kAccumLayout && sizeof(RealT) == 8) ? MCB_CVA_SUBBLOCKS : 1;  // fp64: one table set per SM
    static constexpr int kMinBlocks = kSubBlocks > 1 ? 1 : 3;
    static constexpr int kNpb = NormalsPerBlock<RealT>::value;
    struct Params {
        PhiloxKeys keys;
        Real y0, mu_dt, k, k_pdf;  // k_pdf = K / sqrt(2 pi)
        PolarScale<Real> scale;  // of sig_dt = v sqrt(dt), folded under the Box-Muller square root
        int n_dates;  // kept dates
        int first_date;  // of this job in the device table (0 unless the launch carries several jobs)
    };
    // fp64 pricing kernel: replicated tables; math constants from the constant bank only when asked (MCB_CVA_BANK: the
    // right choice under the 80-register cap of 3 sub-blocks, see device_math64.cuh)
    static constexpr bool kBank = kAccumLayout && sizeof(RealT) == 8 && MCB_CVA_BANK;
    using Shared = std::conditional_t<kBank, SharedTables64RepBank,
                                      std::conditional_t<kAccumLayout, typename SharedAccumFor<Real>::type, typename SharedFor<Real>::type>>;
    // one exposure date; the diffusion sig_dt z arrives as (sig_dt r) * (cos or sin) and folds into the step's FMA
    static __device__ __forceinline__ void step(const Params &P, const CvaDate<Real> &D, Real sr, Real trig, Real &y,
                                                Real &cva, const Shared &sh)
    {
        y = fma(sr, trig, y + P.mu_dt);
        const Real s = P.k * exp_real(y, sh);
        const Real d1 = fma(y, D.inv, D.c1);
        const Real d2 = d1 - D.sig;
        // s phi(d1) = kd phi(d2) = K / sqrt(2 pi) e^{y - d1^2/2}.  The exponent can be -1e14 (a date a few ulps before
        // maturity) or -inf (exact grid, tau = 0): floor it where e^x is already 0 for every purpose, so the
        // table-driven exp stays in range
        const Real a = P.k_pdf * exp_real(floor_at_minus_700(fma((Real)-0.5 * d1, d1, y)), sh);
        const Real tails = with_sign_of(hastings_ratio<Real, kBank>(d1), d1) - with_sign_of(hastings_ratio<Real, kBank>(d2), d2);
        const Real ee = fma(-a, tails, keep_if_positive(s, d1) - keep_if_positive(D.kd, d2));
        cva = fma(D.w, ee, cva);
    }
    static __device__ __forceinline__ void eval(const Params &P, uint32_t path_lo, uint32_t path_hi, Real (&v)[1],
                                                const Shared &sh)
    {
        const CvaDate<Real> *dates = reinterpret_cast<const CvaDate<Real> *>(c_cva_table) + P.first_date;
        Real y = P.y0, cva = 0;
        // whole draw blocks run their kNpb dates back to back with no test in between: only y links one date to the
        // next, so the exponentials and reciprocals of neighbouring dates overlap
        const int n_whole = P.n_dates / kNpb;
#pragma unroll 1
        for (int jb = 0; jb < n_whole; jb++) {
            uint32_t w[4];
            philox4x32_10(path_lo, path_hi, (uint32_t)jb, kTagCva, P.keys, w);
            Real sr[kNpb / 2], cs[kNpb / 2], sn[kNpb / 2];
            polar_from_words<true>(w, sr, cs, sn, sh, P.scale);
#pragma unroll
            for (int q = 0; q < kNpb; q++)
                step(P, dates[jb * kNpb + q], sr[q / 2], (q & 1) ? sn[q / 2] : cs[q / 2], y, cva, sh);
        }
        if (n_whole * kNpb < P.n_dates) {
            uint32_t w[4];
            philox4x32_10(path_lo, path_hi, (uint32_t)n_whole, kTagCva, P.keys, w);
            Real sr[kNpb / 2], cs[kNpb / 2], sn[kNpb / 2];
            polar_from_words<true>(w, sr, cs, sn, sh, P.scale);
#pragma unroll
            for (int q = 0; q < kNpb - 1; q++) {
                const int j = n_whole * kNpb + q;
                if (j < P.n_dates)
                    step(P, dates[j], sr[q / 2], (q & 1) ? sn[q / 2] : cs[q / 2], y, cva, sh);
            }
        }
        v[0] = cva;
    }
};

template <class W>
static typename W::Params narrow(const CvaJob &job)
{
    using Real = typename W::Real;
    typename W::Params p;
    p.keys = job.keys;
    p.y0 = (Real)job.y0;
    p.mu_dt = (Real)job.mu_dt;
    p.scale = polar_scale<Real>(job.sig_dt);
    p.k = (Real)job.k;
    p.k_pdf = (Real)(job.k * 0.39894228040143267793994605993438);
    p.n_dates = job.n_dates;
    p.first_date = 0;
    return p;
}

// the table image of `n_jobs` jobs laid end to end, uploaded unless the device already holds exactly this image
template <typename Real>
static cudaError_t upload_dates(TableUse &use, const std::vector<CvaDate<Real>> &staging, cudaStream_t stream)
{
    if (use.status() != cudaSuccess)
        return use.status();
    if (use.needs_upload()) {
        cudaError_t e = cudaMemcpyToSymbolAsync(c_cva_table, staging.data(), staging.size() * sizeof(CvaDate<Real>), 0,
                                                cudaMemcpyHostToDevice, stream);
        if (e != cudaSuccess) {
            use.invalidate();
            return e;
        }
        use.uploaded();
    }
    return cudaSuccess;
}
template <typename Real>
static bool stage_dates(const CvaJob &job, std::vector<CvaDate<Real>> &staging)
{
    if (job.n_dates < 0 || staging.size() + (size_t)job.n_dates > (size_t)kCvaMaxDates)
        return false;
    for (int j = 0; j < job.n_dates; j++) {
        const CvaDateHost &h = job.dates[j];
        staging.push_back(CvaDate<Real>{(Real)h.w, (Real)h.inv, (Real)h.c1, (Real)h.sig, (Real)h.kd, (Real)0});
    }
    return true;
}

template <typename Real>
static cudaError_t launch_t(const CvaJob &job, const Geometry *geom, int grid, unsigned long long *d_acc,
                            unsigned long long first_unit, unsigned long long n_units, void *d_out,
                            cudaStream_t stream, const LaunchOptions &opt)
{
    std::vector<CvaDate<Real>> staging;
    if (!stage_dates(job, staging))
        return cudaErrorInvalidValue;
    if (staging.empty())
        staging.push_back(CvaDate<Real>{});
    TableUse use(g_cva_lock, stream, staging.data(), staging.size() * sizeof(CvaDate<Real>));
    cudaError_t e = upload_dates(use, staging, stream);
    if (e != cudaSuccess)
        return e;
    if (geom)
        e = accumulate_launch<Cva<Real, true>>(grid, narrow<Cva<Real, true>>(job), *geom, d_acc, stream, opt);
    else {
        const unsigned long long blocks = (n_units + kThreads - 1) / kThreads;
        mc_paths_kernel<Cva<Real>><<<(int)(blocks < 65535ull ? blocks : 65535ull), kThreads, 0, stream>>>(
            narrow<Cva<Real>>(job), first_unit, n_units, (Real *)d_out);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess)
        use.invalidate();
    return e;
}

// many CVA jobs (time grids, options, intensities ...) in one launch: their date tables laid end to end
template <typename Real>
static cudaError_t batch_t(const BatchShape &shape, const CvaJob *jobs, int grid, const BatchTarget &target, cudaStream_t stream,
                           const LaunchOptions &opt)
{
    using W = Cva<Real, true>;
    BatchJobs<W> b{};
    fill_batch_header(b, shape, target);
    std::vector<CvaDate<Real>> staging;
    for (int i = 0; i < shape.n_jobs; i++) {
        b.params[i] = narrow<W>(jobs[i]);
        b.params[i].first_date = (int)staging.size();
        if (!stage_dates(jobs[i], staging))
            return cudaErrorInvalidValue;
    }
    if (staging.empty())
        staging.push_back(CvaDate<Real>{});
    TableUse use(g_cva_lock, stream, staging.data(), staging.size() * sizeof(CvaDate<Real>));
    cudaError_t e = upload_dates(use, staging, stream);
    if (e != cudaSuccess)
        return e;
    e = accumulate_batch_launch<W>(grid, b, stream, opt);
    if (e != cudaSuccess)
        use.invalidate();
    return e;
}

int cva_batch_blocks_per_sm(int precision)
{
    return precision ? accumulate_batch_blocks_per_sm<Cva<double, true>>() : accumulate_batch_blocks_per_sm<Cva<float, true>>();
}

cudaError_t cva_batch_launch(int precision, const BatchShape &shape, const CvaJob *jobs, int grid, const BatchTarget &target,
                             cudaStream_t stream, const LaunchOptions &opt)
{
    if (shape.n_jobs < 1 || shape.n_jobs > kBatchMaxJobs)
        return cudaErrorInvalidValue;
    return precision ? batch_t<double>(shape, jobs, grid, target, stream, opt) : batch_t<float>(shape, jobs, grid, target, stream, opt);
}

int cva_blocks_per_sm(int precision)
{
    return precision ? accumulate_blocks_per_sm<Cva<double, true>>() : accumulate_blocks_per_sm<Cva<float, true>>();
}

cudaError_t cva_launch(int precision, const CvaJob &job, const Geometry &geom, int grid,
                       unsigned long long *d_acc, cudaStream_t stream, const LaunchOptions &opt)
{
    return precision ? launch_t<double>(job, &geom, grid, d_acc, 0, 0, nullptr, stream, opt)
                     : launch_t<float>(job, &geom, grid, d_acc, 0, 0, nullptr, stream, opt);
}

cudaError_t cva_paths(int precision, const CvaJob &job, unsigned long long first_unit,
                      unsigned long long n_units, void *d_out, cudaStream_t stream)
{
    return precision ? launch_t<double>(job, nullptr, 0, nullptr, first_unit, n_units, d_out, stream, LaunchOptions())
                     : launch_t<float>(job, nullptr, 0, nullptr, first_unit, n_units, d_out, stream, LaunchOptions());
}

}  // namespace mcb
