// kernels_cva.cu -- CVA of one European call under Black-Scholes, fp32 and fp64 (sm_100a).
//
// Replaces geomBrownian + cnd + device_bsCall + cvaCallOptMC (DP/MonteCarloKernel.cu:104-129,
// :222-283).  One draw unit = one path; exposure date j uses normal j of the path's sub-stream.
// The state is the log-moneyness y = ln(S/K), and everything is priced per unit of strike.  Per kept date j
// (host-built table, engine.cu):
//   y  += mu_dt + sig_dt z                      exact GBM step
//   s   = e^y                                   = S / K
//   d1  = y inv_j + c1_j,  d2 = d1 - sig_j      inv_j = 1/(v sqrt(tau_j)), c1_j = (r + v^2/2) tau_j inv_j
//   A   = e^{y - d1^2/2}                         one exponential: sqrt(2 pi) s phi(d1), and sqrt(2 pi) kd_j phi(d2) as
//                                               well (kd_j = e^{-r tau_j})
//   q_i = cnd-tail(|d_i|) / (sqrt(2 pi) phi(d_i))   Hastings polynomial in 1 / (1 + 0.2316419 |d_i|), like the
//                                               reference, with 1/sqrt(2 pi) folded into its coefficients
//   ee  = s cnd(d1) - kd_j cnd(d2)
//       = s [d1 > 0] - kd_j [d2 > 0] - A (sgn(d1) q1 - sgn(d2) q2)
//   cva += w_j ee                               w_j = K LGD (e^{-lambda t_{j-1}} - e^{-lambda t_j})
// so a path-step costs one normal, two exponentials and two reciprocals; the reference spends six
// exponentials, a log, three square roots and four divisions on it.  y is carried in the units the exponential
// is cheapest in (exp_scaled: log2 units for fp32, units of ln2/256 for fp64; the table's inv_j and the -1/2 of the
// density exponent absorb the factor), so neither exponential has an argument reduction to pay for and neither
// result is multiplied by anything.  Dates the reference drops (remaining time rounded below zero)
// are simply absent from the table.
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

namespace mcb {

template <typename Real>
struct alignas(16) CvaDate {  // three 16-byte constant loads per date
    Real w, inv, c1, sig, kd, pad;
};

__constant__ __align__(16) unsigned char c_cva_table[kCvaMaxDates * sizeof(CvaDate<double>)];
static TableLock g_cva_lock;

// q(d) = cnd-tail(|d|) / (sqrt(2 pi) phi(d)) = k P(k), k = 1 / (1 + 0.2316419 |d|)
// (Abramowitz-Stegun 26.2.17, the constants of DP/MonteCarloKernel.cu:111-116 times 1/sqrt(2 pi)).  Returned as the
// two factors: the caller needs sgn(d1) q1 - sgn(d2) q2, which is one multiply and one FMA on (k, P) pairs with the
// signs put on the k's -- a multiply, a multiply and a subtraction on finished q's.
template <typename Real>
__device__ __forceinline__ void hastings_factors(Real d, Real &k, Real &poly)
{
    constexpr double c = 0.39894228040143267793994605993438;
    k = rcp_real(fma((Real)0.2316419, fabs(d), (Real)1.0));
    poly = fma(k, (Real)(1.330274429 * c), (Real)(-1.821255978 * c));
    poly = fma(k, poly, (Real)(1.781477937 * c));
    poly = fma(k, poly, (Real)(-0.356563782 * c));
    poly = fma(k, poly, (Real)(0.31938153 * c));
}

// sgn(d) k for k > 0 and x [d > 0], by the sign bit of d on the integer pipe (the fp64 comparisons are
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
// (fp64: only the HIGH word is cleared -- what is left is below 2^-1042, i.e. nothing next to any term it meets here,
// and costs two integer instructions instead of three)
__device__ __forceinline__ double keep_if_positive(double x, double d)
{
    return __hiloint2double(__double2hiint(x) & ~(__double2hiint(d) >> 31), __double2loint(x));
}
__device__ __forceinline__ float keep_if_positive(float x, float d)
{
    return __int_as_float(__float_as_int(x) & ~(__float_as_int(d) >> 31));
}

// max(a, -258048) -- e^-698.7 in units of ln2/256 -- without touching the fp64 pipe: negative doubles order like
// their high words taken as unsigned integers and every non-negative one lies below them, so one integer min on the
// high word does it (-inf and NaN included).  The low word stays as it is: a floored value lies within a quarter of
// -258048, still inside the exponential's range and still e^x = 0 for every purpose.
__device__ __forceinline__ double floor_exponent(double a)
{
    constexpr unsigned kFloorHi = 0xC10F8000u;  // high word of -258048.0
    return __hiloint2double((int)min((unsigned)__double2hiint(a), kFloorHi), __double2loint(a));
}
__device__ __forceinline__ float floor_exponent(float a) { return a; }  // MUFU.EX2(-inf) = 0

// kGlobalDates: the date table does not fit the constant buffer (more than kCvaMaxDates kept dates) and is read from
// device memory through Params::dates_global instead -- the same uniform loads, through L1 instead of the constant cache
template <typename RealT, bool kAccumLayout = false, bool kGlobalDates = false>
struct Cva {
    using Real = RealT;
    static constexpr int kUnitPaths = 1;
    static constexpr int kUnroll = 1;
    static constexpr int kSubBlocks = (kAccumLayout && sizeof(RealT) == 8) ? MCB_CVA_SUBBLOCKS : 1;  // fp64: one table set per SM
    static constexpr int kMinBlocks = kSubBlocks > 1 ? 1 : 3;
    static constexpr int kNpb = NormalsPerBlock<RealT>::value;
    struct Params {
        PhiloxKeys keys;
        Real y0, mu_dt;          // ln(S0/K) and the drift of a step, in the units of exp_scaled
        Real half_unit;          // -1/2 of that unit: the density exponent y - d1^2/2
        PolarScale<Real> scale;  // of sig_dt = v sqrt(dt) (same units), folded under the Box-Muller square root
        int n_dates;  // kept dates
        int first_date;  // of this job in the device table (0 unless the launch carries several jobs)
        const void *dates_global;  // kGlobalDates: CvaDate<Real>[n_dates] in device memory
    };
    using Shared = std::conditional_t<kAccumLayout, typename SharedAccumFor<Real>::type, typename SharedFor<Real>::type>;
    using JobState = typename JobStateFor<Real>::type;
    static constexpr bool kClampAtZero = false;
    static __device__ __forceinline__ void prepare(const Params &P, JobState &job, int tid) { prepare_polar(P.scale, job, tid); }
    // one exposure date; the diffusion sig_dt z arrives as (sig_dt r) * (cos or sin) and folds into the step's FMA
    static __device__ __forceinline__ void step(const Params &P, const CvaDate<Real> &D, Real sr, Real trig, Real &y,
                                                Real &cva, const Shared &sh)
    {
        y = fma(sr, trig, y + P.mu_dt);
        const Real s = exp_scaled<true>(y, sh);
        const Real d1 = fma(y, D.inv, D.c1);
        const Real d2 = d1 - D.sig;
        // sqrt(2 pi) s phi(d1) = sqrt(2 pi) kd phi(d2) = e^{y - d1^2/2}.  The exponent can be -1e14 (a date a few ulps
        // before maturity) or -inf (exact grid, tau = 0): floor it where e^x is already 0 for every purpose, so the
        // table-driven exp stays in range
        const Real a = exp_scaled<true>(floor_exponent(fma(P.half_unit * d1, d1, y)), sh);
        Real k1, p1, k2, p2;
        hastings_factors(d1, k1, p1);
        hastings_factors(d2, k2, p2);
        const Real tails = fma(with_sign_of(k1, d1), p1, -(with_sign_of(k2, d2) * p2));
        const Real ee = fma(-a, tails, keep_if_positive(s, d1) - keep_if_positive(D.kd, d2));
        cva = fma(D.w, ee, cva);
    }
    static __device__ __forceinline__ void eval(const Params &P, uint32_t path_lo, uint32_t path_hi, Real (&v)[1],
                                                const Shared &sh, const JobState &job)
    {
        const CvaDate<Real> *dates = kGlobalDates ? static_cast<const CvaDate<Real> *>(P.dates_global)
                                                  : reinterpret_cast<const CvaDate<Real> *>(c_cva_table) + P.first_date;
        Real y = P.y0, cva = 0;
        // whole draw blocks run their kNpb dates back to back with no test in between: only y links one date to the
        // next, so the exponentials and reciprocals of neighbouring dates overlap
        const int n_whole = P.n_dates / kNpb;
        // the Philox block of the NEXT four dates is drawn while this block's dates are priced: integer work for the
        // issue slots between the fp64 instructions (14.66 -> 14.50 ms, profiles/r02k_ab_experiments.txt; the same
        // prefetch in the European call's unit loop was 1.6 % SLOWER and is not used there).  The last prefetch
        // serves the tail below.
        uint32_t w[4];
        philox4x32_10(path_lo, path_hi, 0u, kTagCva, P.keys, w);
#pragma unroll 1
        for (int jb = 0; jb < n_whole; jb++) {
            uint32_t wn[4];
            philox4x32_10(path_lo, path_hi, (uint32_t)(jb + 1), kTagCva, P.keys, wn);
            Real sr[kNpb / 2], cs[kNpb / 2], sn[kNpb / 2];
            polar_from_words<true>(w, sr, cs, sn, sh, P.scale, job);
#pragma unroll
            for (int q = 0; q < kNpb; q++)
                step(P, dates[jb * kNpb + q], sr[q / 2], (q & 1) ? sn[q / 2] : cs[q / 2], y, cva, sh);
#pragma unroll
            for (int i = 0; i < 4; i++)
                w[i] = wn[i];
        }
        if (n_whole * kNpb < P.n_dates) {
            Real sr[kNpb / 2], cs[kNpb / 2], sn[kNpb / 2];
            polar_from_words<true>(w, sr, cs, sn, sh, P.scale, job);
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
    constexpr double unit = ExpUnit<Real>::value;
    p.keys = job.keys;
    p.y0 = (Real)(job.y0 * unit);
    p.mu_dt = (Real)(job.mu_dt * unit);
    p.half_unit = (Real)(-0.5 * unit);
    p.scale = polar_scale<Real>(job.sig_dt * unit);
    p.n_dates = job.n_dates;
    p.first_date = 0;
    p.dates_global = nullptr;
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
static bool stage_dates(const CvaJob &job, std::vector<CvaDate<Real>> &staging, size_t capacity = (size_t)kCvaMaxDates)
{
    if (job.n_dates < 0 || staging.size() + (size_t)job.n_dates > capacity)
        return false;
    for (int j = 0; j < job.n_dates; j++) {
        const CvaDateHost &h = job.dates[j];
        // per unit of strike, slope per unit of the kernel's y (see the header of this file)
        staging.push_back(CvaDate<Real>{(Real)(h.w * job.k), (Real)(h.inv / ExpUnit<Real>::value), (Real)h.c1, (Real)h.sig,
                                        (Real)(h.kd / job.k), (Real)0});
    }
    return true;
}

// ---- long grids: more kept dates than the constant buffer holds ----
// One growable device buffer per device, shared like the constant table (same lock discipline, same "skip the upload
// when the device already holds this image").  The reference has no limit here (cva->n is a plain int,
// DP/MonteCarloKernel.cu:247); neither has this path, up to MCB200_MAX_DATES.
static TableLock g_cva_wide_lock;
static void *g_cva_wide_buffer[TableLock::kMaxDevices] = {};
static size_t g_cva_wide_capacity[TableLock::kMaxDevices] = {};

template <typename Real>
static cudaError_t launch_wide_t(const CvaJob &job, const Geometry *geom, int grid, unsigned long long *d_acc,
                                 unsigned long long first_unit, unsigned long long n_units, void *d_out,
                                 cudaStream_t stream, const LaunchOptions &opt)
{
    std::vector<CvaDate<Real>> staging;
    if (!stage_dates(job, staging, (size_t)job.n_dates))
        return cudaErrorInvalidValue;
    const size_t bytes = staging.size() * sizeof(CvaDate<Real>);
    TableUse use(g_cva_wide_lock, stream, staging.data(), bytes);
    if (use.status() != cudaSuccess)
        return use.status();
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess)
        return e;
    if (g_cva_wide_capacity[device] < bytes) {
        // (cudaFree waits for everything that may still read the old buffer)
        use.invalidate();
        if (g_cva_wide_buffer[device])
            cudaFree(g_cva_wide_buffer[device]);
        g_cva_wide_buffer[device] = nullptr;
        g_cva_wide_capacity[device] = 0;
        e = cudaMalloc(&g_cva_wide_buffer[device], bytes);
        if (e != cudaSuccess)
            return e;
        g_cva_wide_capacity[device] = bytes;
    }
    if (use.needs_upload() || g_cva_wide_buffer[device] == nullptr) {
        e = cudaMemcpyAsync(g_cva_wide_buffer[device], staging.data(), bytes, cudaMemcpyHostToDevice, stream);
        if (e != cudaSuccess) {
            use.invalidate();
            return e;
        }
        use.uploaded();
    }
    if (geom) {
        using W = Cva<Real, true, true>;
        typename W::Params p = narrow<W>(job);
        p.dates_global = g_cva_wide_buffer[device];
        e = accumulate_launch<W>(grid, p, *geom, d_acc, stream, opt);
    } else {
        using W = Cva<Real, false, true>;
        typename W::Params p = narrow<W>(job);
        p.dates_global = g_cva_wide_buffer[device];
        const unsigned long long blocks = (n_units + kThreads - 1) / kThreads;
        mc_paths_kernel<W><<<(int)(blocks < 65535ull ? blocks : 65535ull), kThreads, 0, stream>>>(p, first_unit, n_units, (Real *)d_out);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess)
        use.invalidate();
    return e;
}

template <typename Real>
static cudaError_t launch_t(const CvaJob &job, const Geometry *geom, int grid, unsigned long long *d_acc,
                            unsigned long long first_unit, unsigned long long n_units, void *d_out,
                            cudaStream_t stream, const LaunchOptions &opt)
{
    if (job.n_dates > kCvaMaxDates)
        return launch_wide_t<Real>(job, geom, grid, d_acc, first_unit, n_units, d_out, stream, opt);
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

int cva_blocks_per_sm(int precision, int n_dates)
{
    if (n_dates > kCvaMaxDates)
        return precision ? accumulate_blocks_per_sm<Cva<double, true, true>>() : accumulate_blocks_per_sm<Cva<float, true, true>>();
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
