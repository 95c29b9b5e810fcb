// kernels_vanilla.cu -- European call: fused Philox -> Box-Muller -> GBM terminal value -> payoff
// -> (sum, sum^2), fp32 and fp64 (sm_100a).
//
// Replaces callPayoff + vanillaOptMonteCarlo (DP/MonteCarloKernel.cu:67-71, :179-220).
// One draw unit = one Philox block = 6 (fp32) or 4 (fp64) consecutive paths:
//   payoff = max(e^(a + b z) - K, 0),  a = ln(S0) + (r - v^2/2) T,  b = v sqrt(T)
// with a and b scaled on the host into the units the exponential is cheapest in (fp32: log2 units for MUFU.EX2; fp64:
// units of ln2/256 for the table-driven exp_units), so a path costs one FMA and one exponential after its normal.
#include "launch.h"
#include "workload_vanilla.cuh"

namespace mcb {

template <typename Real> using VanillaAccum = Vanilla<Real, VanillaTuning<Real>::kMinBlocks, VanillaTuning<Real>::kUnroll, true>;

template <class W>
static typename W::Params narrow(const VanillaJob &job)
{
    using Real = typename W::Real;
    typename W::Params p;
    p.keys = job.keys;
    p.a = (Real)(job.a * ExpUnit<Real>::value);
    p.k = (Real)job.k;
    p.scale = polar_scale<Real>(job.b * ExpUnit<Real>::value);
    return p;
}

int vanilla_blocks_per_sm(int precision)
{
    return precision ? accumulate_blocks_per_sm<VanillaAccum<double>>() : accumulate_blocks_per_sm<VanillaAccum<float>>();
}

cudaError_t vanilla_launch(int precision, const VanillaJob &job, const Geometry &geom, int grid,
                           unsigned long long *d_acc, cudaStream_t stream, const LaunchOptions &opt)
{
    if (precision)
        return accumulate_launch<VanillaAccum<double>>(grid, narrow<VanillaAccum<double>>(job), geom, d_acc, stream, opt);
    return accumulate_launch<VanillaAccum<float>>(grid, narrow<VanillaAccum<float>>(job), geom, d_acc, stream, opt);
}

// many European calls in one launch (mc_accumulate_batch_kernel)
template <class W>
static cudaError_t batch_t(const BatchShape &shape, const VanillaJob *jobs, int grid, const BatchTarget &target, cudaStream_t stream,
                           const LaunchOptions &opt)
{
    BatchJobs<W> b{};
    fill_batch_header(b, shape, target);
    for (int i = 0; i < shape.n_jobs; i++)
        b.params[i] = narrow<W>(jobs[i]);
    return accumulate_batch_launch<W>(grid, b, stream, opt);
}

int vanilla_batch_blocks_per_sm(int precision)
{
    return precision ? accumulate_batch_blocks_per_sm<VanillaAccum<double>>() : accumulate_batch_blocks_per_sm<VanillaAccum<float>>();
}

cudaError_t vanilla_batch_launch(int precision, const BatchShape &shape, const VanillaJob *jobs, int grid,
                                 const BatchTarget &target, cudaStream_t stream, const LaunchOptions &opt)
{
    if (shape.n_jobs < 1 || shape.n_jobs > kBatchMaxJobs)
        return cudaErrorInvalidValue;
    return precision ? batch_t<VanillaAccum<double>>(shape, jobs, grid, target, stream, opt)
                     : batch_t<VanillaAccum<float>>(shape, jobs, grid, target, stream, opt);
}

cudaError_t vanilla_paths(int precision, const VanillaJob &job, unsigned long long first_unit,
                          unsigned long long n_units, void *d_out, cudaStream_t stream)
{
    const int grid = (int)((n_units + kThreads - 1) / kThreads < 65535ull ? (n_units + kThreads - 1) / kThreads
                                                                           : 65535ull);
    if (precision)
        mc_paths_kernel<Vanilla<double>><<<grid, kThreads, 0, stream>>>(narrow<Vanilla<double>>(job), first_unit, n_units,
                                                                        (double *)d_out);
    else
        mc_paths_kernel<Vanilla<float>><<<grid, kThreads, 0, stream>>>(narrow<Vanilla<float>>(job), first_unit, n_units,
                                                                       (float *)d_out);
    return cudaGetLastError();
}

}  // namespace mcb
