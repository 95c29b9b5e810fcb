// kernels_vanilla.cu -- European call: fused Philox -> Box-Muller -> GBM terminal value -> payoff
// -> (sum, sum^2), fp32 and fp64 (sm_100a).
//
// Replaces callPayoff + vanillaOptMonteCarlo (DP/MonteCarloKernel.cu:67-71, :179-220).
// One draw unit = one Philox block = 6 (fp32) or 4 (fp64) consecutive paths:
//   fp32: payoff = max(2^(a + b z) - K, 0),  a = log2(S0) + (r - v^2/2) T log2(e),  b = v sqrt(T) log2(e)
//   fp64: payoff = max(e^(a + b z) - K, 0),  a = ln(S0) + (r - v^2/2) T,            b = v sqrt(T)
// so a path costs one FMA and one exponential after its normal.
#include "launch.h"
#include "workload_vanilla.cuh"

namespace mcb {

template <typename Real>
static typename Vanilla<Real>::Params narrow(const VanillaJob &job)
{
    typename Vanilla<Real>::Params p;
    p.keys = job.keys;
    p.a = (Real)job.a;
    p.k = (Real)job.k;
    p.scale = polar_scale<Real>(job.b);
    return p;
}

int vanilla_blocks_per_sm(int precision)
{
    int n = 0;
    cudaError_t e = precision ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                                    &n, mc_accumulate_kernel<Vanilla<double>>, kThreads, 0)
                              : cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                                    &n, mc_accumulate_kernel<Vanilla<float>>, kThreads, 0);
    return e == cudaSuccess ? n : 0;
}

cudaError_t vanilla_launch(int precision, const VanillaJob &job, const Geometry &geom, int grid,
                           unsigned long long *d_acc, cudaStream_t stream)
{
    if (precision)
        mc_accumulate_kernel<Vanilla<double>><<<grid, kThreads, 0, stream>>>(narrow<double>(job), geom, d_acc);
    else
        mc_accumulate_kernel<Vanilla<float>><<<grid, kThreads, 0, stream>>>(narrow<float>(job), geom, d_acc);
    return cudaGetLastError();
}

cudaError_t vanilla_paths(int precision, const VanillaJob &job, unsigned long long first_unit,
                          unsigned long long n_units, void *d_out, cudaStream_t stream)
{
    const int grid = (int)((n_units + kThreads - 1) / kThreads < 65535ull ? (n_units + kThreads - 1) / kThreads
                                                                           : 65535ull);
    if (precision)
        mc_paths_kernel<Vanilla<double>><<<grid, kThreads, 0, stream>>>(narrow<double>(job), first_unit, n_units,
                                                                        (double *)d_out);
    else
        mc_paths_kernel<Vanilla<float>><<<grid, kThreads, 0, stream>>>(narrow<float>(job), first_unit, n_units,
                                                                       (float *)d_out);
    return cudaGetLastError();
}

}  // namespace mcb
