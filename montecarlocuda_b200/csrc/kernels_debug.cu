// kernels_debug.cu -- parity instrumentation: the generator, the normal transform and the chunk
// reduction exposed on their own, running the SAME device functions as the pricing kernels, so the
// tests can compare them bit for bit (integer stages) or within a stated tolerance (normals) with
// the CPU oracle.
#include "device_math.cuh"
#include "launch.h"

namespace mcb {

__global__ void debug_philox_kernel(unsigned long long n, const uint32_t *__restrict__ ctr,
                                    const __grid_constant__ PhiloxKeys keys, uint32_t *__restrict__ out)
{
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        uint32_t w[4];
        philox4x32_10(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], keys, w);
#pragma unroll
        for (int q = 0; q < 4; q++)
            out[4 * i + q] = w[q];
    }
}

template <typename Real, int kPerBlock>
__global__ void debug_normals_kernel(unsigned long long n, const uint32_t *__restrict__ ctr,
                                     const __grid_constant__ PhiloxKeys keys, Real *__restrict__ out)
{
    __shared__ typename SharedFor<Real>::type sh;
    __shared__ typename JobStateFor<Real>::type job;
    sh.load();
    prepare_polar(polar_scale<Real>(1.0), job, (int)threadIdx.x);
    __syncthreads();
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        uint32_t w[4];
        philox4x32_10(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], keys, w);
        Real z[kPerBlock];
        normals_from_words(w, z, sh, job);
#pragma unroll
        for (int q = 0; q < kPerBlock; q++)
            out[kPerBlock * i + q] = z[q];
    }
}

// One chunk of given per-path values through the pricing kernels' accumulation order and
// chunk_commit / scratch_flush.
__global__ void __launch_bounds__(kThreads)
debug_reduce_kernel(const double *__restrict__ values, unsigned long long n_valid, int unit_paths,
                    bool accumulate_in_float, const __grid_constant__ Geometry G,
                    unsigned long long *__restrict__ acc)
{
    __shared__ BlockScratch sc;
    scratch_init(sc);
    double s = 0, s2 = 0;
    // fp32 with an even number of paths per unit: even and odd paths of a unit in running sums of their own, joined by
    // one float addition at the end (PackedSums, device_common.cuh); otherwise fs[1] stays 0
    float fs[2] = {0, 0}, fs2[2] = {0, 0};
    const bool interleaved = (unit_paths & 1) == 0;
    for (int k = 0; k < G.rounds; k++) {
        const unsigned long long unit = (unsigned long long)k * kThreads + threadIdx.x;
        for (int q = 0; q < unit_paths; q++) {
            const unsigned long long idx = unit * (unsigned long long)unit_paths + q;
            if (idx >= n_valid)
                continue;
            if (accumulate_in_float) {
                const float x = (float)values[idx];
                const int which = interleaved ? (q & 1) : 0;
                fs[which] += x;
                fs2[which] = fmaf(x, x, fs2[which]);
            } else {
                const double x = values[idx];
                s += x;
                s2 = fma(x, x, s2);
            }
        }
    }
    if (accumulate_in_float) {
        s = (double)(fs[0] + fs[1]);
        s2 = (double)(fs2[0] + fs2[1]);
    }
    chunk_commit(s, s2, n_valid, G, sc);
    scratch_flush(sc, acc);
}

// The fp64 special functions on their own: fn 0 = -2 ln(u), 1 = sqrt, 2 = 1/x, 3 = e^x,
// 4 = cos/sin of 2 pi k / 2^52 (input reinterpreted as the 52-bit integer k; two outputs),
// 5 = cos/sin of 2 pi k / 2^20 from the two-level table (k = low word of the input), 6 = sqrt (short iteration),
// 7 = 2^(y/256) (exp_units), 8 = the same with the table entry entering last (exp_units<true>: basket, CVA).
__global__ void debug_math64_kernel(int fn, unsigned long long n, const double *__restrict__ in,
                                    double *__restrict__ out)
{
    __shared__ SharedTables64 sh;
    __shared__ LogScale64 job;
    sh.load();
    prepare_polar(polar_scale<double>(1.0), job, (int)threadIdx.x);
    __syncthreads();
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const double x = in[i];
        double a = 0, b = 0;
        switch (fn) {
            case 0: a = neg2log_unit(x, sh.t, job); break;
            case 7: a = exp_units(x, sh.t); break;
            case 8: a = exp_units<true>(x, sh.t); break;
            case 1: a = sqrt_pos(x); break;
            case 6: a = sqrt_pos<true>(x); break;
            case 2: a = rcp_newton(x); break;
            case 3: a = exp_tab(x, sh.t); break;
            case 4: sincos_turn((uint32_t)__double2hiint(x), (uint32_t)__double2loint(x), a, b); break;
            default: sincos_turn20((uint32_t)__double2loint(x), a, b, sh.t); break;
        }
        out[2 * i] = a;
        out[2 * i + 1] = b;
    }
}

static int grid_for(unsigned long long n)
{
    const unsigned long long blocks = (n + kThreads - 1) / kThreads;
    return (int)(blocks < 1 ? 1 : (blocks < 65535ull ? blocks : 65535ull));
}

cudaError_t debug_philox(unsigned long long n, const uint32_t *d_ctr, PhiloxKeys keys, uint32_t *d_out,
                         cudaStream_t stream)
{
    debug_philox_kernel<<<grid_for(n), kThreads, 0, stream>>>(n, d_ctr, keys, d_out);
    return cudaGetLastError();
}

cudaError_t debug_normals(int precision, unsigned long long n, const uint32_t *d_ctr, PhiloxKeys keys,
                          void *d_out, cudaStream_t stream)
{
    if (precision)
        debug_normals_kernel<double, NormalsPerBlock<double>::value><<<grid_for(n), kThreads, 0, stream>>>(n, d_ctr, keys, (double *)d_out);
    else
        debug_normals_kernel<float, NormalsPerBlock<float>::value><<<grid_for(n), kThreads, 0, stream>>>(n, d_ctr, keys, (float *)d_out);
    return cudaGetLastError();
}

cudaError_t debug_math64(int fn, unsigned long long n, const double *d_in, double *d_out, cudaStream_t stream)
{
    debug_math64_kernel<<<grid_for(n), kThreads, 0, stream>>>(fn, n, d_in, d_out);
    return cudaGetLastError();
}

cudaError_t debug_reduce(const double *d_values, unsigned long long n_valid, int unit_paths, int rounds,
                         bool accumulate_in_float, int scale_exp_sum, int scale_exp_sumsq,
                         unsigned long long *d_acc, cudaStream_t stream)
{
    Geometry g{};
    g.total_paths = n_valid;
    g.chunk_units = (unsigned long long)kThreads * rounds;
    g.first_chunk = 0;
    g.n_chunks = 1;
    g.rounds = rounds;
    g.scale_exp_sum = scale_exp_sum;
    g.scale_exp_sumsq = scale_exp_sumsq;
    debug_reduce_kernel<<<1, kThreads, 0, stream>>>(d_values, n_valid, unit_paths, accumulate_in_float, g, d_acc);
    return cudaGetLastError();
}

}  // namespace mcb
