// tune_vanilla.cu -- times launch-bound / unroll variants of the vanilla kernels on the device
// (bench evidence: profiles/r01_tune_vanilla.txt).  Standalone; not part of libmcb200.
#include <cstdio>
#include <cmath>
#include "workload_vanilla.cuh"

using namespace mcb;

template <class W>
float time_variant(const typename W::Params &p, unsigned long long paths, int sms, int *blocks_per_sm)
{
    Geometry g{};
    const unsigned long long units = paths / W::kUnitPaths;
    g.total_paths = paths;
    g.rounds = 64;
    g.chunk_units = 256ull * g.rounds;
    g.first_chunk = 0;
    g.n_chunks = units / g.chunk_units;
    g.scale_exp_sum = 73;
    g.scale_exp_sumsq = 66;
    *blocks_per_sm = accumulate_blocks_per_sm<W>();
    const int grid = sms * *blocks_per_sm;
    unsigned long long *acc;
    cudaMalloc(&acc, 96);
    cudaMemset(acc, 0, 96);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        accumulate_launch<W>(grid, p, g, acc, 0);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best)
            best = ms;
    }
    cudaFree(acc);
    return best;
}

template <typename Real, int B, int U>
void run(const char *name, unsigned long long paths, int sms)
{
    using W = Vanilla<Real, B, U, true>;
    typename W::Params p{};
    unsigned k0 = 0x30300001u, k1 = 0x6d636232u;
    for (int i = 0; i < 10; i++) {
        p.keys.k0[i] = k0;
        p.keys.k1[i] = k1;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    const double unit = sizeof(Real) == 4 ? 1.4426950408889634 : 1.0;
    p.a = (Real)((std::log(100.0) + (0.05 - 0.02) * 1.0) * unit);
    p.scale = polar_scale<Real>(0.2 * unit);
    p.k = (Real)100.0;
    int bps = 0;
    const float ms = time_variant<W>(p, paths, sms, &bps);
    cudaFuncAttributes attr;
    cudaFuncGetAttributes(&attr, mc_accumulate_kernel<W>);
    std::printf("%s minblocks=%d unroll=%d regs=%d blocks/SM=%d  %.3f ms  %.3e paths/s\n", name, B, U, attr.numRegs, bps, ms,
                paths / (ms * 1e-3));
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    const unsigned long long paths = 1ull << 32;
    run<float, 3, 1>("f32", paths, sms);
    run<float, 4, 1>("f32", paths, sms);
    run<float, 5, 1>("f32", paths, sms);
    run<float, 6, 1>("f32", paths, sms);
    run<float, 8, 1>("f32", paths, sms);
    run<float, 4, 2>("f32", paths, sms);
    run<float, 3, 2>("f32", paths, sms);
    run<float, 2, 2>("f32", paths, sms);
    run<float, 2, 4>("f32", paths, sms);
    run<double, 1, 1>("f64", paths, sms);
    run<double, 2, 1>("f64", paths, sms);
    run<double, 3, 1>("f64", paths, sms);
    run<double, 4, 1>("f64", paths, sms);
    run<double, 2, 2>("f64", paths, sms);
    run<double, 3, 2>("f64", paths, sms);
    return 0;
}
