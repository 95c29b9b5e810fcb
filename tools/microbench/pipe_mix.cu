// pipe_mix.cu -- co-issue ceilings of the instruction classes the pricing kernels are made of:
// independent single-instruction dependency chains (inline PTX, one SASS instruction per step,
// verified with cuobjdump) interleaved pairwise.  Answers "which pipes overlap on an sm_100a
// sub-partition?" -- the question that decides the instruction-mix budget in DESIGN.md.
// Standalone bench evidence, not part of libmcb200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum Cls { DFMA, FFMA, LOP3, IMADW, IMAD, MUFU, NCLS };
static const char *kNames[NCLS] = {"dfma", "ffma", "lop3", "imad.wide", "imad", "mufu"};

template <int C> struct Chain;
template <> struct Chain<DFMA> {
    double x;
    __device__ void init(int i) { x = 1.0 + 1e-3 * i; }
    __device__ void step() { asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(0.999), "d"(1e-3)); }
    __device__ double value() { return x; }
};
template <> struct Chain<FFMA> {
    float x;
    __device__ void init(int i) { x = 1.0f + 1e-3f * i; }
    __device__ void step() { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(0.999f), "f"(1e-3f)); }
    __device__ double value() { return x; }
};
template <> struct Chain<LOP3> {
    uint32_t x, a, b;
    __device__ void init(int i) { x = 12345u + i; a = 0x9E3779B9u + i; b = 0xBB67AE85u ^ i; }
    __device__ void step() { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(a), "r"(b)); }
    __device__ double value() { return x; }
};
template <> struct Chain<IMADW> {
    uint64_t x;
    uint32_t m;
    __device__ void init(int i) { x = 0x9E3779B97F4A7C15ull + i; m = 0xD2511F53u + 2 * i; }
    __device__ void step()
    {
        uint32_t lo = (uint32_t)x;
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x) : "r"(lo), "r"(m));
    }
    __device__ double value() { return (double)x; }
};
template <> struct Chain<IMAD> {
    uint32_t x, m, c;
    __device__ void init(int i) { x = 12345u + i; m = 0xD2511F53u + 2 * i; c = 17u + i; }
    __device__ void step() { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(m), "r"(c)); }
    __device__ double value() { return x; }
};
template <> struct Chain<MUFU> {
    float x;
    __device__ void init(int i) { x = 0.5f + 1e-3f * i; }
    __device__ void step() { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x)); }
    __device__ double value() { return x; }
};

// NA chains of class A and NB chains of class B per thread, interleaved
template <int A, int NA, int B, int NB>
__global__ void __launch_bounds__(256, 4) pair_kernel(int iters, double *out)
{
    Chain<A> a[NA];
    Chain<B> b[NB > 0 ? NB : 1];
    for (int i = 0; i < NA; i++) a[i].init(threadIdx.x + i);
    for (int i = 0; i < NB; i++) b[i].init(threadIdx.x + 7 * i);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < (NA > NB ? NA : NB); i++) {
                if (i < NA) a[i].step();
                if (i < NB) b[i].step();
            }
        }
    }
    double acc = 0;
    for (int i = 0; i < NA; i++) acc += a[i].value();
    for (int i = 0; i < NB; i++) acc += b[i].value();
    if (acc == 123.456) out[0] = acc;
}

template <int A, int NA, int B, int NB>
void run(int sms)
{
    double *out;
    cudaMalloc(&out, 8);
    const int iters = 2048, blocks = sms * 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        pair_kernel<A, NA, B, NB><<<blocks, 256>>>(iters, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    // cycles an SM sub-partition spends per group of (NA + NB) warp-instructions: 8 warps share it
    const double cycles = best * 1e-3 * 1.965e9;
    const double groups = (double)iters * 8 * 8;  // per warp x 8 warps per sub-partition
    if (NB > 0)
        std::printf("%-9s x%d + %-9s x%d : %7.3f ms  %6.2f cycles per group  (%.2f warp-instr/clk per sub-partition)\n", kNames[A], NA,
                    kNames[B], NB, best, cycles / groups, (NA + NB) * groups / cycles);
    else
        std::printf("%-9s x%d alone          : %7.3f ms  %6.2f cycles per group  (%.2f warp-instr/clk per sub-partition)\n", kNames[A], NA, best,
                    cycles / groups, NA * groups / cycles);
    cudaFree(out);
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    std::printf("# 4 CTAs x 256 threads per SM (8 warps per sub-partition), cycles at 1965 MHz\n");
    run<DFMA, 4, DFMA, 0>(sms);
    run<FFMA, 4, DFMA, 0>(sms);
    run<LOP3, 4, DFMA, 0>(sms);
    run<IMADW, 4, DFMA, 0>(sms);
    run<IMAD, 4, DFMA, 0>(sms);
    run<MUFU, 4, DFMA, 0>(sms);
    run<DFMA, 4, LOP3, 4>(sms);
    run<DFMA, 4, IMADW, 4>(sms);
    run<DFMA, 4, IMAD, 4>(sms);
    run<DFMA, 4, FFMA, 4>(sms);
    run<DFMA, 4, MUFU, 1>(sms);
    run<FFMA, 4, LOP3, 4>(sms);
    run<FFMA, 4, IMADW, 4>(sms);
    run<FFMA, 4, IMAD, 4>(sms);
    run<FFMA, 4, MUFU, 1>(sms);
    run<LOP3, 4, IMADW, 4>(sms);
    run<LOP3, 4, IMAD, 4>(sms);
    run<LOP3, 4, MUFU, 1>(sms);
    run<IMADW, 4, MUFU, 1>(sms);
    return 0;
}
