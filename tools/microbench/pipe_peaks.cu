// pipe_peaks.cu -- measured per-SM issue rates of the pipes the Monte Carlo kernels live on
// (FMA, FP64, XU/MUFU, ALU, integer multiply), the compute-side analogue of MEASURED_PEAKS.json.
// the roofline denominators of a compute-bound kernel must be measured, not
// taken from a data sheet.  Prints one JSON object; bench evidence only, not part of the product.
//
// Method: every thread runs kChains independent dependency chains of one instruction, kIters times
// (fully unrolled inner body of 32), 2 CTAs x 1024 threads per SM (16 resident warps per
// sub-partition); the rate is thread-instructions / (SM cycles x SMs), with SM cycles read from
// clock64() inside the kernel (immune to the host's view of the clock) and wall time from CUDA
// events for the absolute ops/s.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <vector>

#define CHECK(x)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (x);                                                                      \
        if (e_ != cudaSuccess) {                                                                   \
            std::fprintf(stderr, "%s failed: %s\n", #x, cudaGetErrorString(e_));                  \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

constexpr int kChains = 8;
constexpr int kInner = 32;

struct Fma32 {
    using T = float;
    static __device__ __forceinline__ T init(int i) { return 1.0f + i * 1e-3f; }
    static __device__ __forceinline__ T op(T x, T a, T b) { return fmaf(x, a, b); }
};
struct Fma64 {
    using T = double;
    static __device__ __forceinline__ T init(int i) { return 1.0 + i * 1e-3; }
    static __device__ __forceinline__ T op(T x, T a, T b) { return fma(x, a, b); }
};
struct Add64 {
    using T = double;
    static __device__ __forceinline__ T init(int i) { return 1.0 + i * 1e-3; }
    static __device__ __forceinline__ T op(T x, T a, T) { return x + a; }
};
#define MUFU_OP(NAME, PTX)                                                                        \
    struct NAME {                                                                                  \
        using T = float;                                                                           \
        static __device__ __forceinline__ T init(int i) { return 0.5f + i * 1e-2f; }             \
        static __device__ __forceinline__ T op(T x, T, T)                                          \
        {                                                                                          \
            float y;                                                                               \
            asm volatile(PTX " %0, %1;" : "=f"(y) : "f"(x));                                       \
            return y;                                                                              \
        }                                                                                          \
    };
MUFU_OP(MufuEx2, "ex2.approx.ftz.f32")
MUFU_OP(MufuLg2, "lg2.approx.ftz.f32")
MUFU_OP(MufuSin, "sin.approx.ftz.f32")
MUFU_OP(MufuCos, "cos.approx.ftz.f32")
MUFU_OP(MufuSqrt, "sqrt.approx.ftz.f32")
MUFU_OP(MufuRsqrt, "rsqrt.approx.ftz.f32")
struct MufuRcp {  // rcp(rcp(x)) folds to x in ptxas: keep an FADD between them (other pipe, not binding)
    using T = float;
    static __device__ __forceinline__ T init(int i) { return 0.5f + i * 1e-2f; }
    static __device__ __forceinline__ T op(T x, T a, T)
    {
        float y;
        asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
        return y + a;
    }
};
struct Rcp64h {  // MUFU.RCP64H: the fp64 reciprocal seed
    using T = double;
    static __device__ __forceinline__ T init(int i) { return 1.5 + i * 1e-2; }
    static __device__ __forceinline__ T op(T x, T, T)
    {
        double y;
        asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
        return y;
    }
};
struct I2F {
    using T = uint32_t;
    static __device__ __forceinline__ T init(int i) { return 12345u + i; }
    static __device__ __forceinline__ T op(T x, T, T) { return __float_as_uint(__uint2float_rn(x)); }
};
struct F2I {
    using T = uint32_t;
    static __device__ __forceinline__ T init(int i) { return 0x3f800000u + i; }
    static __device__ __forceinline__ T op(T x, T, T) { return (uint32_t)__float2int_rn(__uint_as_float(x)) + 0x3f800000u; }
};
struct ImadWide {
    using T = uint64_t;
    static __device__ __forceinline__ T init(int i) { return 0x9E3779B97F4A7C15ull + i; }
    static __device__ __forceinline__ T op(T x, T, T) { return (uint64_t)(uint32_t)x * 0xD2511F53u + (x >> 32); }
};
struct Imad {
    using T = uint32_t;
    static __device__ __forceinline__ T init(int i) { return 12345u + i; }
    static __device__ __forceinline__ T op(T x, T a, T b) { return x * a + b; }
};
struct Lop3 {
    using T = uint32_t;
    static __device__ __forceinline__ T init(int i) { return 12345u + i; }
    static __device__ __forceinline__ T op(T x, T a, T b)
    {
        uint32_t y;
        asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(y) : "r"(x), "r"(a), "r"(b));
        return y;
    }
};

template <class Op>
__global__ void __launch_bounds__(1024, 2) chain_kernel(int iters, typename Op::T a, typename Op::T b,
                                                     typename Op::T *out, long long *cycles)
{
    typename Op::T x[kChains];
#pragma unroll
    for (int c = 0; c < kChains; c++)
        x[c] = Op::init(c + threadIdx.x);
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < kInner; u++) {
#pragma unroll
            for (int c = 0; c < kChains; c++)
                x[c] = Op::op(x[c], a, b);
        }
    }
    const long long t1 = clock64();
    typename Op::T acc = x[0];
#pragma unroll
    for (int c = 1; c < kChains; c++)
        acc = acc + x[c];
    if (acc == (typename Op::T)123456789)
        out[0] = acc;  // keep the chains alive
    if (threadIdx.x == 0)
        cycles[blockIdx.x] = t1 - t0;
}

template <class Op>
int measure(const char *name, int sms, typename Op::T a, typename Op::T b, bool last)
{
    const int blocks = sms * 2, threads = 1024, iters = 256;
    typename Op::T *d_out;
    long long *d_cycles;
    CHECK(cudaMalloc(&d_out, sizeof(typename Op::T)));
    CHECK(cudaMalloc(&d_cycles, sizeof(long long) * blocks));
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0));
    CHECK(cudaEventCreate(&e1));
    for (int w = 0; w < 2; w++)
        chain_kernel<Op><<<blocks, threads>>>(iters, a, b, d_out, d_cycles);
    CHECK(cudaDeviceSynchronize());
    float best_ms = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        CHECK(cudaEventRecord(e0));
        chain_kernel<Op><<<blocks, threads>>>(iters, a, b, d_out, d_cycles);
        CHECK(cudaEventRecord(e1));
        CHECK(cudaEventSynchronize(e1));
        float ms;
        CHECK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best_ms)
            best_ms = ms;
    }
    std::vector<long long> cyc(blocks);
    CHECK(cudaMemcpy(cyc.data(), d_cycles, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
    double mean_cycles = 0;
    for (long long c : cyc)
        mean_cycles += (double)c;
    mean_cycles /= blocks;
    // 2 CTAs of 1024 threads share an SM for the whole kernel: per-SM thread-instructions / cycles
    const double per_thread = (double)iters * kInner * kChains;
    const double per_sm = per_thread * threads * 2;
    const double total = per_thread * threads * blocks;
    // clock64() does not tick at the SM clock on this platform (it reads ~1461 MHz while FFMA
    // retires 124/clk/SM at clocks.max.sm), so the per-clock figure is also given at clocks.max.sm
    cudaDeviceProp prop;
    CHECK(cudaGetDeviceProperties(&prop, 0));
    const double gops = total / (best_ms * 1e6);
    std::printf("  \"%s\": {\"gops\": %.1f, \"per_clk_per_sm_at_max_clock\": %.2f, \"per_clock64_tick_per_sm\": %.2f, \"ms\": %.4f, "
                "\"clock64_mhz\": %.0f}%s\n", name, gops, gops * 1e9 / ((double)sms * prop.clockRate * 1e3),
                per_sm / mean_cycles, best_ms, mean_cycles / (best_ms * 1e3), last ? "" : ",");
    cudaFree(d_out);
    cudaFree(d_cycles);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return 0;
}

int main()
{
    cudaDeviceProp prop;
    CHECK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    std::printf("{\n  \"device\": \"%s\", \"sms\": %d, \"clock_khz_reported\": %d,\n", prop.name, sms, prop.clockRate);
    int rc = 0;
    rc |= measure<Fma32>("ffma", sms, 0.999f, 1e-3f, false);
    rc |= measure<Fma64>("dfma", sms, 0.999, 1e-3, false);
    rc |= measure<Add64>("dadd", sms, 1e-3, 0.0, false);
    rc |= measure<MufuEx2>("mufu_ex2", sms, 0.f, 0.f, false);
    rc |= measure<MufuLg2>("mufu_lg2", sms, 0.f, 0.f, false);
    rc |= measure<MufuSin>("mufu_sin", sms, 0.f, 0.f, false);
    rc |= measure<MufuCos>("mufu_cos", sms, 0.f, 0.f, false);
    rc |= measure<MufuSqrt>("mufu_sqrt", sms, 0.f, 0.f, false);
    rc |= measure<MufuRsqrt>("mufu_rsqrt", sms, 0.f, 0.f, false);
    rc |= measure<MufuRcp>("mufu_rcp_plus_fadd", sms, 0.25f, 0.f, false);
    rc |= measure<Rcp64h>("mufu_rcp64h", sms, 0.0, 0.0, false);
    rc |= measure<I2F>("i2f_u32", sms, 0u, 0u, false);
    rc |= measure<F2I>("f2i_s32", sms, 0u, 0u, false);
    rc |= measure<ImadWide>("imad_wide_u32", sms, 0ull, 0ull, false);
    rc |= measure<Imad>("imad", sms, 0x9E3779B9u, 0xBB67AE85u, false);
    rc |= measure<Lop3>("lop3", sms, 0x9E3779B9u, 0xBB67AE85u, true);
    std::printf("}\n");
    return rc;
}
