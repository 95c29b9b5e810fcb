// pipe_ffma2.cu -- does the packed FP32 FMA of sm_100 (fma.rn.f32x2 -> FFMA2) buy issue slots?
// Chains per thread: NF x FFMA or FFMA2, interleaved with NL x LOP3 and NW x (mul.wide.u32 + xor).
// Reports cycles per step per sub-partition and FMAs per clock per SM (bench evidence, standalone).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <bool kPacked, int NF, int NL, int NW>
__global__ void __launch_bounds__(256, 2) kern(int iters, float a, float b, uint32_t la, uint32_t lb, float *out)
{
    unsigned long long x2[NF];
    float x1[NF];
    uint32_t l[NL > 0 ? NL : 1], w[NW > 0 ? NW : 1], y[NW > 0 ? NW : 1];
    const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b);
    const unsigned long long aa = *reinterpret_cast<const unsigned long long *>(&a2);
    const unsigned long long bb = *reinterpret_cast<const unsigned long long *>(&b2);
    for (int i = 0; i < NF; i++) {
        float2 v = make_float2(1.0f + threadIdx.x * 1e-3f + i, 0.5f + i);
        x2[i] = *reinterpret_cast<unsigned long long *>(&v);
        x1[i] = v.x;
    }
    for (int i = 0; i < NL; i++) l[i] = 12345u + threadIdx.x + i;
    for (int i = 0; i < NW; i++) { w[i] = 777u + threadIdx.x + i; y[i] = i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < NF; i++) {
                if (kPacked)
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x2[i]) : "l"(aa), "l"(bb));
                else
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x1[i]) : "f"(a), "f"(b));
                if (i < NL) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(l[i]) : "r"(la), "r"(lb));
                if (i < NW) {
                    unsigned long long p;
                    asm volatile("mul.wide.u32 %0, %1, 0xD2511F53;" : "=l"(p) : "r"(w[i]));
                    w[i] = (uint32_t)p ^ y[i];
                    y[i] = (uint32_t)(p >> 32);
                }
            }
        }
    }
    float acc = 0;
    for (int i = 0; i < NF; i++) { float2 v = *reinterpret_cast<float2 *>(&x2[i]); acc += v.x + v.y + x1[i]; }
    for (int i = 0; i < NL; i++) acc += l[i];
    for (int i = 0; i < NW; i++) acc += w[i] + y[i];
    if (acc == 123.456f) out[0] = acc;
}

template <bool kPacked, int NF, int NL, int NW>
void run(int sms)
{
    float *out;
    cudaMalloc(&out, 4);
    const int iters = 2048, blocks = sms * 2;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        kern<kPacked, NF, NL, NW><<<blocks, 256>>>(iters, 0.999f, 1e-3f, 0x9E3779B9u, 0xBB67AE85u, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double cycles = best * 1e-3 * 1.965e9;
    const double groups = (double)iters * 8 * 4;  // 4 warps per sub-partition (2 CTAs x 8 warps / 4)
    const double fmas = (double)iters * 8 * NF * (kPacked ? 2 : 1) * 256.0 * blocks;
    std::printf("%-6s x%d + lop3 x%d + (mul.wide+xor) x%d : %7.3f ms  %6.2f cycles per group  %6.1f FMA/clk/SM\n", kPacked ? "ffma2" : "ffma",
                NF, NL, NW, best, cycles / groups, fmas / cycles / sms);
    cudaFree(out);
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    run<false, 8, 0, 0>(sms);
    run<true, 8, 0, 0>(sms);
    run<false, 8, 2, 0>(sms);
    run<true, 8, 2, 0>(sms);
    run<false, 8, 4, 0>(sms);
    run<true, 8, 4, 0>(sms);
    run<false, 8, 0, 1>(sms);
    run<true, 8, 0, 1>(sms);
    run<false, 8, 0, 2>(sms);
    run<true, 8, 0, 2>(sms);
    run<false, 8, 2, 2>(sms);
    run<true, 8, 2, 2>(sms);
    return 0;
}
