// pipe_mul.cu -- what does a 32x32->64 multiply cost on sm_100a?  The Philox round needs both halves
// of two such products; this measures the candidate encodings (bench evidence, standalone).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int V>
__global__ void __launch_bounds__(256, 4) mul_kernel(int iters, uint32_t m_reg, uint32_t *out)
{
    uint32_t x[4], y[4];
    for (int i = 0; i < 4; i++) { x[i] = 12345u + threadIdx.x + i; y[i] = 777u + i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (V == 0) {  // mul.wide.u32 by an immediate, both halves consumed by the next step
                    uint64_t p;
                    asm volatile("mul.wide.u32 %0, %1, 0xD2511F53;" : "=l"(p) : "r"(x[i]));
                    x[i] = (uint32_t)p ^ y[i];
                    y[i] = (uint32_t)(p >> 32);
                } else if (V == 1) {  // mul.wide.u32 by a register
                    uint64_t p;
                    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x[i]), "r"(m_reg));
                    x[i] = (uint32_t)p ^ y[i];
                    y[i] = (uint32_t)(p >> 32);
                } else if (V == 2) {  // mul.hi + mul.lo separately
                    uint32_t hi, lo;
                    asm volatile("mul.hi.u32 %0, %1, 0xD2511F53;" : "=r"(hi) : "r"(x[i]));
                    asm volatile("mul.lo.u32 %0, %1, 0xD2511F53;" : "=r"(lo) : "r"(x[i]));
                    x[i] = lo ^ y[i];
                    y[i] = hi;
                } else if (V == 3) {  // mul.hi only
                    uint32_t hi;
                    asm volatile("mul.hi.u32 %0, %1, 0xD2511F53;" : "=r"(hi) : "r"(x[i]));
                    x[i] = hi ^ y[i];
                } else if (V == 4) {  // mul.lo only (+ xor), the reference point
                    uint32_t lo;
                    asm volatile("mul.lo.u32 %0, %1, 0xD2511F53;" : "=r"(lo) : "r"(x[i]));
                    x[i] = lo ^ y[i];
                } else if (V == 5) {  // xor only
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(m_reg));
                }
            }
        }
    }
    uint32_t acc = 0;
    for (int i = 0; i < 4; i++) acc += x[i] + y[i];
    if (acc == 123456789u) out[0] = acc;
}

template <int V>
void run(const char *name, int sms)
{
    uint32_t *out;
    cudaMalloc(&out, 4);
    const int iters = 2048, blocks = sms * 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        mul_kernel<V><<<blocks, 256>>>(iters, 0xCD9E8D57u, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double cycles = best * 1e-3 * 1.965e9;
    const double steps = (double)iters * 8 * 4 * 8;  // per sub-partition: 8 warps x 4 chains
    std::printf("%-48s %7.3f ms  %5.2f cycles per step per sub-partition\n", name, best, cycles / steps);
    cudaFree(out);
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    run<5>("lop3 (reference: 2 cycles)", sms);
    run<4>("mul.lo.u32 imm + xor", sms);
    run<3>("mul.hi.u32 imm + xor", sms);
    run<2>("mul.hi.u32 + mul.lo.u32 imm + xor", sms);
    run<0>("mul.wide.u32 imm + xor", sms);
    run<1>("mul.wide.u32 reg + xor", sms);
    return 0;
}
