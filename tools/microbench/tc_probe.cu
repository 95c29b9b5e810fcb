// tc_probe.cu -- stand-alone check of the tcgen05 building blocks the tensor-core basket kernel is
// made of: K-major un-swizzled shared-memory operand layout, shared-memory and instruction
// descriptors for kind::tf32, TMEM allocation, single-thread MMA issue, commit -> mbarrier, and
// tcgen05.ld of one accumulator row per thread.  D[128 x 64] = A[128 x 64] * B[64 x 64]^T with the
// 3xTF32 split (hi*hi + lo*hi + hi*lo), compared on the host with an fp64 product.
// Stand-alone development tool, not part of libmcb200.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

constexpr int kM = 128, kN = 64, kK = 64;
// K-major, no swizzle, in bytes: element (row, k) of an R-row operand sits at
//   (k / 4) * lbo + (row / 8) * 128 + (row % 8) * 16 + (k % 4) * 4,   lbo = (R / 8) * 128
// (core matrix = 8 rows x 16 bytes, contiguous; 8-row groups 128 bytes apart; 16-byte K chunks lbo apart)
__host__ __device__ constexpr int operand_offset(int rows, int row, int k)
{
    return (k / 4) * (rows / 8) * 128 + (row / 8) * 128 + (row % 8) * 16 + (k % 4) * 4;
}

__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    // cute::UMMA::SmemDescriptor: start >> 4 [0,14), LBO >> 4 [16,30), SBO >> 4 [32,46), version 1 [46,48),
    // base offset 0, layout type SWIZZLE_NONE (0) [61,64)
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46);
}

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate));
}

__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    return ok != 0;
}


__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate));
}

#define TMEM_ST_X32(taddr, v, o)                                                                                                      \
    asm volatile(                                                                                                                     \
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "   \
        "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),                                   \
        "r"(v[o + 0]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), "r"(v[o + 5]), "r"(v[o + 6]), "r"(v[o + 7]),       \
        "r"(v[o + 8]), "r"(v[o + 9]), "r"(v[o + 10]), "r"(v[o + 11]), "r"(v[o + 12]), "r"(v[o + 13]), "r"(v[o + 14]), "r"(v[o + 15]), \
        "r"(v[o + 16]), "r"(v[o + 17]), "r"(v[o + 18]), "r"(v[o + 19]), "r"(v[o + 20]), "r"(v[o + 21]), "r"(v[o + 22]),               \
        "r"(v[o + 23]), "r"(v[o + 24]), "r"(v[o + 25]), "r"(v[o + 26]), "r"(v[o + 27]), "r"(v[o + 28]), "r"(v[o + 29]),               \
        "r"(v[o + 30]), "r"(v[o + 31])                                                                                                \
        : "memory")

#define TMEM_LD_X32(taddr, r, o)                                                                                                      \
    asm volatile(                                                                                                                     \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "     \
        "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                                               \
        : "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]),             \
          "=r"(r[o + 7]), "=r"(r[o + 8]), "=r"(r[o + 9]), "=r"(r[o + 10]), "=r"(r[o + 11]), "=r"(r[o + 12]), "=r"(r[o + 13]),         \
          "=r"(r[o + 14]), "=r"(r[o + 15]), "=r"(r[o + 16]), "=r"(r[o + 17]), "=r"(r[o + 18]), "=r"(r[o + 19]), "=r"(r[o + 20]),      \
          "=r"(r[o + 21]), "=r"(r[o + 22]), "=r"(r[o + 23]), "=r"(r[o + 24]), "=r"(r[o + 25]), "=r"(r[o + 26]), "=r"(r[o + 27]),      \
          "=r"(r[o + 28]), "=r"(r[o + 29]), "=r"(r[o + 30]), "=r"(r[o + 31])                                                          \
        : "r"(taddr))

// mode bit 0: swap the LBO / SBO fields of the shared-memory descriptors
// mode bit 1: A operand from tensor memory (tcgen05.st by the thread that owns the row) instead of shared memory
// mode bit 2: (with bit 1) the production schedule: two K halves of 32 through ONE 32-column A_hi/A_lo buffer, the second
//             half restricted to outputs 32..63 (lower-triangular B), TMEM = 128 columns: D [0,64), A_hi [64,96), A_lo [96,128)
__global__ void __launch_bounds__(128, 1) probe_kernel(const float *A, const float *B, float *D, int mode)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    float *a_hi = reinterpret_cast<float *>(smem);
    float *a_lo = reinterpret_cast<float *>(smem + 32768);
    float *b_hi = reinterpret_cast<float *>(smem + 65536);
    float *b_lo = reinterpret_cast<float *>(smem + 65536 + 16384);
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool swap = mode & 1, a_tmem = mode & 2, halves = mode & 4;

    for (int idx = tid; idx < kM * kK; idx += blockDim.x) {
        const int m = idx / kK, k = idx % kK;
        const float v = A[idx];
        const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        a_hi[operand_offset(kM, m, k) / 4] = hi;
        a_lo[operand_offset(kM, m, k) / 4] = v - hi;
    }
    for (int idx = tid; idx < kN * kK; idx += blockDim.x) {
        const int n = idx / kK, k = idx % kK;
        const float v = B[idx];
        const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        b_hi[operand_offset(kN, n, k) / 4] = hi;
        b_lo[operand_offset(kN, n, k) / 4] = v - hi;
    }
    // generic-proxy writes -> visible to the async proxy (the tensor core reads shared memory through it)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&mbar);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base)),
                     "n"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    // idesc (cute::UMMA::InstrDescriptor): D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major,
    // N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc64 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
    const uint32_t idesc32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
    const uint32_t a_hi_s = (uint32_t)__cvta_generic_to_shared(a_hi), a_lo_s = (uint32_t)__cvta_generic_to_shared(a_lo);
    const uint32_t b_hi_s = (uint32_t)__cvta_generic_to_shared(b_hi), b_lo_s = (uint32_t)__cvta_generic_to_shared(b_lo);
    constexpr uint32_t lbo_a = (kM / 8) * 128, lbo_b = (kN / 8) * 128;
    auto desc = [&](uint32_t addr, uint32_t lbo) { return swap ? smem_desc(addr, 128, lbo) : smem_desc(addr, lbo, 128); };
    bool done = true;
    uint32_t phase = 0;
    auto wait = [&]() {
        bool ok = false;
        for (int spin = 0; spin < (1 << 22) && !ok; spin++)  // bounded: a wrong descriptor must not hang the GPU box
            ok = mbar_try_wait(bar, phase);
        phase ^= 1;
        done = done && ok;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    };

    if (!a_tmem) {
        if (tid == 0) {
            uint32_t acc = 0;
            for (int ks = 0; ks < kK / 8; ks++) {  // one MMA covers K = 8 tf32 = two 16-byte chunks
                const uint64_t ah = desc(a_hi_s + ks * 2 * lbo_a, lbo_a), al = desc(a_lo_s + ks * 2 * lbo_a, lbo_a);
                const uint64_t bh = desc(b_hi_s + ks * 2 * lbo_b, lbo_b), bl = desc(b_lo_s + ks * 2 * lbo_b, lbo_b);
                mma_tf32(tmem, ah, bh, idesc64, acc);
                acc = 1;
                mma_tf32(tmem, al, bh, idesc64, 1);
                mma_tf32(tmem, ah, bl, idesc64, 1);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
        }
        wait();
    } else {
        // this thread owns row m = tid of A
        const int nh = halves ? 2 : 1, kh = halves ? 32 : 64;
        for (int h = 0; h < nh; h++) {
            uint32_t vhi[64], vlo[64];
#pragma unroll
            for (int k = 0; k < 64; k++) {
                const float v = (k < kh) ? A[tid * kK + h * kh + k] : 0.f;
                const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
                vhi[k] = __float_as_uint(hi);
                vlo[k] = __float_as_uint(v - hi);
            }
            const uint32_t t_ahi = tmem + lane_base + 64, t_alo = tmem + lane_base + (halves ? 96 : 128);
            TMEM_ST_X32(t_ahi, vhi, 0);
            TMEM_ST_X32(t_alo, vlo, 0);
            if (!halves) {
                TMEM_ST_X32(t_ahi + 32, vhi, 32);
                TMEM_ST_X32(t_alo + 32, vlo, 32);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t n0 = (h == 1) ? 32 : 0;       // second half only reaches outputs 32..63
                const uint32_t idesc = (h == 1) ? idesc32 : idesc64;
                const uint32_t brow = (n0 / 8) * 128;        // byte offset of row n0 inside a K chunk
                for (int ks = 0; ks < kh / 8; ks++) {
                    const int kg = (h * kh) / 8 + ks;        // global K step (for B)
                    const uint64_t bh = desc(b_hi_s + brow + kg * 2 * lbo_b, lbo_b), bl = desc(b_lo_s + brow + kg * 2 * lbo_b, lbo_b);
                    const uint32_t ah = tmem + 64 + ks * 8, al = tmem + (halves ? 96 : 128) + ks * 8;
                    mma_tf32_ts(tmem + n0, ah, bh, idesc, (h > 0 || ks > 0) ? 1u : 0u);
                    mma_tf32_ts(tmem + n0, al, bh, idesc, 1);
                    mma_tf32_ts(tmem + n0, ah, bl, idesc, 1);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
            }
            wait();  // A buffer free again / accumulator complete
        }
    }
    if (!done) {
        if (tid == 0)
            D[0] = -12345.0f;  // marker: the MMA never completed
    } else {
        uint32_t r[64];
        const uint32_t taddr = tmem + lane_base;
        TMEM_LD_X32(taddr, r, 0);
        TMEM_LD_X32(taddr + 32, r, 32);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < 64; c++)
            D[(warp * 32 + lane) * kN + c] = __uint_as_float(r[c]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256));
}

int main()
{
    std::vector<float> A(kM * kK), B(kN * kK), D(kM * kN, 0.f);
    srand(1);
    for (auto &v : A) v = (float)rand() / RAND_MAX * 8.f - 4.f;      // normals-like
    for (int n = 0; n < kN; n++)
        for (int k = 0; k < kK; k++)
            B[n * kK + k] = (k <= n) ? (float)rand() / RAND_MAX * 0.4f - 0.2f : 0.f;  // lower-triangular factor
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4);
    cudaMalloc(&dB, B.size() * 4);
    cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    const int smem = 65536 + 32768;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int failures = 0;
    const int modes[] = {0, 2, 6};
    for (int mode : modes) {
        cudaMemset(dD, 0, D.size() * 4);
        probe_kernel<<<1, 128, smem>>>(dA, dB, dD, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            std::printf("mode %d: CUDA error: %s\n", mode, cudaGetErrorString(e));
            return 1;
        }
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        if (D[0] == -12345.0f) {
            std::printf("mode %d: MMA never completed (mbarrier wait timed out)\n", mode);
            failures++;
            continue;
        }
        double max_err = 0, max_ref = 0;
        for (int m = 0; m < kM; m++)
            for (int n = 0; n < kN; n++) {
                double ref = 0;
                for (int k = 0; k < kK; k++)
                    ref += (double)A[m * kK + k] * (double)B[n * kK + k];
                max_err = std::fmax(max_err, std::fabs(ref - (double)D[m * kN + n]));
                max_ref = std::fmax(max_ref, std::fabs(ref));
            }
        std::printf("mode %d (%s, %s%s): 3xTF32 128x64x64 max |err| = %.3e (max |ref| = %.3f) -> %s\n", mode,
                    (mode & 1) ? "LBO/SBO swapped" : "LBO = K-chunk stride, SBO = 8-row stride", (mode & 2) ? "A in TMEM" : "A in smem",
                    (mode & 4) ? ", two K halves + triangular N" : "", max_err, max_ref, max_err < 2e-5 ? "OK" : "MISMATCH");
        if (!(max_err < 2e-5))
            failures++;
    }
    std::printf("failures: %d\n", failures);
    return 0;
}
