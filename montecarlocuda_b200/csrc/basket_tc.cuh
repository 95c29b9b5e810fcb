// basket_tc.cuh -- wide fp32 basket (32 < N <= 64) with the correlation mat-vec on the 5th-generation
// tensor cores (tcgen05, sm_100a).  Included by kernels_basket.cu, which owns the constant table.
//
// Same estimator as Basket<float, 64, *> (brownianVect + basketPayoff + basketOptMonteCarlo,
// DP/MonteCarloKernel.cu:74-101, :133-177) and the same Philox stream: draw block j of a path's
// sub-stream gives normals 6j .. 6j+5.  What changes is who multiplies by the factor.  On the FFMA
// path the triangular mat-vec is 1056 FFMA2 + 565 LDS.128 of the 3144 instructions of a path
// (profiles/r01f_basket64_f32_smem.txt): over a tile of 128 paths it is the dense contraction
//     X[128 paths x 64 assets] = Z[128 x 64 normals] * F^T,      F = diag(v sqrt(T)) L  (log2 units)
// so it goes to the tensor cores and the CUDA cores keep what only they can do (Philox, Box-Muller, 2^x).
//
//   * precision: kind::tf32 keeps 11 significant bits of each operand, far too few for a price at 3
//     standard errors, so both operands are split exactly, v = hi + lo with hi = v & 0xffffe000, and
//     three products are accumulated in fp32: hi*hi + lo*hi + hi*lo ("3xTF32"; the dropped lo*lo term
//     and the truncation of lo are < 2^-21 relative: measured 7.6e-6 absolute on |x| <= 8,
//     profiles/r01h_tc_probe.txt, the same order as the MUFU chain around it).
//   * A operand (the normals) never touches shared memory: the thread that owns path m writes its row
//     with tcgen05.st into tensor memory (lane m, one column per normal) and the MMA reads A from there.
//   * B operand (the factor, hi and lo parts) sits in shared memory for the whole kernel, K-major,
//     un-swizzled canonical layout (core matrix = 8 rows x 16 bytes).
//   * the accumulator D[128 x 64] lives in tensor memory; tcgen05.ld 32x32b hands every thread the 64
//     exponents of ITS path, so the payoff needs no cross-thread traffic either.
//   * K is walked in two halves of 32 normals through one 32-column A buffer, so a tile needs
//     64 (D) + 32 (A hi) + 32 (A lo) = 128 tensor-memory columns; the SM's one CTA runs four tiles side by side
//     (all 512 columns): two sub-blocks of 256 worker threads, two tiles each.  For a triangular factor the second half only
//     reaches assets 32..63: its MMAs run with N = 32 on the upper half of D.
//   * pipeline: the MMAs of a half (12 x UMMA 128xNx8) are issued by one thread per tile and complete
//     asynchronously (tcgen05.commit -> mbarrier) while all threads generate the next normals: the second half's
//     while the first half multiplies, and the first three Philox blocks of the NEXT round while the second half
//     multiplies (the payoff of a round is collected after them).  The hand-off "A buffer written" is a shared-memory
//     counter: the warp of the tile that arrives last issues the MMAs, the other three go on.
//
// The chunk structure, the per-thread accumulation order and the exact-integer combine are those of
// mc_accumulate_kernel (device_common.cuh): price and half-width stay bit-identical for any grid
// shape and GPU count.  They are NOT bit-identical to the FFMA kernel's (different rounding of the
// mat-vec), which is why the engine choice is part of the job description (mcb200_set_basket_engine).
#pragma once

#include "device_math.cuh"

namespace mcb {

constexpr int kTcWidth = 64;                       // assets (padded) == normals per path
constexpr int kTcFactorBytes = kTcWidth * kTcWidth * 4;
constexpr int kTcABase = kTcFactorBytes;           // a_i  (log2 units)
constexpr int kTcMBase = kTcABase + kTcWidth * 4;  // m_i = w_i s_i
constexpr int kTcKBase = kTcMBase + kTcWidth * 4;  // strike
struct BasketTcTable {
    float f[kTcWidth * kTcWidth];  // row-major F[i][k], zero above the diagonal / in the padding
    float a[kTcWidth];
    float m[kTcWidth];
    float k;
};

// K-major, un-swizzled canonical operand layout (cute: ((8,n),(4,2)):((16 B, SBO),(4 B, LBO))):
// element (row, k) of an R-row operand, in bytes
__host__ __device__ constexpr int tc_operand_offset(int rows, int row, int k)
{
    return (k / 4) * (rows / 8) * 128 + (row / 8) * 128 + (row % 8) * 16 + (k % 4) * 4;
}
constexpr uint32_t kTcLbo = (kTcWidth / 8) * 128;  // between the 16-byte K chunks
constexpr uint32_t kTcSbo = 128;                   // between 8-row groups

struct BasketTcShared {
    __align__(128) float b_hi[kTcWidth * kTcWidth];
    __align__(128) float b_lo[kTcWidth * kTcWidth];
    __align__(8) unsigned long long mbar[4];  // one per tile (128 threads)
    unsigned int arrivals[4][2];              // per (tile, K half): warps that have written their A columns, ever
    uint32_t tmem_base;
};

namespace tc {

__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr)
{
    // cute::UMMA::SmemDescriptor: start >> 4 [0,14), LBO >> 4 [16,30), SBO >> 4 [32,46), version 1 [46,48),
    // base offset 0, layout SWIZZLE_NONE (0) [61,64)
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(kTcLbo >> 4) << 16) | ((uint64_t)(kTcSbo >> 4) << 32) | (1ull << 46);
}
// cute::UMMA::InstrDescriptor: D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major, N >> 3 at bit 17,
// M >> 4 at bit 24
__host__ __device__ constexpr uint32_t instr_desc(int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); }

// D[tmem] (+)= A[tmem] * B[smem]^T, 128 x N x 8, issued by one thread
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
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
// Bounded: a tensor-core operation that never completes must end as an error flag, not as a hung device.
// `dead` is sticky, so a failed launch drains in milliseconds.
__device__ __forceinline__ void wait_phase(uint32_t bar, uint32_t parity, bool &dead)
{
    if (!dead) {
        bool ok = false;
        for (int spin = 0; spin < (1 << 20) && !ok; spin++)
            ok = mbar_try_wait(bar, parity);
        dead = !ok;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// "A buffer written" hand-off inside a tile (4 warps): every warp counts itself in with an acquire-release
// shared-memory atomic once its A columns are written; the warp that arrives LAST -- the one the tile would have
// waited for anyway -- enqueues the tile's MMAs, the other three go straight on drawing normals.  The counter only
// grows (4 per hand-off), so nothing has to be reset between rounds.  No extra warp: 16 warps per SM, 4 per
// scheduler, leave the full 128 registers per thread, and this kernel lives on instruction-level parallelism
// (with a dedicated 17th issuing warp one scheduler holds 5 warps and ptxas is capped at 96 registers).
__device__ __forceinline__ bool tile_arrive_is_last(unsigned int *counter)
{
    unsigned int old = 0;
    if ((threadIdx.x & 31) == 0)
        asm volatile("atom.acq_rel.cta.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"((uint32_t)__cvta_generic_to_shared(counter)) : "memory");
    old = __shfl_sync(0xffffffffu, old, 0);
    return (old & 3u) == 3u;
}

// one lane of a converged warp, chosen by the hardware (elect.sync): unlike `lane == 0` the compiler knows the
// branch holds exactly one thread, so a UMMA inside it is issued straight from uniform registers
__device__ __forceinline__ bool elect_one()
{
    uint32_t picked;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(picked));
    return picked != 0;
}

// 16 consecutive columns of this thread's tensor-memory lane
__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
// 32 consecutive columns of this thread's lane, as 16 packed pairs {col 2i, col 2i + 1}
__device__ __forceinline__ void ld32(uint32_t taddr, unsigned long long (&d)[16])
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
        "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++)
        asm("mov.b64 %0, {%1, %2};" : "=l"(d[i]) : "r"(r[2 * i]), "r"(r[2 * i + 1]));
}

template <int kByteOffset> __device__ __forceinline__ unsigned long long const_pair()
{
    unsigned long long t;
    asm volatile("ld.const.b64 %0, [mcb_basket_table+%1];" : "=l"(t) : "n"(kByteOffset));
    return t;
}
template <int kByteOffset> __device__ __forceinline__ float const_f32()
{
    float t;
    asm volatile("ld.const.f32 %0, [mcb_basket_table+%1];" : "=f"(t) : "n"(kByteOffset));
    return t;
}

}  // namespace tc

// Per-thread view of its tile: tensor-memory addresses and the tile's mbarrier.
struct BasketTcTile {
    uint32_t lane_d;   // D columns [0,64) of this thread's lane
    uint32_t tile_d;   // same, lane field 0 (MMA operand)
    uint32_t bar;      // shared-memory address of the tile's mbarrier
    uint32_t b_hi, b_lo;
    int tile;          // 0 .. 3 inside the CTA
    unsigned int *arrivals;  // this tile's two hand-off counters
    bool dead;         // a tensor-core wait timed out
};

__device__ __forceinline__ BasketTcTile basket_tc_setup(BasketTcShared &sh)
{
    const int tid = threadIdx.x, warp = tid >> 5;
    // factor: constant table -> hi / lo parts in the operand layout (once per CTA)
    const float *f = reinterpret_cast<const float *>(mcb_basket_table);
    for (int idx = tid; idx < kTcWidth * kTcWidth; idx += blockDim.x) {
        const int n = idx / kTcWidth, k = idx % kTcWidth;
        const float v = f[idx];
        const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        sh.b_hi[tc_operand_offset(kTcWidth, n, k) / 4] = hi;
        sh.b_lo[tc_operand_offset(kTcWidth, n, k) / 4] = v - hi;
    }
    // generic-proxy writes -> visible to the async proxy (the tensor core reads shared memory through it)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(&sh.tmem_base)),
                     "n"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid < 8)
        (&sh.arrivals[0][0])[tid] = 0u;
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 4; i++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&sh.mbar[i])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    BasketTcTile t;
    // warp-uniform by construction, and said so to the compiler (a shuffle from lane 0): the MMA issue code then takes
    // its tensor-memory addresses from uniform registers instead of an ELECT / R2UR.BROADCAST loop around every UMMA
    t.tile = __shfl_sync(0xffffffffu, (tid >> 7) & 3, 0);  // sub-block * 2 + half of the sub-block
    t.tile_d = __shfl_sync(0xffffffffu, sh.tmem_base + (uint32_t)t.tile * 128u, 0);
    t.lane_d = t.tile_d + ((uint32_t)((warp & 3) * 32) << 16);
    t.bar = (uint32_t)__cvta_generic_to_shared(&sh.mbar[t.tile]);
    t.arrivals = &sh.arrivals[t.tile][0];
    t.b_hi = (uint32_t)__cvta_generic_to_shared(sh.b_hi);
    t.b_lo = (uint32_t)__cvta_generic_to_shared(sh.b_lo);
    t.dead = false;
    return t;
}

__device__ __forceinline__ void basket_tc_teardown(BasketTcShared &sh)
{
    tc::fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(sh.tmem_base), "n"(512) : "memory");
}

// 16 normals -> exact hi / lo split -> 16 columns of the A_hi and A_lo buffers of this thread's lane
__device__ __forceinline__ void basket_tc_store16(const float *z, uint32_t lane_d, int col)
{
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
        hi[i] = __float_as_uint(z[i]) & 0xffffe000u;
        hi[i + 1] = __float_as_uint(z[i + 1]) & 0xffffe000u;
        // lo = z - hi (exact), two columns per packed instruction
        float l0, l1;
        sub_f32x2(z[i], z[i + 1], __uint_as_float(hi[i]), __uint_as_float(hi[i + 1]), l0, l1);
        lo[i] = __float_as_uint(l0);
        lo[i + 1] = __float_as_uint(l1);
    }
    tc::st16(lane_d + 64 + col, hi);
    tc::st16(lane_d + 96 + col, lo);
}

// Normal generation of one K half (32 normals), in two stages so that the kernel can slide the first stage of the
// NEXT round in front of the wait for this round's accumulator.
// A Philox block gives six normals (device_math.cuh), so the halves do not fall on block boundaries: half 0
// draws blocks 0..5 (normals 0..35) and hands normals 32..35 to half 1, which draws blocks 6..10 (36..65, the
// last two unused).  W[i] is normal 32 * kHalf + i.
template <int kHalf>
struct BasketTcHalf {
    static constexpr int kNpb = NormalsPerBlock<float>::value;
    static_assert(kNpb == 6, "the block schedule below is written for six normals per Philox block");
    static constexpr int kPre = kHalf == 0 ? 0 : 4;            // normals inherited from the previous half
    static constexpr int kFirstBlock = kHalf == 0 ? 0 : 6;
    static constexpr int kBlocksHere = kHalf == 0 ? 6 : 5;
    static constexpr int kBlocksFirst16 = (16 - kPre + kNpb - 1) / kNpb;  // blocks needed before columns 0..15 are complete
    static constexpr int kWindow = kPre + kNpb * kBlocksHere;

    static __device__ __forceinline__ void draw(const PhiloxKeys &keys, uint32_t path_lo, uint32_t path_hi, int lb, float *W)
    {
        const NoShared none;
        uint32_t w[4];
        philox4x32_10(path_lo, path_hi, (uint32_t)(kFirstBlock + lb), kTagBasket, keys, w);
        float z[kNpb];
        normals_from_words(w, z, none);
#pragma unroll
        for (int i = 0; i < kNpb; i++)
            W[kPre + kNpb * lb + i] = z[i];
    }
    // stage 1: the blocks that complete columns 0..15 (W[0 .. kPre + 6 * kBlocksFirst16))
    static __device__ __forceinline__ void first(const PhiloxKeys &keys, uint32_t path_lo, uint32_t path_hi, float *W,
                                                 const float (&carry)[4])
    {
        if (kPre) {
#pragma unroll
            for (int i = 0; i < kPre; i++)
                W[i] = carry[i];
        }
#pragma unroll
        for (int lb = 0; lb < kBlocksFirst16; lb++)
            draw(keys, path_lo, path_hi, lb, W);
    }
    // stage 2: store columns 0..15, draw the rest, store columns 16..31, count this warp in; the last warp of the tile
    // to do so enqueues the tile's 12 MMAs.  The A buffer must be free when this is called.
    template <bool kFull>
    static __device__ __forceinline__ void rest(const PhiloxKeys &keys, uint32_t path_lo, uint32_t path_hi, float *W,
                                                float (&carry)[4], BasketTcTile &t)
    {
        basket_tc_store16(W, t.lane_d, 0);
#pragma unroll
        for (int lb = kBlocksFirst16; lb < kBlocksHere; lb++)
            draw(keys, path_lo, path_hi, lb, W);
        basket_tc_store16(W + 16, t.lane_d, 16);
        if (kHalf == 0) {
#pragma unroll
            for (int i = 0; i < 4; i++)
                carry[i] = W[32 + i];
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc::fence_before();
        if (tc::tile_arrive_is_last(t.arrivals + kHalf)) {
            // the tile's A buffer is complete: one lane enqueues the 12 MMAs (3xTF32 over four K steps of 8) + commit
            if (tc::elect_one()) {
                tc::fence_after();
                constexpr bool kUpper = kHalf == 1 && !kFull;     // a triangular factor: normals 32..63 only reach assets 32..63
                constexpr uint32_t n0 = kUpper ? 32 : 0;
                constexpr uint32_t idesc = tc::instr_desc(kUpper ? 32 : 64);
                constexpr uint32_t brow = (n0 / 8) * 128;          // byte offset of row n0 inside a K chunk
#pragma unroll
                for (int ks = 0; ks < 4; ks++) {                   // one MMA covers K = 8 tf32 = two 16-byte chunks
                    const uint32_t koff = brow + (uint32_t)(kHalf * 4 + ks) * 2u * kTcLbo;
                    const uint64_t bh = tc::smem_desc(t.b_hi + koff), bl = tc::smem_desc(t.b_lo + koff);
                    const uint32_t ah = t.tile_d + 64 + ks * 8, al = t.tile_d + 96 + ks * 8;
                    tc::mma_ts(t.tile_d + n0, ah, bh, idesc, (kHalf > 0 || ks > 0) ? 1u : 0u);
                    tc::mma_ts(t.tile_d + n0, al, bh, idesc, 1u);
                    tc::mma_ts(t.tile_d + n0, ah, bl, idesc, 1u);
                }
                tc::commit(t.bar);
            }
            __syncwarp();
        }
    }
};

template <int... kI>
__device__ __forceinline__ void basket_tc_payoff_half(const unsigned long long (&d)[16], unsigned long long &sum2,
                                                      std::integer_sequence<int, kI...>, std::integral_constant<int, 0>)
{
    (([&] {
         unsigned long long x2;
         asm("add.rn.f32x2 %0, %1, %2;" : "=l"(x2) : "l"(d[kI]), "l"(tc::const_pair<kTcABase + kI * 8>()));
         const float e0 = mufu_ex2(__uint_as_float((uint32_t)x2)), e1 = mufu_ex2(__uint_as_float((uint32_t)(x2 >> 32)));
         asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(sum2) : "l"(tc::const_pair<kTcMBase + kI * 8>()), "l"(pack2(e0, e1)));
     }()),
     ...);
}
template <int... kI>
__device__ __forceinline__ void basket_tc_payoff_half(const unsigned long long (&d)[16], unsigned long long &sum2,
                                                      std::integer_sequence<int, kI...>, std::integral_constant<int, 1>)
{
    (([&] {
         unsigned long long x2;
         asm("add.rn.f32x2 %0, %1, %2;" : "=l"(x2) : "l"(d[kI]), "l"(tc::const_pair<kTcABase + 128 + kI * 8>()));
         const float e0 = mufu_ex2(__uint_as_float((uint32_t)x2)), e1 = mufu_ex2(__uint_as_float((uint32_t)(x2 >> 32)));
         asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(sum2) : "l"(tc::const_pair<kTcMBase + 128 + kI * 8>()), "l"(pack2(e0, e1)));
     }()),
     ...);
}

// Everything of a round after the first draws of its half 0 (W0 filled by BasketTcHalf<0>::first): both halves'
// stores and MMAs.  All 256 threads of the CTA call this together.
template <bool kFull>
__device__ __forceinline__ void basket_tc_enqueue(const PhiloxKeys &keys, uint32_t path_lo, uint32_t path_hi, float *W0,
                                                  BasketTcTile &t)
{
    float carry[4];
    BasketTcHalf<0>::rest<kFull>(keys, path_lo, path_hi, W0, carry, t);
    float W1[BasketTcHalf<1>::kWindow];
    BasketTcHalf<1>::first(keys, path_lo, path_hi, W1, carry);
    tc::wait_phase(t.bar, 0, t.dead);  // the first half's MMAs have read the A buffer
    BasketTcHalf<1>::rest<kFull>(keys, path_lo, path_hi, W1, carry, t);
}

// The payoff of the round whose MMAs were enqueued last: waits for the accumulator (which also frees the A buffer).
__device__ __forceinline__ float basket_tc_collect(BasketTcTile &t)
{
    tc::wait_phase(t.bar, 1, t.dead);
    unsigned long long sum2 = pack2(-tc::const_f32<kTcKBase>(), 0.0f);
    {
        unsigned long long d[16];
        tc::ld32(t.lane_d, d);
        basket_tc_payoff_half(d, sum2, std::make_integer_sequence<int, 16>{}, std::integral_constant<int, 0>{});
    }
    {
        unsigned long long d[16];
        tc::ld32(t.lane_d + 32, d);
        basket_tc_payoff_half(d, sum2, std::make_integer_sequence<int, 16>{}, std::integral_constant<int, 1>{});
    }
    tc::fence_before();  // the next round's MMAs (issued after the tile hand-off) overwrite D
    return positive_part(__uint_as_float((uint32_t)sum2) + __uint_as_float((uint32_t)(sum2 >> 32)));
}

struct BasketTcParams {
    PhiloxKeys keys;
};

// mc_accumulate_kernel with the tile machinery around it.  ONE CTA per SM: two sub-blocks of 256 threads (each
// the "block" of the stream definition with its own chunk walk and scratch; every thread runs every round, paths
// beyond the job's total are computed and not counted), i.e. four tiles on the SM's 512 tensor-memory columns and
// exactly 16 warps, which leaves 128 registers per thread.  History of the MMA issue code (~12 instructions per UMMA:
// descriptors, uniform-register moves, election): on a fixed worker warp the other three warps of its tile waited
// for it a quarter of the time (profiles/r01i_basket64_f32_tensor_worker_issue.txt); on a dedicated 17th warp the
// register cap fell to 96 and the latency-bound worker code lost its instruction-level parallelism (at 138
// registers ONE CTA of 288 threads per SM was only 14 % slower than two at 96, profiles/r01i_tc_experiments.txt);
// now the last warp to arrive at a hand-off issues.
constexpr int kTcSubBlocks = 2;
constexpr int kTcThreads = kTcSubBlocks * kThreads;

template <bool kFull>
__global__ void __launch_bounds__(kTcThreads, 1)
basket_tc_accumulate_kernel(const __grid_constant__ BasketTcParams P, const __grid_constant__ Geometry G,
                            unsigned long long *__restrict__ acc)
{
    __shared__ BlockScratch scs[kTcSubBlocks];
    __shared__ BasketTcShared sh;
    pdl_launch_dependents();
    const int sub = (int)(threadIdx.x / kThreads), tid = (int)(threadIdx.x % kThreads);
    BlockScratch &sc = scs[sub];
    if (tid < kAccWords)
        sc.acc[tid] = 0ull;
    BasketTcTile t = basket_tc_setup(sh);  // ends with a CTA-wide barrier
    for (ChunkWalk<kTcSubBlocks> walk(sub); walk.live(G); walk.advance(sc)) {
        walk.claim_ahead(G, tid);
        const unsigned long long base = (G.first_chunk + walk.chunk) * G.chunk_units;
        const bool whole = base + G.chunk_units <= G.total_paths;
        const unsigned long long n_valid = whole ? G.chunk_units : (G.total_paths > base ? G.total_paths - base : 0ull);
        float s = 0, s2 = 0;
        const uint32_t lo0 = (uint32_t)base + tid, hi = (uint32_t)(base >> 32);
        const float no_carry[4] = {0.f, 0.f, 0.f, 0.f};
        float W0[BasketTcHalf<0>::kWindow];
        BasketTcHalf<0>::first(P.keys, lo0, hi, W0, no_carry);
#pragma unroll 1
        for (int k = 0; k < G.rounds; k++) {
            const unsigned long long unit = base + (unsigned long long)k * kThreads + tid;
            basket_tc_enqueue<kFull>(P.keys, lo0 + (uint32_t)(k * kThreads), hi, W0, t);
            // software pipeline: the next round's first draws run while this round's last MMAs complete
            if (k + 1 < G.rounds)
                BasketTcHalf<0>::first(P.keys, lo0 + (uint32_t)((k + 1) * kThreads), hi, W0, no_carry);
            const float v = basket_tc_collect(t);
            if (whole || unit < G.total_paths) {
                s += v;
                s2 = fmaf(v, v, s2);
            }
        }
        chunk_commit<kTcSubBlocks>((double)s, (double)s2, n_valid, G, sc, sub, tid, walk.ahead);
    }
    if (t.dead && (tid & 127) == 0)
        atomicAdd(&sc.acc[11], 1ull);
    __syncthreads();
    finish(sc, acc, G, tid);
    basket_tc_teardown(sh);
}

// Per-path values of units [first_unit, first_unit + n_units) through the same tile machinery.
template <bool kFull>
__global__ void __launch_bounds__(kTcThreads, 1)
basket_tc_paths_kernel(const __grid_constant__ BasketTcParams P, unsigned long long first_unit, unsigned long long n_units,
                       float *__restrict__ out)
{
    __shared__ BasketTcShared sh;
    const int sub = (int)(threadIdx.x / kThreads), tid = (int)(threadIdx.x % kThreads);
    BasketTcTile t = basket_tc_setup(sh);
    const unsigned long long n_blocks = (n_units + kThreads - 1) / kThreads;
    const unsigned long long stride = (unsigned long long)gridDim.x * kTcSubBlocks;
    for (unsigned long long blk = (unsigned long long)blockIdx.x * kTcSubBlocks + sub; blk < n_blocks; blk += stride) {
        const unsigned long long i = blk * kThreads + tid;
        const unsigned long long unit = first_unit + i;
        const float no_carry[4] = {0.f, 0.f, 0.f, 0.f};
        float W0[BasketTcHalf<0>::kWindow];
        BasketTcHalf<0>::first(P.keys, (uint32_t)unit, (uint32_t)(unit >> 32), W0, no_carry);
        basket_tc_enqueue<kFull>(P.keys, (uint32_t)unit, (uint32_t)(unit >> 32), W0, t);
        const float v = basket_tc_collect(t);
        if (i < n_units)
            out[i] = t.dead ? __int_as_float(0x7fc00000) : v;
    }
    basket_tc_teardown(sh);
}

}  // namespace mcb
