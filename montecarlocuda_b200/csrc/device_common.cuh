// device_common.cuh -- generator, chunk geometry and the order-free reduction shared by the
// three pricing kernels (sm_100a).
//
// Replaces, for every workload, the reference's per-thread XORWOW state + randomSetup kernel
// (DP/MonteCarloKernel.cu:285-290, :189), its shared-memory tree reduction (:157-176, :200-219,
// :263-282) and the sequential host sum over blocks (:416-419).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>
#include <type_traits>

namespace mcb {

constexpr int kThreads = 256;  // threads per CTA == units per chunk round (part of the stream definition)
constexpr int kWarps = kThreads / 32;
constexpr int kLanes = 5;
constexpr int kAccWords = 12;

// Philox4x32-10 round keys, precomputed on the host (key + i * Weyl): read as constant-bank
// operands of the three-input XOR, so the key schedule costs no instructions per draw.
struct PhiloxKeys {
    uint32_t k0[10];
    uint32_t k1[10];
};

// Peer-memory combine fused into the pricing kernel (multi-GPU, one process per GPU or one process for all).
// Every rank owns a small mailbox in its device memory that its peers can write over NVLink (CUDA IPC /
// peer access); mail[r] is rank r's mailbox as addressed from THIS device.  Slot (seq % kPeerRing, src) holds
// src's 12 accumulator words as 24 "flagged halves": every 8-byte mailbox word is {low: 32 bits of data, high: the
// launch's flag}, written by ONE 8-byte store (atomic), so a word is valid exactly when its flag matches -- no
// fence, no separate "ready" flag and no second NVLink round trip (the LL protocol of NCCL).
constexpr int kPeerMax = 8;
constexpr int kPeerRing = 4;
constexpr int kPeerHalves = 2 * kAccWords;   // 24 flagged halves per (slot, source)
constexpr int kPeerSlotWords = 32;           // 24 used, padded to 256 bytes
constexpr size_t kPeerMailboxBytes = (size_t)kPeerRing * kPeerMax * kPeerSlotWords * sizeof(unsigned long long);
struct PeerLink {
    int world;                            // <= 1: no combine, the kernel leaves this device's partial in acc
    int rank;
    unsigned long long seq;               // this launch's sequence number (>= 1, the same on every rank)
    unsigned long long *mail[kPeerMax];
    unsigned int *ticket;                 // this device's "CTAs finished" counter (zero between launches)
};

// What a launch covers.  All of it derives from the JOB (total paths), never from the GPU count.
struct Geometry {
    unsigned long long total_paths;
    unsigned long long chunk_units;   // kThreads * rounds
    unsigned long long first_chunk;   // this shard
    unsigned long long n_chunks;
    int rounds;                       // units per thread per chunk
    int scale_exp_sum;                // value * 2^e before the integer split
    int scale_exp_sumsq;
    PeerLink peer;
};

// One Philox4x32-10 block.  The 64-bit products compile to IMAD.WIDE.U32, the mixes to LOP3.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              const PhiloxKeys &K, uint32_t (&out)[4])
{
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
    for (int i = 0; i < 10; i++) {
        const unsigned long long p0 = (unsigned long long)M0 * c0;
        const unsigned long long p1 = (unsigned long long)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ K.k0[i];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ K.k1[i];
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Split a non-negative double, scaled by 2^e, into five 32-bit limbs (a 160-bit fixed-point
// window) and add them to `lanes`.  Every step is exact except the final truncation below the
// window's least significant bit, so sums of limbs are order-independent.  Returns false when
// the value is negative, NaN or does not fit.
__device__ __forceinline__ bool lanes_add(double value, int scale_exp, unsigned long long *lanes)
{
    const double t = scalbn(value, scale_exp);
    const double th = t * 0x1p-96;
    if (!(value >= 0.0) || !(th < 0x1p63))
        return false;
    const unsigned long long hi = __double2ull_rz(th);
    const double rem = t - __ull2double_rn(hi) * 0x1p96;
    const unsigned long long mid = __double2ull_rz(rem * 0x1p-32);
    const double lo_d = rem - __ull2double_rn(mid) * 0x1p32;
    const unsigned long long lo = __double2ull_rz(lo_d);
    lanes[0] += lo;
    lanes[1] += mid & 0xffffffffull;
    lanes[2] += mid >> 32;
    lanes[3] += hi & 0xffffffffull;
    lanes[4] += hi >> 32;
    return true;
}

struct BlockScratch {
    double s[kWarps];
    double s2[kWarps];
    unsigned long long acc[kAccWords];
};

// Barrier of one sub-block: a CTA of the pricing kernel is kSubBlocks independent groups of kThreads threads
// (each the "block" of the stream definition) that only share read-only tables.  One group: __syncthreads.
template <int kSubBlocks>
__device__ __forceinline__ void sub_barrier(int sub)
{
    if constexpr (kSubBlocks == 1)
        __syncthreads();
    else if (sub == 0)
        asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory");
    else if (sub == 1)
        asm volatile("bar.sync 2, %0;" ::"n"(kThreads) : "memory");
    else if (sub == 2)
        asm volatile("bar.sync 3, %0;" ::"n"(kThreads) : "memory");
    else
        asm volatile("bar.sync 4, %0;" ::"n"(kThreads) : "memory");
    static_assert(kSubBlocks <= 4, "one named barrier per sub-block");
}

template <int kSubBlocks = 1>
__device__ __forceinline__ void scratch_init(BlockScratch &sc, int sub = 0, int tid = threadIdx.x)
{
    if (tid < kAccWords)
        sc.acc[tid] = 0ull;
    sub_barrier<kSubBlocks>(sub);
}

// Fixed-shape reduction of one chunk: xor butterfly inside each warp (offsets 16..1, every lane
// ends with the same value), then warps 0..7 in order on thread 0, which turns the chunk partial
// into integer limbs held in shared memory.  The chunk partial depends only on the chunk's
// per-path values, not on which CTA, SM or GPU ran it.
template <int kSubBlocks = 1>
__device__ __forceinline__ void chunk_commit(double s, double s2, unsigned long long n_valid,
                                             const Geometry &G, BlockScratch &sc, int sub = 0, int tid = threadIdx.x)
{
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, off);
        s2 += __shfl_xor_sync(0xffffffffu, s2, off);
    }
    const int warp = tid >> 5;
    if ((tid & 31) == 0) {
        sc.s[warp] = s;
        sc.s2[warp] = s2;
    }
    sub_barrier<kSubBlocks>(sub);
    if (tid == 0) {
        double S = sc.s[0], S2 = sc.s2[0];
#pragma unroll
        for (int w = 1; w < kWarps; w++) {
            S += sc.s[w];
            S2 += sc.s2[w];
        }
        bool ok = lanes_add(S, G.scale_exp_sum, sc.acc);
        ok = lanes_add(S2, G.scale_exp_sumsq, sc.acc + kLanes) && ok;
        sc.acc[10] += n_valid;
        if (!ok)
            sc.acc[11] += 1ull;
    }
    sub_barrier<kSubBlocks>(sub);
}

// After the (sub-)block's last chunk: 12 integer atomics.  Integer addition is associative, so
// the device-wide (and, after the combine, job-wide) totals do not depend on arrival order.
__device__ __forceinline__ void scratch_flush(const BlockScratch &sc, unsigned long long *acc, int tid = threadIdx.x)
{
    if (tid < kAccWords) {
        const unsigned long long v = sc.acc[tid];
        if (v != 0ull)
            atomicAdd(acc + tid, v);
    }
}

// ---- the collective, fused into the kernel that feeds it ----------------------------------------
// The (sum, sum^2) combine across GPUs is 96 bytes: as a separate NCCL all-reduce it costs a kernel launch
// and ~20-30 us of latency after a pricing kernel that, sharded over 8 GPUs, runs for 0.5-10 ms.  Here the
// LAST CTA of each device's pricing kernel pushes the device's limbs straight into every peer's mailbox
// (24 flagged 8-byte stores per peer over NVLink), polls its own mailbox for the peers' and adds the
// integer limbs: when the kernel ends, acc holds the JOB's totals on every rank, bit-identical everywhere
// (integer addition; the order of arrival cannot matter).  Slot reuse is safe with a ring of 2 or more: a rank
// cannot finish step s + 1 before every peer has pushed step s + 1, which each peer does after reading step s.
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ void peer_combine(unsigned long long *acc, const PeerLink &L)
{
    __shared__ bool s_last;
    __shared__ unsigned int s_late;
    __shared__ unsigned int s_half[kPeerMax * kPeerHalves];
    __threadfence();  // this CTA's atomics on acc are ordered before its ticket
    __syncthreads();
    if (threadIdx.x == 0) {
        s_last = atomicAdd(L.ticket, 1u) == gridDim.x - 1;
        s_late = 0u;
    }
    __syncthreads();
    if (!s_last)
        return;
    // ---- last CTA of this device ----
    __threadfence();
    const int tid = threadIdx.x;
    if (tid == 0)
        *L.ticket = 0u;  // ready for the next launch on this device (stream order)
    const size_t slot = (size_t)(L.seq % kPeerRing) * kPeerMax;
    // never 0 (fresh mailbox memory) and different from the flag this slot carried kPeerRing launches ago
    const unsigned long long flag = ((unsigned long long)((unsigned int)L.seq | 0x80000000u)) << 32;
    if (tid < kPeerHalves * L.world) {
        const int peer = tid / kPeerHalves, h = tid % kPeerHalves;
        // push: half h of this device's accumulator into rank `peer`'s mailbox
        const unsigned long long word = *((volatile unsigned long long *)acc + (h >> 1));
        const unsigned long long half = (h & 1) ? (word >> 32) : (word & 0xffffffffull);
        st_relaxed_sys(L.mail[peer] + (slot + L.rank) * kPeerSlotWords + h, flag | half);
        // pull: half h of rank `peer`'s accumulator from this device's mailbox.  Bounded: a peer that never
        // launches must end as an error flag, not as a hung device.
        const unsigned long long *src = L.mail[L.rank] + (slot + peer) * kPeerSlotWords + h;
        unsigned long long v = 0ull;
        bool ok = false;
        for (int spin = 0; spin < (1 << 22) && !ok; spin++) {
            v = ld_relaxed_sys(src);
            ok = (v & 0xffffffff00000000ull) == flag;
            if (!ok)
                __nanosleep(64);
        }
        if (!ok)
            atomicAdd(&s_late, 1u);
        s_half[tid] = (unsigned int)v;
    }
    __syncthreads();
    if (tid < kAccWords) {
        unsigned long long total = 0ull;
        for (int src = 0; src < L.world; src++)
            total += (unsigned long long)s_half[src * kPeerHalves + 2 * tid] |
                     ((unsigned long long)s_half[src * kPeerHalves + 2 * tid + 1] << 32);
        if (tid == kAccWords - 1 && s_late)
            total += 1ull;  // error flag: a peer did not answer
        acc[tid] = total;
    }
}

// end of a pricing kernel: CTA totals -> device accumulator (-> job totals on every rank)
__device__ __forceinline__ void finish(const BlockScratch &sc, unsigned long long *acc, const Geometry &G, int tid = threadIdx.x)
{
    scratch_flush(sc, acc, tid);
    if (G.peer.world > 1)
        peer_combine(acc, G.peer);
}

// The pricing kernel skeleton.  W is a workload policy:
//   W::Real            float or double: the per-path arithmetic type
//   W::Params          by-value parameter block (constant bank)
//   W::kUnitPaths      paths served by one draw unit
//   W::kMinBlocks      CTAs per SM the register budget is sized for; W::kUnroll  unroll of the unit loop
//   W::Shared          per-CTA shared-memory state (the fp64 math tables; empty for fp32)
//   W::eval(P, unit_lo, unit_hi, v, sh) fills v[kUnitPaths] with the per-path values of a draw unit;
//                      inside a chunk unit_hi is the same for every thread (chunks are aligned), so the
//                      part of the first two Philox rounds that depends only on it runs on the uniform datapath
// A CTA walks chunks first_chunk + blockIdx.x, + gridDim.x, ...; thread t of a chunk owns units
// base + k * 256 + t for k < rounds, in that order, and accumulates value and value^2 in W::Real
// (short runs: at most rounds * kUnitPaths <= 384 terms) before the fp64 block reduction.
template <class W>
__global__ void __launch_bounds__(kThreads * W::kSubBlocks, W::kMinBlocks)
mc_accumulate_kernel(const __grid_constant__ typename W::Params P, const __grid_constant__ Geometry G,
                     unsigned long long *__restrict__ acc)
{
    using Real = typename W::Real;
    constexpr int kSub = W::kSubBlocks;
    // W::Shared in dynamic shared memory (the replicated fp64 tables are 96 KB); nothing for fp32
    extern __shared__ __align__(16) unsigned char mcb_dynamic_smem[];
    typename W::Shared &sh = *reinterpret_cast<typename W::Shared *>(mcb_dynamic_smem);
    __shared__ BlockScratch scs[kSub];
    // sub-block = the "block" of the stream definition: kThreads threads, its own scratch, its own chunks
    const int sub = kSub == 1 ? 0 : (int)(threadIdx.x / kThreads);
    const int tid = kSub == 1 ? (int)threadIdx.x : (int)(threadIdx.x % kThreads);
    BlockScratch &sc = scs[sub];
    sh.load();
    if (tid < kAccWords)
        sc.acc[tid] = 0ull;
    __syncthreads();
    const unsigned long long last = G.first_chunk + G.n_chunks;
    const unsigned long long stride = (unsigned long long)gridDim.x * kSub;
    for (unsigned long long chunk = G.first_chunk + (unsigned long long)blockIdx.x * kSub + sub; chunk < last; chunk += stride) {
        const unsigned long long base = chunk * G.chunk_units;
        const unsigned long long path_end = (base + G.chunk_units) * (unsigned long long)W::kUnitPaths;
        Real s = 0, s2 = 0;
        unsigned long long n_valid;
        const bool whole = path_end <= G.total_paths;
        if (whole) {
            n_valid = G.chunk_units * (unsigned long long)W::kUnitPaths;
        } else {
            const unsigned long long path0 = base * (unsigned long long)W::kUnitPaths;
            n_valid = G.total_paths > path0 ? G.total_paths - path0 : 0ull;
        }
        if (whole || W::kUnitPaths == 1) {
            // one path per unit: the same loop serves the job's last (partial) chunk, so the
            // (large) estimator body is instantiated once
#pragma unroll W::kUnroll
            for (int k = 0; k < G.rounds; k++) {
                const unsigned long long unit = base + (unsigned long long)k * kThreads + tid;
                if (W::kUnitPaths == 1 && !whole && unit >= G.total_paths)
                    break;
                Real v[W::kUnitPaths];
                W::eval(P, (uint32_t)base + (uint32_t)(k * kThreads) + tid, (uint32_t)(base >> 32), v, sh);
#pragma unroll
                for (int q = 0; q < W::kUnitPaths; q++) {
                    s += v[q];
                    s2 = fma(v[q], v[q], s2);
                }
            }
        } else {
            // the job's last chunk with several paths per unit: mask paths beyond the total
#pragma unroll 1
            for (int k = 0; k < G.rounds; k++) {
                const unsigned long long unit = base + (unsigned long long)k * kThreads + tid;
                if (unit * (unsigned long long)W::kUnitPaths >= G.total_paths)
                    break;
                Real v[W::kUnitPaths];
                W::eval(P, (uint32_t)base + (uint32_t)(k * kThreads) + tid, (uint32_t)(base >> 32), v, sh);
#pragma unroll
                for (int q = 0; q < W::kUnitPaths; q++) {
                    if (unit * (unsigned long long)W::kUnitPaths + q < G.total_paths) {
                        s += v[q];
                        s2 = fma(v[q], v[q], s2);
                    }
                }
            }
        }
        chunk_commit<kSub>((double)s, (double)s2, n_valid, G, sc, sub, tid);
    }
    if constexpr (kSub > 1)
        __syncthreads();  // every sub-block has committed its last chunk before the CTA-wide tail
    finish(sc, acc, G, tid);
}

// Host side of a launch: dynamic shared memory = W::Shared (opt-in above 48 KB, set once per instantiation).
template <class W> constexpr size_t accumulate_smem_bytes() { return std::is_empty<typename W::Shared>::value ? 0 : sizeof(typename W::Shared); }
template <class W> inline cudaError_t accumulate_prepare()
{
    // a function attribute is per device: set it once on each device this process prices on
    static bool done[64] = {};
    if (accumulate_smem_bytes<W>() <= 48 * 1024)
        return cudaSuccess;
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess || device < 0 || device >= 64)
        return e != cudaSuccess ? e : cudaErrorInvalidDevice;
    if (!done[device]) {
        e = cudaFuncSetAttribute(mc_accumulate_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)accumulate_smem_bytes<W>());
        if (e != cudaSuccess)
            return e;
        done[device] = true;
    }
    return cudaSuccess;
}
template <class W>
inline cudaError_t accumulate_launch(int grid, const typename W::Params &p, const Geometry &g, unsigned long long *d_acc,
                                     cudaStream_t stream)
{
    cudaError_t e = accumulate_prepare<W>();
    if (e != cudaSuccess)
        return e;
    mc_accumulate_kernel<W><<<grid, kThreads * W::kSubBlocks, accumulate_smem_bytes<W>(), stream>>>(p, g, d_acc);
    return cudaGetLastError();
}
template <class W> inline int accumulate_blocks_per_sm()
{
    int n = 0;
    if (accumulate_prepare<W>() != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, mc_accumulate_kernel<W>, kThreads * W::kSubBlocks,
                                                      accumulate_smem_bytes<W>()) != cudaSuccess)
        return 0;
    return n;
}

// Per-path values of units [first_unit, first_unit + n_units): the same W::eval as above.
template <class W>
__global__ void __launch_bounds__(kThreads)
mc_paths_kernel(const __grid_constant__ typename W::Params P, unsigned long long first_unit,
                unsigned long long n_units, typename W::Real *__restrict__ out)
{
    __shared__ typename W::Shared sh;
    sh.load();
    __syncthreads();
    for (unsigned long long i = blockIdx.x * (unsigned long long)kThreads + threadIdx.x; i < n_units;
         i += (unsigned long long)gridDim.x * kThreads) {
        typename W::Real v[W::kUnitPaths];
        W::eval(P, (uint32_t)(first_unit + i), (uint32_t)((first_unit + i) >> 32), v, sh);
#pragma unroll
        for (int q = 0; q < W::kUnitPaths; q++)
            out[i * W::kUnitPaths + q] = v[q];
    }
}

}  // namespace mcb
