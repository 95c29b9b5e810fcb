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
#include <utility>

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
// Two modes: kPeerWait -- the last CTA pushes, waits for its peers' halves and leaves the JOB's totals in the output
// block (the kernel is the collective); kPeerPush -- it only pushes and ends (split phase): ranks do not lock-step on
// the slowest one, and the totals of a launch are summed out of the mailbox when somebody asks for them
// (peer_pull_kernel).
constexpr int kPeerMax = 8;
constexpr int kPeerRing = 8;
constexpr int kPeerHalves = 2 * kAccWords;   // 24 flagged halves per (slot, source)
constexpr int kPeerSlotWords = 32;           // 24 used, padded to 256 bytes
constexpr size_t kPeerMailboxBytes = (size_t)kPeerRing * kPeerMax * kPeerSlotWords * sizeof(unsigned long long);
enum : int { kPeerWait = 0, kPeerPush = 1 };
constexpr unsigned long long kErrPeerTimeout = 1ull << 32;   // added to accumulator word 11: a peer never answered
struct PeerLink {
    int world;                            // <= 1: no combine, the kernel leaves this device's partial in the output block
    int rank;
    int mode;                             // kPeerWait / kPeerPush
    unsigned long long seq;               // this launch's sequence number (>= 1, the same on every rank)
    unsigned long long timeout_ns;        // bound of the wait for the peers (wall clock, %globaltimer)
    unsigned long long *mail[kPeerMax];
};
__host__ __device__ inline unsigned long long peer_flag(unsigned long long seq)
{
    // never 0 (fresh mailbox memory) and different from the flags this slot carried on its previous uses
    return ((unsigned long long)((unsigned int)seq | 0x80000000u)) << 32;
}

// Per-launch control block in device memory.  All zero between launches: the last CTA of a launch takes the totals
// out and leaves it so (no memset in front of a launch).
struct LaunchCtl {
    unsigned long long acc[kAccWords];    // integer limbs added by the CTAs of this launch
    unsigned int next;                    // chunks claimed beyond the first gridDim.x * kSubBlocks
    unsigned int ticket;                  // CTAs that have finished
    unsigned long long pad[3];
};
static_assert(sizeof(LaunchCtl) == 128, "one control block per 128 bytes");

// The part of a launch that depends on the JOB only (total paths), never on the GPU count or the grid.
struct JobGeometry {
    unsigned long long total_paths;
    unsigned long long chunk_units;   // kThreads * rounds
    int rounds;                       // units per thread per chunk
    int scale_exp_sum;                // value * 2^e before the integer split
    int scale_exp_sumsq;
    int pad;
};

// What a launch covers and where its result goes.
struct Geometry : JobGeometry {
    unsigned long long first_chunk;   // this shard
    unsigned long long n_chunks;      // < 2^31 per launch (the host splits larger shards)
    LaunchCtl *ctl;
    // result, written by the launch's last CTA: 24 flagged halves (like a mailbox slot) into mapped pinned HOST
    // memory, so a blocking call needs no copy and no stream synchronisation -- the host polls the flags
    unsigned long long *host_slot;    // or nullptr
    unsigned long long host_flag;     // peer_flag(launch number of the context)
    PeerLink peer;
};

// Programmatic dependent launch (sm_90+): the next kernel of the stream may start filling the SMs as soon as every
// CTA of this one has started (its tail overlaps our tail), and nothing of ours is complete for it until pdl_wait().
// Both are no-ops for a kernel launched without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

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
// Integer arithmetic on the double's fields: V = floor(m 2^(x + e)) for value = m 2^x touches at most three limbs,
// found by a shift count -- ~30 short-latency instructions and no branches on the way, where scalbn and three
// conversions each way (this function's first version; the oracle still restates it that way) kept the rest of the
// sub-block waiting at the chunk's second barrier.
__device__ __forceinline__ bool lanes_add(double value, int scale_exp, unsigned long long *lanes)
{
    const unsigned long long bits = (unsigned long long)__double_as_longlong(value);
    const int field = (int)(bits >> 52) & 0x7ff;
    const unsigned long long frac = bits & 0x000fffffffffffffull;
    const unsigned long long m = field ? (frac | 0x0010000000000000ull) : frac;     // value = m 2^x
    const int p = (field ? field - 1075 : -1074) + scale_exp;                          // V = floor(m 2^p)
    if (m == 0ull)
        return field != 0x7ff;                                                         // +-0: nothing to add
    const int top = 64 - __clzll((long long)m) + p;                                    // V < 2^top
    if ((long long)bits < 0 || field == 0x7ff || top > 159)
        return false;
    if (p < 0) {
        const unsigned long long v = p > -64 ? m >> (-p) : 0ull;
        lanes[0] += v & 0xffffffffull;
        lanes[1] += v >> 32;
        return true;
    }
    const int limb = p >> 5, off = p & 31;                                             // limb <= 4 because top <= 159
    const unsigned long long lo = m << off, hi = off > 11 ? m >> (64 - off) : 0ull;    // m < 2^53: hi < 2^21
    lanes[limb] += lo & 0xffffffffull;
    if (limb < 4)
        lanes[limb + 1] += lo >> 32;
    if (limb < 3)
        lanes[limb + 2] += hi;
    return true;
}

struct BlockScratch {
    double s[kWarps];
    double s2[kWarps];
    unsigned long long acc[kAccWords];
    unsigned int next;   // the chunk this sub-block runs after the current one (relative to the shard)
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
// per-path values, not on which CTA, SM or GPU ran it.  `next` (thread 0's) is handed to the whole sub-block
// through sc.next on the way: the chunk it runs after this one.
template <int kSubBlocks = 1>
__device__ __forceinline__ void chunk_commit(double s, double s2, unsigned long long n_valid,
                                             const JobGeometry &G, BlockScratch &sc, int sub = 0, int tid = threadIdx.x,
                                             unsigned int next = 0u)
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
    // the serial part, on two threads of different warps side by side: sum -> thread 0, sum of squares -> thread 32
    if (tid == 0) {
        double S = sc.s[0];
#pragma unroll
        for (int w = 1; w < kWarps; w++)
            S += sc.s[w];
        const bool ok = lanes_add(S, G.scale_exp_sum, sc.acc);
        sc.acc[10] += n_valid;
        if (!ok)
            atomicAdd(&sc.acc[11], 1ull);
        sc.next = next;
    } else if (tid == 32) {
        double S2 = sc.s2[0];
#pragma unroll
        for (int w = 1; w < kWarps; w++)
            S2 += sc.s2[w];
        if (!lanes_add(S2, G.scale_exp_sumsq, sc.acc + kLanes))
            atomicAdd(&sc.acc[11], 1ull);
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

// Which chunk a sub-block runs next.  The first one is static (blockIdx.x * kSubBlocks + sub); the following ones are
// claimed from the launch's counter, ONE AHEAD: thread 0 asks for the next chunk when it starts the current one and
// only looks at the answer when the current one is committed, so the L2 round trip of the atomic never waits.
// Legal because the combine is order-free (integer limbs): which CTA ran which chunk cannot change a bit.  Against
// the static stride this removes the wave tail -- 10 922 chunks on 1 184 CTAs are 9.2 waves, and whoever drew ten
// chunks kept the device waiting (profiles/r01k_scale_notes.txt).
template <int kSubBlocks>
struct ChunkWalk {
    unsigned int chunk;   // relative to G.first_chunk
    unsigned int ahead;   // thread 0 only: the claim in flight
    __device__ __forceinline__ ChunkWalk(int sub) : chunk(blockIdx.x * kSubBlocks + sub), ahead(0u) {}
    __device__ __forceinline__ bool live(const Geometry &G) const { return chunk < (unsigned int)G.n_chunks; }
    __device__ __forceinline__ void claim_ahead(const Geometry &G, int tid)
    {
#ifdef MCB_STATIC_STRIDE   // A/B switch (tools/build_variant.sh): the static stride of the first version
        ahead = chunk + gridDim.x * kSubBlocks;
#else
        if (tid == 0)
            ahead = gridDim.x * kSubBlocks + atomicAdd(&G.ctl->next, 1u);
#endif
    }
    __device__ __forceinline__ void advance(const BlockScratch &sc) { chunk = sc.next; }
};

// ---- the collective, fused into the kernel that feeds it ----------------------------------------
// The (sum, sum^2) combine across GPUs is 96 bytes: as a separate NCCL all-reduce it costs a kernel launch
// and ~20-30 us of latency after a pricing kernel that, sharded over 8 GPUs, runs for 0.5-10 ms.  Here the
// LAST CTA of each device's pricing kernel pushes the device's limbs straight into every peer's mailbox
// (24 flagged 8-byte stores per peer over NVLink) and either ends there (kPeerPush) or polls its own mailbox
// for the peers' and adds the integer limbs (kPeerWait): then the output block holds the JOB's totals on every
// rank when the kernel ends, bit-identical everywhere (integer addition; the order of arrival cannot matter).
// Slot reuse: in kPeerWait a rank cannot finish step s + 1 before every peer has pushed step s + 1, which each does
// after reading step s; in kPeerPush nothing holds a rank back, a slot is simply overwritten kPeerRing launches
// later, and a pull that comes too late sees a newer flag and says so instead of waiting for ever.
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
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned long long flagged_half(const unsigned long long *words, int h, unsigned long long flag)
{
    const unsigned long long w = words[h >> 1];
    return flag | ((h & 1) ? (w >> 32) : (w & 0xffffffffull));
}

// Wait (bounded, wall clock) for half h of source `src` of launch `seq` in this device's mailbox.
// Returns 0 = arrived, 1 = timed out, 2 = overwritten by a later launch (a pull that came too late).
__device__ __forceinline__ int peer_poll(const unsigned long long *mailbox, unsigned long long seq, int src, int h,
                                         unsigned long long timeout_ns, unsigned int &data)
{
    const unsigned long long flag = peer_flag(seq);
    const unsigned long long *p = mailbox + ((size_t)(seq % kPeerRing) * kPeerMax + src) * kPeerSlotWords + h;
    unsigned long long t0 = 0ull;
    for (unsigned int spin = 0;; spin++) {
        const unsigned long long v = ld_relaxed_sys(p);
        if ((v & 0xffffffff00000000ull) == flag) {
            data = (unsigned int)v;
            return 0;
        }
        // a flag of a LATER launch of the same slot (31-bit sequence numbers, compared modulo 2^31)
        const unsigned int theirs = (unsigned int)(v >> 32), mine = (unsigned int)(flag >> 32);
        if ((theirs & 0x80000000u) && ((theirs - mine) & 0x7fffffffu) != 0u && ((theirs - mine) & 0x7fffffffu) < 0x40000000u) {
            data = 0u;
            return 2;
        }
        if ((spin & 63u) == 63u) {
            const unsigned long long now = globaltimer_ns();
            if (t0 == 0ull)
                t0 = now;
            else if (now - t0 > timeout_ns) {
                data = 0u;
                return 1;
            }
        }
        __nanosleep(32);
    }
}

// last CTA of a device: tot[] (shared memory, this device's 12 limbs) -> every mailbox; kPeerWait: += the peers'
__device__ __forceinline__ void peer_exchange(unsigned long long *tot, const PeerLink &L)
{
    __shared__ unsigned int s_late;
    __shared__ unsigned int s_half[kPeerMax * kPeerHalves];
    const int tid = threadIdx.x;
    if (tid == 0)
        s_late = 0u;
    __syncthreads();
    const unsigned long long flag = peer_flag(L.seq);
    const size_t slot = (size_t)(L.seq % kPeerRing) * kPeerMax;
    if (tid < kPeerHalves * L.world) {
        const int peer = tid / kPeerHalves, h = tid % kPeerHalves;
        // push: half h of this device's limbs into rank `peer`'s mailbox
        st_relaxed_sys(L.mail[peer] + (slot + L.rank) * kPeerSlotWords + h, flagged_half(tot, h, flag));
        if (L.mode == kPeerWait) {
            // pull: half h of rank `peer`'s limbs from this device's mailbox.  Bounded: a peer that never launches
            // must end as an error flag, not as a hung device.
            unsigned int v;
            if (peer_poll(L.mail[L.rank], L.seq, peer, h, L.timeout_ns, v) != 0)
                atomicAdd(&s_late, 1u);
            s_half[tid] = v;
        }
    }
    if (L.mode != kPeerWait)
        return;
    __syncthreads();
    if (tid < kAccWords) {
        unsigned long long total = 0ull;
        for (int src = 0; src < L.world; src++)
            total += (unsigned long long)s_half[src * kPeerHalves + 2 * tid] |
                     ((unsigned long long)s_half[src * kPeerHalves + 2 * tid + 1] << 32);
        if (tid == kAccWords - 1 && s_late)
            total += kErrPeerTimeout;
        tot[tid] = total;
    }
    __syncthreads();
}

// End of a pricing kernel: CTA totals -> the launch's control block; the LAST CTA takes the device totals out of it
// (leaving it zeroed for the next launch), runs the cross-GPU exchange if there is one, adds the result to the
// caller's accumulator block and/or publishes it to the host.
__device__ __forceinline__ void finish(const BlockScratch &sc, unsigned long long *out, const Geometry &G, int tid = threadIdx.x)
{
    __shared__ bool s_last;
    __shared__ unsigned long long s_tot[kAccWords];
    pdl_wait();   // the previous launch of the stream is complete before anything of this one becomes visible
    scratch_flush(sc, G.ctl->acc, tid);
    __threadfence();  // this CTA's atomics are ordered before its ticket
    __syncthreads();
    if (threadIdx.x == 0)
        s_last = atomicAdd(&G.ctl->ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last)
        return;
    // ---- last CTA of this device ----
    __threadfence();
    if (threadIdx.x < kAccWords)
        s_tot[threadIdx.x] = atomicExch(&G.ctl->acc[threadIdx.x], 0ull);
    if (threadIdx.x == 0) {
        G.ctl->next = 0u;
        G.ctl->ticket = 0u;
    }
    __syncthreads();
    if (G.peer.world > 1)
        peer_exchange(s_tot, G.peer);
    if (G.peer.world > 1 && G.peer.mode == kPeerPush)
        return;   // the totals are in the mailboxes; peer_pull_kernel delivers them
    if (out != nullptr && threadIdx.x < kAccWords && s_tot[threadIdx.x] != 0ull)
        atomicAdd(out + threadIdx.x, s_tot[threadIdx.x]);
    if (G.host_slot != nullptr && threadIdx.x < kPeerHalves)
        st_relaxed_sys(G.host_slot + threadIdx.x, flagged_half(s_tot, threadIdx.x, G.host_flag));
}

// value -> (sum, sum of squares).  kClamp: v stands for max(v, 0) (a payoff the workload left unclamped).  fp64 has no
// cheap max (DSETP + selects on the pipe that binds; a sign test with predicated accumulation is turned into four
// selects by ptxas), so a negative v only has its HIGH word clamped -- one integer max -- which leaves a number below
// 2^-1042 in place of the zero: its square is exactly 0, and it cannot change a sum that holds anything else (a sum of
// nothing but such leftovers truncates to zero limbs in lanes_add).  The limbs are those of the exact clamp, which the
// per-path kernel applies (mc_paths_kernel), and the tests compare the two bit for bit.
__device__ __forceinline__ bool sign_bit_set(double v) { return __double2hiint(v) < 0; }
__device__ __forceinline__ bool sign_bit_set(float v) { return __float_as_int(v) < 0; }
template <bool kClamp>
__device__ __forceinline__ void add_value(float v, float &s, float &s2)
{
    if (kClamp)
        v = fmaxf(v, 0.0f);
    s += v;
    s2 = fmaf(v, v, s2);
}
template <bool kClamp>
__device__ __forceinline__ void add_value(double v, double &s, double &s2)
{
    if (kClamp)
        v = __hiloint2double(max(__double2hiint(v), 0), __double2loint(v));
    s += v;
    s2 = fma(v, v, s2);
}

// fp32 with an even number of paths per draw unit (the European call: six): a thread keeps TWO interleaved pairs of
// running sums -- the even paths of every unit go into one, the odd paths into the other, and one float addition joins
// them at the end of the chunk -- so that the accumulations of two paths are ONE FADD2 and ONE FFMA2 (packed fp32,
// sm_100) instead of two FADD and two FFMA: the fp32 call is bound by issue slots.  This order is part
// of the stream definition; the oracle and the debug reduction restate it (orc_chunk_reduce, debug_reduce_kernel).
struct PackedSums {
    unsigned long long s = 0ull, s2 = 0ull;   // {even paths, odd paths} of value and of value^2
    __device__ __forceinline__ void add(float even, float odd)
    {
        unsigned long long v;
        asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(even), "f"(odd));
        asm("add.rn.f32x2 %0, %0, %1;" : "+l"(s) : "l"(v));
        asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(s2) : "l"(v));
    }
    static __device__ __forceinline__ float joined(unsigned long long pair)
    {
        float even, odd;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(even), "=f"(odd) : "l"(pair));
        return even + odd;
    }
};
template <class W> struct accumulates_packed {
    static constexpr bool value = std::is_same<typename W::Real, float>::value && W::kUnitPaths % 2 == 0 && !W::kClampAtZero;
};

// One chunk of the job: thread t owns units base + k * 256 + t for k < rounds, in that order, and accumulates value
// and value^2 in W::Real (short runs: at most rounds * kUnitPaths <= 384 terms) before the fp64 block reduction.
template <class W>
__device__ __forceinline__ void run_chunk(const typename W::Params &P, const JobGeometry &G, unsigned long long chunk, int tid,
                                          const typename W::Shared &sh, const typename W::JobState &job,
                                          typename W::Real &s_out, typename W::Real &s2_out, unsigned long long &n_valid)
{
    using Real = typename W::Real;
    const unsigned long long base = chunk * G.chunk_units;
    const unsigned long long path_end = (base + G.chunk_units) * (unsigned long long)W::kUnitPaths;
    Real s = 0, s2 = 0;
    PackedSums packed;   // accumulates_packed<W> only
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
            W::eval(P, (uint32_t)base + (uint32_t)(k * kThreads) + tid, (uint32_t)(base >> 32), v, sh, job);
            if constexpr (accumulates_packed<W>::value) {
#pragma unroll
                for (int q = 0; q + 1 < W::kUnitPaths; q += 2)
                    packed.add(v[q], v[q + 1]);
            } else {
#pragma unroll
                for (int q = 0; q < W::kUnitPaths; q++)
                    add_value<W::kClampAtZero>(v[q], s, s2);
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
            W::eval(P, (uint32_t)base + (uint32_t)(k * kThreads) + tid, (uint32_t)(base >> 32), v, sh, job);
            if constexpr (accumulates_packed<W>::value) {
                // (a path beyond the total adds +0 to its running sums: exact)
#pragma unroll
                for (int q = 0; q + 1 < W::kUnitPaths; q += 2)
                    packed.add(unit * (unsigned long long)W::kUnitPaths + q < G.total_paths ? v[q] : Real(0),
                               unit * (unsigned long long)W::kUnitPaths + q + 1 < G.total_paths ? v[q + 1] : Real(0));
            } else {
#pragma unroll
                for (int q = 0; q < W::kUnitPaths; q++) {
                    if (unit * (unsigned long long)W::kUnitPaths + q < G.total_paths)
                        add_value<W::kClampAtZero>(v[q], s, s2);
                }
            }
        }
    }
    if constexpr (accumulates_packed<W>::value) {
        s = PackedSums::joined(packed.s);
        s2 = PackedSums::joined(packed.s2);
    }
    s_out = s;
    s2_out = s2;
}

// The pricing kernel skeleton.  W is a workload policy:
//   W::Real            float or double: the per-path arithmetic type
//   W::Params          by-value parameter block (constant bank)
//   W::kUnitPaths      paths served by one draw unit
//   W::kMinBlocks      CTAs per SM the register budget is sized for; W::kUnroll  unroll of the unit loop
//   W::Shared          per-CTA shared-memory state (the fp64 math tables; empty for fp32)
//   W::JobState        per-job shared-memory state of one sub-block (fp64: the exponent table of the job's scaled
//                      logarithm; empty for fp32), filled by W::prepare(P, job, tid)
//   W::kClampAtZero    the per-path value is max(v, 0) of what eval returns (add_value clamps for free)
//   W::eval(P, unit_lo, unit_hi, v, sh, job) fills v[kUnitPaths] with the per-path values of a draw unit;
//                      inside a chunk unit_hi is the same for every thread (chunks are aligned), so the
//                      part of the first two Philox rounds that depends only on it runs on the uniform datapath
// Persistent CTAs; every sub-block of 256 threads walks chunks of the shard (ChunkWalk) through run_chunk.
template <class W>
__global__ void __launch_bounds__(kThreads * W::kSubBlocks, W::kMinBlocks)
mc_accumulate_kernel(const __grid_constant__ typename W::Params P, const __grid_constant__ Geometry G,
                     unsigned long long *__restrict__ acc)
{
    using Real = typename W::Real;
    constexpr int kSub = W::kSubBlocks;
    pdl_launch_dependents();
    // W::Shared in dynamic shared memory (the replicated fp64 tables are 192 KB); nothing for fp32
    extern __shared__ __align__(16) unsigned char mcb_dynamic_smem[];
    typename W::Shared &sh = *reinterpret_cast<typename W::Shared *>(mcb_dynamic_smem);
    __shared__ BlockScratch scs[kSub];
    __shared__ typename W::JobState jobs[kSub];
    // sub-block = the "block" of the stream definition: kThreads threads, its own scratch, its own chunks
    const int sub = kSub == 1 ? 0 : (int)(threadIdx.x / kThreads);
    const int tid = kSub == 1 ? (int)threadIdx.x : (int)(threadIdx.x % kThreads);
    BlockScratch &sc = scs[sub];
    sh.load();
    W::prepare(P, jobs[sub], tid);
    if (tid < kAccWords)
        sc.acc[tid] = 0ull;
    __syncthreads();
    for (ChunkWalk<kSub> walk(sub); walk.live(G); walk.advance(sc)) {
        walk.claim_ahead(G, tid);
        Real s, s2;
        unsigned long long n_valid;
        run_chunk<W>(P, G, G.first_chunk + walk.chunk, tid, sh, jobs[sub], s, s2, n_valid);
        chunk_commit<kSub>((double)s, (double)s2, n_valid, G, sc, sub, tid, walk.ahead);
    }
    if constexpr (kSub > 1)
        __syncthreads();  // every sub-block has committed its last chunk before the CTA-wide tail
    finish(sc, acc, G, tid);
}

// ---- many jobs in ONE launch (sweeps) -------------------------------------------------------------
// The reference's cvaOpt driver prices 5 time grids x 4 thread counts as 20 blocking calls
// (double_precision/cvaOpt.cu:70-109).  Here the jobs of a sweep that share a kernel (workload and precision) are
// laid end to end in one chunk index space; a sub-block claims global chunks from one counter, finds the job a chunk
// belongs to (jobs are few: a linear scan of the offsets) and runs it exactly as the one-job kernel would -- same
// chunks, same reduction, same integer limbs, so every job's result is bit-identical to its one-call price.  A job's
// parameters are read from the kernel's parameter block through a job index (constant bank, one indexed load per
// use), its limbs go to its own 12 words of `acc`, and the launch's last CTA publishes all of them to the host.
constexpr int kBatchMaxJobs = 24;
template <class W>
struct BatchJobs {
    int n_jobs;
    int pad;
    LaunchCtl *ctl;
    unsigned long long *acc;              // n_jobs x 12 words of device scratch, zero between launches
    unsigned long long *host_slots;       // n_jobs x kPeerSlotWords words of mapped host memory (flagged halves)
    unsigned long long host_flag;
    unsigned int first[kBatchMaxJobs + 1];    // global index of each job's first chunk; first[n_jobs] = total
    unsigned short slot[kBatchMaxJobs];       // which host slot a job reports to
    JobGeometry geo[kBatchMaxJobs];
    typename W::Params params[kBatchMaxJobs];
};

template <class W>
__global__ void __launch_bounds__(kThreads * W::kSubBlocks, W::kMinBlocks)
mc_accumulate_batch_kernel(const __grid_constant__ BatchJobs<W> B)
{
    using Real = typename W::Real;
    constexpr int kSub = W::kSubBlocks;
    pdl_launch_dependents();
    extern __shared__ __align__(16) unsigned char mcb_dynamic_smem[];
    typename W::Shared &sh = *reinterpret_cast<typename W::Shared *>(mcb_dynamic_smem);
    __shared__ BlockScratch scs[kSub];
    __shared__ typename W::JobState jobs[kSub];
    __shared__ bool s_last;
    const int sub = kSub == 1 ? 0 : (int)(threadIdx.x / kThreads);
    const int tid = kSub == 1 ? (int)threadIdx.x : (int)(threadIdx.x % kThreads);
    BlockScratch &sc = scs[sub];
    sh.load();
    W::prepare(B.params[0], jobs[sub], tid);
    if (tid < kAccWords)
        sc.acc[tid] = 0ull;
    __syncthreads();
    const unsigned int total = B.first[B.n_jobs];
    unsigned int chunk = blockIdx.x * kSub + sub, ahead = 0u;
    int job = 0;
    while (chunk < total) {
        if (tid == 0)
            ahead = gridDim.x * kSub + atomicAdd(&B.ctl->next, 1u);
        // chunks are claimed in increasing order, so the job index only moves forward; when it moves, the limbs
        // collected so far belong to the previous job
        int j = job;
        while (chunk >= B.first[j + 1])
            j++;
        if (j != job) {
            // (thread t < 12 reads and clears its own word; thread 0's next write to sc.acc comes after a barrier)
            scratch_flush(sc, B.acc + (size_t)job * kAccWords, tid);
            if (tid < kAccWords)
                sc.acc[tid] = 0ull;
            job = j;
            // the sub-block's per-job state follows (everybody is past the previous chunk's last barrier)
            W::prepare(B.params[job], jobs[sub], tid);
            if constexpr (!std::is_empty<typename W::JobState>::value)
                sub_barrier<kSub>(sub);
        }
        Real s, s2;
        unsigned long long n_valid;
        run_chunk<W>(B.params[job], B.geo[job], (unsigned long long)(chunk - B.first[job]), tid, sh, jobs[sub], s, s2, n_valid);
        chunk_commit<kSub>((double)s, (double)s2, n_valid, B.geo[job], sc, sub, tid, ahead);
        chunk = sc.next;
    }
    pdl_wait();
    scratch_flush(sc, B.acc + (size_t)job * kAccWords, tid);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0)
        s_last = atomicAdd(&B.ctl->ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last)
        return;
    __threadfence();
    if (threadIdx.x == 0) {
        B.ctl->next = 0u;
        B.ctl->ticket = 0u;
    }
    // every job's 24 flagged halves to its host slot; the device scratch is left zeroed
    for (int i = threadIdx.x; i < B.n_jobs * kPeerHalves; i += blockDim.x) {
        const int jb = i / kPeerHalves, h = i % kPeerHalves;
        unsigned long long *word = B.acc + (size_t)jb * kAccWords + (h >> 1);
        const unsigned long long w = *((volatile unsigned long long *)word);
        st_relaxed_sys(B.host_slots + (size_t)B.slot[jb] * kPeerSlotWords + h, B.host_flag | ((h & 1) ? (w >> 32) : (w & 0xffffffffull)));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < B.n_jobs * kAccWords; i += blockDim.x)
        B.acc[i] = 0ull;
}

// Host side of a launch: dynamic shared memory = W::Shared (opt-in above 48 KB, set once per instantiation).
template <class W> constexpr size_t accumulate_smem_bytes() { return std::is_empty<typename W::Shared>::value ? 0 : sizeof(typename W::Shared); }
template <class W, class Kernel> inline cudaError_t accumulate_prepare(Kernel kernel, bool *done)
{
    // a function attribute is per device: set it once on each device this process prices on
    if (accumulate_smem_bytes<W>() <= 48 * 1024)
        return cudaSuccess;
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess || device < 0 || device >= 64)
        return e != cudaSuccess ? e : cudaErrorInvalidDevice;
    if (!done[device]) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)accumulate_smem_bytes<W>());
        if (e != cudaSuccess)
            return e;
        done[device] = true;
    }
    return cudaSuccess;
}
template <class W> inline cudaError_t accumulate_prepare()
{
    static bool done[64] = {};
    return accumulate_prepare<W>(mc_accumulate_kernel<W>, done);
}
template <class W> inline cudaError_t accumulate_batch_prepare()
{
    static bool done[64] = {};
    return accumulate_prepare<W>(mc_accumulate_batch_kernel<W>, done);
}

// launch with or without programmatic dependent launch (the attribute lets this kernel start while the previous one
// of the stream drains; see pdl_launch_dependents)
template <class... KArgs, class... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, bool overlap,
                                 Args &&...args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = overlap ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// per-launch options the engine passes down to the launchers
struct LaunchOptions {
    bool overlap = false;   // programmatic dependent launch
};

template <class W>
inline cudaError_t accumulate_launch(int grid, const typename W::Params &p, const Geometry &g, unsigned long long *d_acc,
                                     cudaStream_t stream, const LaunchOptions &opt = LaunchOptions())
{
    cudaError_t e = accumulate_prepare<W>();
    if (e != cudaSuccess)
        return e;
    return launch_kernel(mc_accumulate_kernel<W>, grid, kThreads * W::kSubBlocks, accumulate_smem_bytes<W>(), stream, opt.overlap,
                         p, g, d_acc);
}
template <class W>
inline cudaError_t accumulate_batch_launch(int grid, const BatchJobs<W> &b, cudaStream_t stream, const LaunchOptions &opt = LaunchOptions())
{
    cudaError_t e = accumulate_batch_prepare<W>();
    if (e != cudaSuccess)
        return e;
    return launch_kernel(mc_accumulate_batch_kernel<W>, grid, kThreads * W::kSubBlocks, accumulate_smem_bytes<W>(), stream, opt.overlap, b);
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

template <class W> inline int accumulate_batch_blocks_per_sm()
{
    int n = 0;
    if (accumulate_batch_prepare<W>() != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, mc_accumulate_batch_kernel<W>, kThreads * W::kSubBlocks,
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
    __shared__ typename W::JobState job;
    sh.load();
    W::prepare(P, job, (int)threadIdx.x);
    __syncthreads();
    for (unsigned long long i = blockIdx.x * (unsigned long long)kThreads + threadIdx.x; i < n_units;
         i += (unsigned long long)gridDim.x * kThreads) {
        typename W::Real v[W::kUnitPaths];
        W::eval(P, (uint32_t)(first_unit + i), (uint32_t)((first_unit + i) >> 32), v, sh, job);
#pragma unroll
        for (int q = 0; q < W::kUnitPaths; q++)
            out[i * W::kUnitPaths + q] = (W::kClampAtZero && sign_bit_set(v[q])) ? (typename W::Real)0 : v[q];
    }
}

}  // namespace mcb
