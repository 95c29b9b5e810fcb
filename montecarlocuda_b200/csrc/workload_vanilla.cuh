// workload_vanilla.cuh -- the European-call policy of mc_accumulate_kernel (see kernels_vanilla.cu).
#pragma once

#include <type_traits>

#include "device_common.cuh"
#include "device_math.cuh"

namespace mcb {

constexpr uint32_t kVanillaTag = 1u;

#ifndef MCB_VANILLA_SHORT_SQRT
#define MCB_VANILLA_SHORT_SQRT true   // 5-instruction square root (device_math64.cuh); false: 7 (A/B switch, tools/build_variant.sh)
#endif

// tuned on B200 (profiles/r01_tune_vanilla.txt): CTAs per SM / unroll of the unit loop
template <typename Real> struct VanillaTuning;
template <> struct VanillaTuning<float> { static constexpr int kMinBlocks = 8, kUnroll = 1; };
// fp64: sub-blocks of 256 threads per CTA around one replicated table set (192 KB).  2 sub-blocks (128 registers
// available, 74 used) against 3 (80 available): 8.38 vs 9.36 ms without a spill, 10.16 ms with the round's first
// kernel, which spilled (profiles/r02k_ab_experiments.txt) -- the kernel lives on instruction-level parallelism, not
// on warps.  Unroll 2: 8.56 vs 8.42 ms.
#ifndef MCB_VANILLA_UNROLL
#define MCB_VANILLA_UNROLL 1
#endif
#ifndef MCB_VANILLA_SUBBLOCKS
#define MCB_VANILLA_SUBBLOCKS 2
#endif
template <> struct VanillaTuning<double> { static constexpr int kMinBlocks = MCB_VANILLA_SUBBLOCKS, kUnroll = MCB_VANILLA_UNROLL; };

// kAccumLayout: instantiated by mc_accumulate_kernel (bank-conflict-free fp64 tables shared by kSubBlocks x 256
// threads per CTA) rather than by the per-path / instrumentation kernels (plain tables, 256 threads)
template <typename RealT, int kMinBlocksT = VanillaTuning<RealT>::kMinBlocks, int kUnrollT = VanillaTuning<RealT>::kUnroll,
          bool kAccumLayout = false>
struct Vanilla {
    using Real = RealT;
    static constexpr int kUnitPaths = NormalsPerBlock<RealT>::value;
    // fp64: the kMinBlocksT CTAs of 256 threads an SM holds become ONE CTA of kMinBlocksT sub-blocks around one table set
    static constexpr int kSubBlocks = (kAccumLayout && sizeof(RealT) == 8) ? kMinBlocksT : 1;
    static constexpr int kMinBlocks = kSubBlocks > 1 ? 1 : kMinBlocksT;
    static constexpr int kUnroll = kUnrollT;
    // a and b = v sqrt(T) in the units the exponential is cheapest in (exp_scaled: log2 units for fp32, units of
    // ln2/256 for fp64); b rides under the Box-Muller square root
    struct Params {
        PhiloxKeys keys;
        Real a, k;
        PolarScale<Real> scale;
    };
    using Shared = std::conditional_t<kAccumLayout, typename SharedAccumFor<Real>::type, typename SharedFor<Real>::type>;
    using JobState = typename JobStateFor<Real>::type;
    // fp64: eval returns S_T - K and add_value clamps (sign test + predicated accumulation); fp32: one FMNMX here
    static constexpr bool kClampAtZero = sizeof(Real) == 8;
    static __device__ __forceinline__ void prepare(const Params &P, JobState &job, int tid) { prepare_polar(P.scale, job, tid); }
    static __device__ __forceinline__ void eval(const Params &P, uint32_t unit_lo, uint32_t unit_hi,
                                                Real (&v)[kUnitPaths], const Shared &sh, const JobState &job)
    {
        uint32_t w[4];
        philox4x32_10(unit_lo, unit_hi, 0u, kVanillaTag, P.keys, w);
        // the exponent a + b z needs b r, never z itself (device_math.cuh, polar_from_words): a path costs one FMA
        // and one exponential after its pair's (b r, cos, sin)
        constexpr int kPairs = kUnitPaths / 2;
        Real br[kPairs], cs[kPairs], sn[kPairs];
        polar_from_words<MCB_VANILLA_SHORT_SQRT>(w, br, cs, sn, sh, P.scale, job);
#pragma unroll
        for (int i = 0; i < kPairs; i++) {
            const Real s0 = exp_scaled(fma(br[i], cs[i], P.a), sh), s1 = exp_scaled(fma(br[i], sn[i], P.a), sh);
            if constexpr (sizeof(Real) == 4) {
                add_f32x2(s0, s1, -P.k, v[2 * i], v[2 * i + 1]);     // one FADD2 for the pair's two S_T - K
                v[2 * i] = positive_part(v[2 * i]);
                v[2 * i + 1] = positive_part(v[2 * i + 1]);
            } else {
                v[2 * i] = s0 - P.k;          // add_value clamps
                v[2 * i + 1] = s1 - P.k;
            }
        }
    }
};


}  // namespace mcb
