// workload_vanilla.cuh -- the European-call policy of mc_accumulate_kernel (see kernels_vanilla.cu).
#pragma once

#include <type_traits>

#include "device_common.cuh"
#include "device_math.cuh"

namespace mcb {

constexpr uint32_t kVanillaTag = 1u;

// tuned on B200 (profiles/r01_tune_vanilla.txt): CTAs per SM / unroll of the unit loop
template <typename Real> struct VanillaTuning;
template <> struct VanillaTuning<float> { static constexpr int kMinBlocks = 8, kUnroll = 1; };
// fp64: sub-blocks of 256 threads per CTA around one replicated table set (96 KB).  3 sub-blocks leave 80 registers
// per thread and the kernel spills (11.5 ms); 2 leave 128 (94 used): 9.70 ms (profiles/r01k_tune_vanilla.txt)
template <> struct VanillaTuning<double> { static constexpr int kMinBlocks = 2, kUnroll = 1; };

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
    struct Params {
        PhiloxKeys keys;
        Real a, k;
        PolarScale<Real> scale;  // of b = v sqrt(T) (in the exponent's units), folded under the Box-Muller square root
    };
    using Shared = std::conditional_t<kAccumLayout, typename SharedAccumFor<Real>::type, typename SharedFor<Real>::type>;
    template <class Sh> static __device__ __forceinline__ Real grow(Real x, const Sh &sh)
    {
        if constexpr (sizeof(Real) == 4)
            return mufu_ex2(x);
        else
            return exp_tab(x, sh.t);
    }
    static __device__ __forceinline__ void eval(const Params &P, uint32_t unit_lo, uint32_t unit_hi,
                                                Real (&v)[kUnitPaths], const Shared &sh)
    {
        uint32_t w[4];
        philox4x32_10(unit_lo, unit_hi, 0u, kVanillaTag, P.keys, w);
        // the exponent a + b z needs b r, never z itself (device_math.cuh, polar_from_words): a path costs one FMA
        // and one exponential after its pair's (b r, cos, sin)
        constexpr int kPairs = kUnitPaths / 2;
        Real br[kPairs], cs[kPairs], sn[kPairs];
        polar_from_words(w, br, cs, sn, sh, P.scale);
#pragma unroll
        for (int i = 0; i < kPairs; i++) {
            v[2 * i] = positive_part(grow(fma(br[i], cs[i], P.a), sh) - P.k);
            v[2 * i + 1] = positive_part(grow(fma(br[i], sn[i], P.a), sh) - P.k);
        }
    }
};


}  // namespace mcb
