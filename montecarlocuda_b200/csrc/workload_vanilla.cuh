// workload_vanilla.cuh -- the European-call policy of mc_accumulate_kernel (see kernels_vanilla.cu).
#pragma once

#include "device_common.cuh"
#include "device_math.cuh"

namespace mcb {

constexpr uint32_t kVanillaTag = 1u;

// tuned on B200 (profiles/r01_tune_vanilla.txt): CTAs per SM / unroll of the unit loop
template <typename Real> struct VanillaTuning;
template <> struct VanillaTuning<float> { static constexpr int kMinBlocks = 8, kUnroll = 1; };
template <> struct VanillaTuning<double> { static constexpr int kMinBlocks = 3, kUnroll = 1; };

template <typename RealT, int kMinBlocksT = VanillaTuning<RealT>::kMinBlocks, int kUnrollT = VanillaTuning<RealT>::kUnroll>
struct Vanilla {
    using Real = RealT;
    static constexpr int kUnitPaths = NormalsPerBlock<RealT>::value;
    static constexpr int kMinBlocks = kMinBlocksT;
    static constexpr int kUnroll = kUnrollT;
    struct Params {
        PhiloxKeys keys;
        Real a, k;
        PolarScale<Real> scale;  // of b = v sqrt(T) (in the exponent's units), folded under the Box-Muller square root
    };
    using Shared = typename SharedFor<Real>::type;
    static __device__ __forceinline__ float grow(float x, const NoShared &) { return mufu_ex2(x); }
    static __device__ __forceinline__ double grow(double x, const SharedTables64 &sh) { return exp_tab(x, sh.t); }
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
