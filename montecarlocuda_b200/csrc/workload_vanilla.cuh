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
        Real a, b, k;
    };
    using Shared = typename SharedFor<Real>::type;
    static __device__ __forceinline__ float grow(float x, const NoShared &) { return mufu_ex2(x); }
    static __device__ __forceinline__ double grow(double x, const SharedTables64 &sh) { return exp_tab(x, sh.t); }
    static __device__ __forceinline__ void eval(const Params &P, uint32_t unit_lo, uint32_t unit_hi,
                                                Real (&v)[kUnitPaths], const Shared &sh)
    {
        uint32_t w[4];
        philox4x32_10(unit_lo, unit_hi, 0u, kVanillaTag, P.keys, w);
        Real z[kUnitPaths];
        normals_from_words(w, z, sh);
#pragma unroll
        for (int q = 0; q < kUnitPaths; q++) {
            const Real st = grow(fma(P.b, z[q], P.a), sh);
            v[q] = positive_part(st - P.k);
        }
    }
};


}  // namespace mcb
