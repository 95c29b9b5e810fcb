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
        Real c;  // fp32 only: -2 ln2 * b^2, the diffusion scale folded under the Box-Muller square root
    };
    using Shared = typename SharedFor<Real>::type;
    static __device__ __forceinline__ float grow(float x, const NoShared &) { return mufu_ex2(x); }
    static __device__ __forceinline__ double grow(double x, const SharedTables64 &sh) { return exp_tab(x, sh.t); }
    static __device__ __forceinline__ void eval(const Params &P, uint32_t unit_lo, uint32_t unit_hi,
                                                Real (&v)[kUnitPaths], const Shared &sh)
    {
        uint32_t w[4];
        philox4x32_10(unit_lo, unit_hi, 0u, kVanillaTag, P.keys, w);
        if constexpr (sizeof(Real) == 4) {
            // the exponent a + b z needs b r, never z itself: b r = sqrt(lg2(u) * (-2 ln2 b^2)), so the scale rides
            // on the constant under the square root and a pair costs 9 instructions (5 FP32 + 4 MUFU) before its 2^x
            float f[6];
            uniforms_f32(w, f);
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const float br = mufu_sqrt(mufu_lg2(2.0f - f[2 * i]) * P.c);
                const float ang = fmaf(f[2 * i + 1], 6.283185307179586f, -9.42477796076938f);
                v[2 * i] = positive_part(mufu_ex2(fmaf(br, mufu_cos(ang), P.a)) - P.k);
                v[2 * i + 1] = positive_part(mufu_ex2(fmaf(br, mufu_sin(ang), P.a)) - P.k);
            }
        } else {
            Real z[kUnitPaths];
            normals_from_words(w, z, sh);
#pragma unroll
            for (int q = 0; q < kUnitPaths; q++) {
                const Real st = grow(fma(P.b, z[q], P.a), sh);
                v[q] = positive_part(st - P.k);
            }
        }
    }
};


}  // namespace mcb
