// launch.h -- host-callable launchers of the pricing kernels (one translation unit per workload,
// each owning its __constant__ table).  Internal to libmcb200; the public ABI is include/mcb200.h.
#pragma once

#include "device_common.cuh"

namespace mcb {

enum : uint32_t { kTagVanilla = 1u, kTagBasket = 2u, kTagCva = 3u };

// ---- many jobs of one kernel in one launch (device_common.cuh: mc_accumulate_batch_kernel) ----
struct BatchShape {
    int n_jobs;
    JobGeometry geo[kBatchMaxJobs];
    unsigned int n_chunks[kBatchMaxJobs];
    unsigned short slot[kBatchMaxJobs];   // index of the job's host slot
};
struct BatchTarget {
    LaunchCtl *ctl;
    unsigned long long *d_acc;         // kBatchMaxJobs x 12 words of zeroed device scratch
    unsigned long long *host_slots;    // device alias of the mapped host slots
    unsigned long long host_flag;
};
template <class B>
inline void fill_batch_header(B &b, const BatchShape &shape, const BatchTarget &target)
{
    b.n_jobs = shape.n_jobs;
    b.ctl = target.ctl;
    b.acc = target.d_acc;
    b.host_slots = target.host_slots;
    b.host_flag = target.host_flag;
    unsigned int first = 0;
    for (int i = 0; i < shape.n_jobs; i++) {
        b.first[i] = first;
        b.geo[i] = shape.geo[i];
        b.slot[i] = shape.slot[i];
        first += shape.n_chunks[i];
    }
    for (int i = shape.n_jobs; i <= kBatchMaxJobs; i++)
        b.first[i] = first;
}

// ---- vanilla (DP/MonteCarloKernel.cu:67-71, :179-220) ----
// a = ln S0 + (r - v^2/2) T, b = v sqrt(T) in natural-log units; the launcher rescales for the kernel (kernels_vanilla.cu)
struct VanillaJob {
    PhiloxKeys keys;
    double a, b, k;
};
int vanilla_blocks_per_sm(int precision);
cudaError_t vanilla_launch(int precision, const VanillaJob &job, const Geometry &geom, int grid,
                           unsigned long long *d_acc, cudaStream_t stream, const LaunchOptions &opt);
int vanilla_batch_blocks_per_sm(int precision);
cudaError_t vanilla_batch_launch(int precision, const BatchShape &shape, const VanillaJob *jobs, int grid,
                                 const BatchTarget &target, cudaStream_t stream, const LaunchOptions &opt);
cudaError_t vanilla_paths(int precision, const VanillaJob &job, unsigned long long first_unit,
                          unsigned long long n_units, void *d_out, cudaStream_t stream);

// ---- basket (DP/MonteCarloKernel.cu:74-101, :133-177) ----
// all arrays in natural-log units and fp64; the launcher narrows / rescales for the kernel
struct BasketJob {
    PhiloxKeys keys;
    int n;                 // assets
    bool full;             // factor has entries above the diagonal
    const double *factor;  // row-major n x n, already scaled: v_i sqrt(T) L_ij
    const double *a;       // (r - v_i^2/2) T + v_i sqrt(T) d_i
    const double *m;       // w_i s_i
    double k;
};
int basket_padded_width(int n);  // register-template width that serves n assets, 0 beyond the widest one (64)
int basket_max_width();           // widest basket any route prices (the wide route: local-memory normals, device-memory factor)
int basket_blocks_per_sm(int precision, int n, bool full);
// which kernel serves wide fp32 baskets (32 < n <= 64): 0 = tensor cores (tcgen05, basket_tc.cuh), 1 = FFMA2 only.
// Process-wide; first read falls back to the environment variable MCB200_BASKET_ENGINE (0 / 1).
int basket_engine_get();
void basket_engine_set(int engine);
bool basket_uses_tensor_cores(int precision, int n);
cudaError_t basket_launch(int precision, const BasketJob &job, const Geometry &geom, int grid,
                          unsigned long long *d_acc, cudaStream_t stream, const LaunchOptions &opt);
cudaError_t basket_paths(int precision, const BasketJob &job, unsigned long long first_unit,
                         unsigned long long n_units, void *d_out, cudaStream_t stream);

// ---- CVA (DP/MonteCarloKernel.cu:104-129, :222-283) ----
constexpr int kCvaMaxDates = 1024;   // kept dates the constant table holds; longer grids are read from device memory
struct CvaDateHost {
    double w, inv, c1, sig, kd;
};
struct CvaJob {
    PhiloxKeys keys;
    double y0, mu_dt, sig_dt, k;
    int n_dates;               // kept dates
    const CvaDateHost *dates;  // n_dates entries
};
int cva_blocks_per_sm(int precision, int n_dates);
cudaError_t cva_launch(int precision, const CvaJob &job, const Geometry &geom, int grid,
                       unsigned long long *d_acc, cudaStream_t stream, const LaunchOptions &opt);
// jobs[i].dates are concatenated into the device table (kCvaMaxDates entries in all)
int cva_batch_blocks_per_sm(int precision);
cudaError_t cva_batch_launch(int precision, const BatchShape &shape, const CvaJob *jobs, int grid,
                             const BatchTarget &target, cudaStream_t stream, const LaunchOptions &opt);
cudaError_t cva_paths(int precision, const CvaJob &job, unsigned long long first_unit,
                      unsigned long long n_units, void *d_out, cudaStream_t stream);

// ---- instrumentation (kernels_debug.cu) ----
cudaError_t debug_philox(unsigned long long n, const uint32_t *d_ctr, PhiloxKeys keys, uint32_t *d_out,
                         cudaStream_t stream);
cudaError_t debug_normals(int precision, unsigned long long n, const uint32_t *d_ctr, PhiloxKeys keys,
                          void *d_out, cudaStream_t stream);
cudaError_t debug_math64(int fn, unsigned long long n, const double *d_in, double *d_out, cudaStream_t stream);
cudaError_t debug_reduce(const double *d_values, unsigned long long n_valid, int unit_paths, int rounds,
                         bool accumulate_in_float, int scale_exp_sum, int scale_exp_sumsq,
                         unsigned long long *d_acc, cudaStream_t stream);


}  // namespace mcb
