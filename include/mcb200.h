/*
 * mcb200.h -- extended C ABI of the mcb200 Monte Carlo pricing engine (libmcb200.so).
 *
 * Plain C: opaque handle, pointers and sizes, no CUDA or torch types in any signature (a stream
 * is passed as void*).  Every entry point states the reference interface it replaces
 * (DP/ = /root/reference/double_precision/).  The three reference entry points themselves
 * (dev_vanillaOpt, dev_basketOpt, dev_cvaEquityOption) are declared in MonteCarlo.h and live in
 * libmcb200_dp.so / libmcb200_sp.so, thin shims over this API.
 *
 * What the reference API cannot express and this one adds: 64-bit path counts,
 * an explicit seed, a runtime basket width, a status code instead of exit(1), a discounted
 * standard error next to the reference's `Confidence`, a persistent context (the reference
 * allocates, seeds XORWOW and frees on every call, DP/MonteCarloKernel.cu:296-363), and path-range
 * shards whose partial results are exact integers, so that any number of GPUs gives the same bits.
 *
 * There is no CPU fallback: without a usable CUDA device every compute call fails with
 * MCB200_ERR_NO_DEVICE / MCB200_ERR_CUDA.
 */
#ifndef MCB200_H_
#define MCB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCB200_VERSION 100

/* status codes (0 = success) */
enum {
    MCB200_OK = 0,
    MCB200_ERR_INVALID = 1,     /* bad argument (null pointer, n <= 0, non-finite parameter ...) */
    MCB200_ERR_CUDA = 2,        /* a CUDA runtime call failed; see mcb200_last_error() */
    MCB200_ERR_NO_DEVICE = 3,   /* no CUDA device / device index out of range */
    MCB200_ERR_OVERFLOW = 4,    /* a partial sum left the 160-bit fixed-point window, or was NaN */
    MCB200_ERR_UNSUPPORTED = 5, /* e.g. basket wider than MCB200_MAX_ASSETS, grid longer than MCB200_MAX_DATES */
    MCB200_ERR_ALIGNMENT = 6,   /* shard boundary not on a chunk boundary */
    MCB200_ERR_PEER_TIMEOUT = 7 /* a peer rank's partial sums did not arrive in time (fused combine), or were pulled too late */
};

enum { MCB200_F32 = 0, MCB200_F64 = 1 };
enum { MCB200_VANILLA = 1, MCB200_BASKET = 2, MCB200_CVA = 3 };

#define MCB200_MAX_ASSETS 256 /* up to 64 assets: accumulators in registers (tensor cores for wide fp32); beyond: a generic route */
#define MCB200_MAX_DATES (1 << 20) /* exposure dates of a CVA grid (up to 1024 kept dates sit in constant memory, longer grids in device memory) */
#define MCB200_LANES 5
/* accumulator block: [0..4] sum limbs, [5..9] sum-of-squares limbs, [10] paths counted,
 * [11] error flags.  Each limb carries a 32-bit payload in a 64-bit word, so blocks from
 * different chunks, launches or GPUs are combined by plain integer addition (an int64 SUM
 * all-reduce) in any order. */
#define MCB200_ACC_WORDS 12

typedef struct mcb200_ctx mcb200_ctx;

/* mirrors OptionData (DP/MonteCarlo.h:32-38); always double here, the kernels narrow it */
typedef struct {
    double s, k, r, v, t;
} mcb200_option_t;

/* runtime-width MultiOptionData (DP/MonteCarlo.h:41-50): p is row-major n x n and holds the
 * Cholesky factor, exactly what the reference expects on entry (DP/basketOpt.cu:96-99). */
typedef struct {
    int n;
    const double *s, *v, *p, *d, *w;
    double k, t, r;
} mcb200_basket_t;

/* mirrors CVA (DP/MonteCarlo.h:57-65).  grid_mode 0 = the reference's time grid (remaining time
 * by repeated subtraction in the working precision, last date kept or dropped by its rounding,
 * SURVEY.md 2.4 Q3); 1 = exact grid, every date kept, exposure at maturity = intrinsic value. */
typedef struct {
    double def_int, lgd;
    mcb200_option_t option;
    int n_dates;
    int grid_mode;
} mcb200_cva_t;

typedef struct {
    uint64_t n_paths;    /* paths simulated */
    double sum, sumsq;   /* exact sums of the per-path value and its square, rounded once */
    double mean;         /* sum / n (undiscounted) */
    double expected;     /* reference `Expected`: e^{-rT} * mean (pricing) or mean (CVA) */
    double confidence;   /* reference `Confidence`: 1.96 * s / sqrt(n) of the UNdiscounted value */
    double std_error;    /* standard error of `expected` (discounted where expected is) */
    double kernel_ms;    /* device time of the pricing kernel(s), CUDA events, max over devices; 0 unless MCB200_OPT_TIMING */
} mcb200_result_t;

/* Everything a shard launch and the final combine must agree on; a pure function of the job
 * (never of the GPU count), filled by mcb200_plan_*(). */
typedef struct {
    int workload, precision;
    uint64_t total_paths;
    int unit_paths;        /* paths served by one Philox draw unit */
    int rounds;            /* units per thread per chunk */
    uint64_t total_units, chunk_units, n_chunks;
    int scale_exp_sum;     /* fixed-point scale 2^e applied before the integer split */
    int scale_exp_sumsq;
    double discount;       /* e^{-rT} for pricing, 1 for CVA */
} mcb200_plan_t;

/* ---- context ---- */
int mcb200_device_count(void);
int mcb200_create(mcb200_ctx **out, int device);
int mcb200_destroy(mcb200_ctx *ctx);
int mcb200_device(const mcb200_ctx *ctx);
int mcb200_sm_count(const mcb200_ctx *ctx);
const char *mcb200_strerror(int status);
const char *mcb200_last_error(const mcb200_ctx *ctx);
/* kernels launched by this context since creation (bench.py's gpu_launches) */
uint64_t mcb200_launch_count(const mcb200_ctx *ctx);
/* Context options.
 *   MCB200_OPT_TIMING   (default 0)  record CUDA events around the kernels of the blocking calls and report
 *                                    mcb200_result_t.kernel_ms (two more stream operations per call; 0 otherwise).
 *                                    The reference times every call (DP/MonteCarloKernel.cu:380-386, :447-453).
 *   MCB200_OPT_OVERLAP  (default 0; environment MCB200_OVERLAP=1)  launch with programmatic dependent launch:
 *                                    consecutive launches of a stream overlap the tail of one with the start of the
 *                                    next (independent jobs only share read-only state). */
enum { MCB200_OPT_TIMING = 1, MCB200_OPT_OVERLAP = 2 };
int mcb200_set_option(mcb200_ctx *ctx, int option, int value);
int mcb200_get_option(const mcb200_ctx *ctx, int option);
/* Which kernel prices a wide single-precision basket (32 < n <= 64 assets):
 *   MCB200_BASKET_TENSOR (default)  the correlated draw L*g of brownianVect (DP/MonteCarloKernel.cu:74-87) runs on the
 *                                   tcgen05 tensor cores as a 3xTF32 128-path tile product (csrc/basket_tc.cuh);
 *   MCB200_BASKET_FFMA              the in-register packed-FMA column sweep used for every other width.
 * Both consume the same Philox stream; they round the mat-vec differently, so their prices agree to fp32
 * accuracy per path, not bit for bit.  Each is bit-reproducible across grid shapes and GPU counts.
 * Process-wide; before the first call the environment variable MCB200_BASKET_ENGINE (0 / 1) decides. */
enum { MCB200_BASKET_TENSOR = 0, MCB200_BASKET_FFMA = 1 };
int mcb200_set_basket_engine(int engine);
int mcb200_get_basket_engine(void);

/* ---- one-call pricing: host structs in, host result out, synchronous ----
 * replaces dev_vanillaOpt / dev_basketOpt / dev_cvaEquityOption
 * (DP/MonteCarloKernel.cu:500, :483, :517) including their host-side closing (:412-423, :459-469) */
int mcb200_vanilla(mcb200_ctx *ctx, int precision, const mcb200_option_t *opt, uint64_t n_paths,
                   uint64_t seed, mcb200_result_t *out);
int mcb200_basket(mcb200_ctx *ctx, int precision, const mcb200_basket_t *opt, uint64_t n_paths,
                  uint64_t seed, mcb200_result_t *out);
int mcb200_cva(mcb200_ctx *ctx, int precision, const mcb200_cva_t *cva, uint64_t n_paths,
               uint64_t seed, mcb200_result_t *out);
/* the same over several devices of this process: contiguous chunk ranges, one launch per device,
 * partials combined by exact integer addition on the host */
int mcb200_vanilla_multi(mcb200_ctx **ctxs, int n_ctx, int precision, const mcb200_option_t *opt,
                         uint64_t n_paths, uint64_t seed, mcb200_result_t *out);
int mcb200_basket_multi(mcb200_ctx **ctxs, int n_ctx, int precision, const mcb200_basket_t *opt,
                        uint64_t n_paths, uint64_t seed, mcb200_result_t *out);
int mcb200_cva_multi(mcb200_ctx **ctxs, int n_ctx, int precision, const mcb200_cva_t *cva,
                     uint64_t n_paths, uint64_t seed, mcb200_result_t *out);

/* ---- batched pricing: many jobs, one synchronisation ----
 * replaces the serial sweep of the reference's cvaOpt driver (double_precision/cvaOpt.cu:70-109: one
 * blocking dev_cvaEquityOption call per grid size and thread count).  params points to the
 * mcb200_option_t / mcb200_basket_t / mcb200_cva_t of the job's workload.  out[i] is exactly what the
 * one-call entry point returns for job i (kernel_ms = device time of the whole batch); status_out
 * (optional, n_jobs entries) receives the per-job status, the return value the first failure. */
typedef struct {
    int workload, precision;
    const void *params;
    uint64_t n_paths, seed;
} mcb200_job_t;
int mcb200_price_batch(mcb200_ctx *ctx, int n_jobs, const mcb200_job_t *jobs, mcb200_result_t *out,
                       int *status_out);

/* ---- sharded pricing (one process per GPU; the combine is the caller's all-reduce) ----
 * plan -> shard range for (rank, world) -> asynchronous launch accumulating into a DEVICE
 * accumulator block (MCB200_ACC_WORDS zero-initialised 64-bit words owned by the caller) on the
 * given stream (a cudaStream_t passed as void*; NULL is CUDA's default stream, as everywhere) ->
 * [int64 SUM all-reduce of the block] -> mcb200_finalize on a host copy. */
int mcb200_plan_vanilla(int precision, const mcb200_option_t *opt, uint64_t n_paths, mcb200_plan_t *plan);
int mcb200_plan_basket(int precision, const mcb200_basket_t *opt, uint64_t n_paths, mcb200_plan_t *plan);
int mcb200_plan_cva(int precision, const mcb200_cva_t *cva, uint64_t n_paths, mcb200_plan_t *plan);
int mcb200_shard_range(const mcb200_plan_t *plan, int rank, int world, uint64_t *first_chunk,
                       uint64_t *n_chunks);
int mcb200_vanilla_launch(mcb200_ctx *ctx, const mcb200_plan_t *plan, const mcb200_option_t *opt,
                          uint64_t seed, uint64_t first_chunk, uint64_t n_chunks,
                          uint64_t *d_acc, void *stream);
int mcb200_basket_launch(mcb200_ctx *ctx, const mcb200_plan_t *plan, const mcb200_basket_t *opt,
                         uint64_t seed, uint64_t first_chunk, uint64_t n_chunks,
                         uint64_t *d_acc, void *stream);
int mcb200_cva_launch(mcb200_ctx *ctx, const mcb200_plan_t *plan, const mcb200_cva_t *cva,
                      uint64_t seed, uint64_t first_chunk, uint64_t n_chunks,
                      uint64_t *d_acc, void *stream);
/* host-side closing on a (summed) accumulator block: DP/MonteCarloKernel.cu:412-423, :459-469 */
/* ---- peer groups: the (sum, sum^2) combine across GPUs fused into the pricing kernel ----
 * The reference has no multi-GPU path; the collective this replaces is the single 96-byte SUM all-reduce that
 * follows a sharded launch.  Every rank owns a mailbox in device memory that its peers write over NVLink
 * (CUDA IPC between processes, peer access inside one).  While a connected group is attached to a context, the
 * *_launch entry points above end with the combine: the last CTA of each device's kernel pushes the device's
 * integer limbs into every mailbox, waits for its peers' and adds them, so d_acc holds the JOB's totals on every
 * rank when the kernel ends (bit-identical everywhere; no separate collective, no extra launch).  Every rank must
 * issue the same sequence of sharded launches, on streams that can run concurrently with its peers'.
 *   one process per GPU:   peer_create on every rank -> exchange the 64-byte handles (any transport) -> peer_connect
 *   one process, many GPUs (or many contexts on one GPU): peer_create for every rank -> peer_connect_local */
#define MCB200_PEER_HANDLE_BYTES 64
#define MCB200_PEER_MAX 8
typedef struct mcb200_peer mcb200_peer;
int mcb200_peer_create(mcb200_ctx *ctx, int rank, int world, mcb200_peer **out,
                       unsigned char handle[MCB200_PEER_HANDLE_BYTES]);
int mcb200_peer_connect(mcb200_peer *peer, const unsigned char *handles /* world x 64 bytes, rank-major */);
int mcb200_peer_connect_local(mcb200_peer **group, int world);
int mcb200_peer_attach(mcb200_ctx *ctx, mcb200_peer *peer /* NULL detaches */);
/* Split phase.  MCB200_PEER_WAIT (default): as above, the kernel is the collective.  MCB200_PEER_PUSH: the last CTA only
 * pushes the device's limbs into the mailboxes and the kernel ends -- no rank waits for the slowest one, back-to-back
 * jobs do not lock-step; d_acc of the launch is not written.  mcb200_peer_pull then enqueues one small kernel that
 * sums the group's MOST RECENT launch out of this rank's mailbox into d_acc (12 words, overwritten).  A mailbox keeps
 * the last 8 launches; a pull that comes later than that, or a peer that does not answer within the timeout
 * (default 10 s, environment MCB200_PEER_TIMEOUT_MS), ends in MCB200_ERR_PEER_TIMEOUT from mcb200_finalize. */
enum { MCB200_PEER_WAIT = 0, MCB200_PEER_PUSH = 1 };
int mcb200_peer_set_mode(mcb200_peer *peer, int mode);
int mcb200_peer_set_timeout_ms(mcb200_peer *peer, double ms);
int mcb200_peer_pull(mcb200_peer *peer, uint64_t *d_acc, void *stream);
int mcb200_peer_destroy(mcb200_peer *peer);

int mcb200_finalize(const mcb200_plan_t *plan, const uint64_t acc[MCB200_ACC_WORDS],
                    mcb200_result_t *out);

/* ---- per-path values (parity instrumentation; same device code as the pricing kernels) ----
 * out_host receives n_paths values of the working precision (float or double): the undiscounted
 * payoff (vanilla, basket) or the path CVA.  first_path must be a multiple of the draw-unit size
 * (vanilla: 6 in single precision, 4 in double; 1 otherwise). */
int mcb200_vanilla_paths(mcb200_ctx *ctx, int precision, const mcb200_option_t *opt, uint64_t seed,
                         uint64_t first_path, uint64_t n_paths, void *out_host);
int mcb200_basket_paths(mcb200_ctx *ctx, int precision, const mcb200_basket_t *opt, uint64_t seed,
                        uint64_t first_path, uint64_t n_paths, void *out_host);
int mcb200_cva_paths(mcb200_ctx *ctx, int precision, const mcb200_cva_t *cva, uint64_t seed,
                     uint64_t first_path, uint64_t n_paths, void *out_host);
/* raw generator output: n counters (4 words each) under one key -> 4 words each, and the
 * normals made from them (6 floats or 4 doubles per counter: one Philox block is three single-precision
 * or two double-precision Box-Muller pairs) */
int mcb200_debug_philox(mcb200_ctx *ctx, uint64_t n, const uint32_t *ctr_host, const uint32_t key[2],
                        uint32_t *out_host);
int mcb200_debug_normals(mcb200_ctx *ctx, int precision, uint64_t n, const uint32_t *ctr_host,
                         const uint32_t key[2], void *out_host);
/* the hand-built fp64 special functions of the kernels, element-wise on n host doubles:
 * fn 0 = -2 ln(u), 1 = sqrt, 2 = 1/x, 3 = e^x, 4 = cos and sin of 2 pi k / 2^52 (the input's bit
 * pattern is the 52-bit integer k), 5 = the same for a 20-bit k (two-level table; k = low word of the
 * input's bit pattern), 6 = sqrt by the short iteration the pricing kernels use, 7 = 2^(y/256) (the exponential
 * with its argument in table units, as the pricing kernels call it), 8 = its second variant (table entry last).
 * out_host receives 2 doubles per element (second = sin for fn 4, 5). */
int mcb200_debug_math64(mcb200_ctx *ctx, int fn, uint64_t n, const double *in_host, double *out_host);
/* reduce ONE chunk of given per-path values with the pricing kernels' block reduction and
 * integer split; acc_host receives the accumulator block */
int mcb200_debug_reduce(mcb200_ctx *ctx, const double *values_host, uint64_t n_valid, int unit_paths,
                        int rounds, int accumulate_in_float, int scale_exp_sum, int scale_exp_sumsq,
                        uint64_t acc_host[MCB200_ACC_WORDS]);

#ifdef __cplusplus
}
#endif
#endif /* MCB200_H_ */
