// engine.cu -- host side of libmcb200: persistent context, job planning, host-built parameter
// tables, shard launches, the exact-integer combine and the closing formulas.
//
// Replaces the host orchestration of the reference GPU engine: MonteCarlo_init / MonteCarlo /
// cvaMonteCarlo / MonteCarlo_closing and the three extern "C" wrappers
// (DP/MonteCarloKernel.cu:296-532).  The reference allocates device + pinned memory, seeds
// 65 536 XORWOW states, copies parameters into __constant__ symbols, prints timing lines and
// frees everything on EVERY call; here a context owns a stream, a ring of self-cleaning launch control blocks
// and a few result slots in mapped host memory for its whole life, the generator is stateless, and a blocking
// call is ONE kernel launch: no memset in front of it (the launch's last CTA leaves the control block zeroed),
// no copy and no stream synchronisation behind it (that CTA writes the 96-byte result into the host slot as
// flagged words and the host polls the flags), no events unless the caller asked for kernel times.
#include "../../include/mcb200.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#endif

#include "launch.h"

using namespace mcb;

constexpr int kCtlRing = 8;     // launch control blocks: a launch may still be draining when the next ones start
constexpr int kHostRing = 4;    // result slots of the blocking calls

struct mcb200_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    unsigned long long *d_acc = nullptr;  // debug_reduce
    unsigned long long *h_acc = nullptr;  // pinned
    LaunchCtl *d_ctl = nullptr;                  // kCtlRing blocks, zero between launches
    unsigned long long *d_batch_acc = nullptr;   // kCtlRing x kBatchMaxJobs x 12 words, zero between launches
    unsigned long long *h_slots = nullptr;       // mapped pinned: kHostRing result slots of kPeerSlotWords words
    unsigned long long *d_slots = nullptr;       // ... as the device addresses them
    // result slots of a batch (mcb200_price_batch): grown on demand, kept for the context's life
    unsigned long long *h_batch_slots = nullptr, *d_batch_slots = nullptr;
    size_t batch_capacity = 0;  // in jobs
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    std::string last_error;
    uint64_t launches = 0;
    uint64_t results = 0;       // blocking results delivered through a host slot (flag = peer_flag(results))
    bool timing = false;        // record CUDA events around the kernels of the blocking calls (kernel_ms)
    bool overlap = false;       // programmatic dependent launch: consecutive launches of a stream overlap their tails
    cudaStream_t last_stream = nullptr;
    bool launched = false;
    std::mutex mu;
    mcb200_peer *peer = nullptr;  // attached peer group: sharded launches combine inside the kernel
};

// One rank's membership of a peer group (see PeerLink, device_common.cuh).
struct mcb200_peer {
    mcb200_ctx *ctx = nullptr;
    int rank = 0, world = 1;
    int mode = kPeerWait;
    unsigned long long timeout_ns = 10ull * 1000 * 1000 * 1000;
    unsigned long long *mailbox = nullptr;          // this rank's mailbox (device memory of ctx->device)
    unsigned long long *mail[kPeerMax] = {};        // every rank's mailbox as addressed from this device
    bool opened[kPeerMax] = {};                     // mail[r] came from cudaIpcOpenMemHandle
    bool connected = false;
    unsigned long long seq = 0;                     // fused launches so far
};

// The second phase of a kPeerPush launch, when somebody wants the totals: one small CTA sums the `world` sources of
// launch `seq` out of this device's mailbox into out[12] (device memory: overwritten, not added to).
__global__ void __launch_bounds__(kThreads)
peer_pull_kernel(const unsigned long long *mailbox, int world, unsigned long long seq, unsigned long long timeout_ns,
                 unsigned long long *out)
{
    __shared__ unsigned int s_bad;
    __shared__ unsigned int s_half[kPeerMax * kPeerHalves];
    const int tid = threadIdx.x;
    if (tid == 0)
        s_bad = 0u;
    __syncthreads();
    if (tid < kPeerHalves * world) {
        unsigned int v;
        if (peer_poll(mailbox, seq, tid / kPeerHalves, tid % kPeerHalves, timeout_ns, v) != 0)
            atomicAdd(&s_bad, 1u);
        s_half[tid] = v;
    }
    __syncthreads();
    if (tid < kAccWords) {
        unsigned long long total = 0ull;
        for (int src = 0; src < world; src++)
            total += (unsigned long long)s_half[src * kPeerHalves + 2 * tid] |
                     ((unsigned long long)s_half[src * kPeerHalves + 2 * tid + 1] << 32);
        if (tid == kAccWords - 1 && s_bad)
            total += kErrPeerTimeout;
        out[tid] = total;
    }
}

namespace {

class DeviceGuard {
public:
    explicit DeviceGuard(int device)
    {
        if (cudaGetDevice(&previous_) != cudaSuccess)
            previous_ = -1;
        status_ = cudaSetDevice(device);
    }
    ~DeviceGuard()
    {
        if (previous_ >= 0)
            cudaSetDevice(previous_);
    }
    cudaError_t status() const { return status_; }

private:
    int previous_ = -1;
    cudaError_t status_ = cudaSuccess;
};

int fail_cuda(mcb200_ctx *ctx, cudaError_t e, const char *what)
{
    if (ctx) {
        ctx->last_error = std::string(what) + ": " + cudaGetErrorString(e);
    }
    cudaGetLastError();  // clear the sticky-free error state
    return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice)
               ? MCB200_ERR_NO_DEVICE
               : MCB200_ERR_CUDA;
}

int fail(mcb200_ctx *ctx, int status, const char *what)
{
    if (ctx)
        ctx->last_error = what;
    return status;
}

#define MCB_CUDA(ctx, call)                                  \
    do {                                                     \
        cudaError_t _e = (call);                             \
        if (_e != cudaSuccess)                               \
            return fail_cuda((ctx), _e, #call);              \
    } while (0)

bool finite_all(std::initializer_list<double> xs)
{
    for (double x : xs)
        if (!std::isfinite(x))
            return false;
    return true;
}

PhiloxKeys make_keys(uint64_t seed)
{
    // key schedule of Philox4x32-10: round i uses key + i * (W0, W1)
    PhiloxKeys k;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int i = 0; i < 10; i++) {
        k.k0[i] = k0;
        k.k1[i] = k1;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return k;
}

// units per thread per chunk: a pure function of the job size (never of the GPU count), aiming
// at >= 2^16 chunks for large jobs and at most 64 units per thread
int chunk_rounds(uint64_t total_units)
{
#ifndef MCB_EXP_ROUNDS_SHIFT
#define MCB_EXP_ROUNDS_SHIFT 0   // experiment only (tools/build_variant.sh): longer chunks, to size the cost of a commit
#endif
    const uint64_t q = total_units >> (24 - MCB_EXP_ROUNDS_SHIFT);
    int r = 1;
    while (r < (64 << MCB_EXP_ROUNDS_SHIFT) && (uint64_t)(r * 2) <= q)
        r *= 2;
    return r;
}

int fill_plan(int workload, int precision, uint64_t n_paths, double scale_ref, double discount,
              mcb200_plan_t *plan)
{
    if (!plan || n_paths == 0 || (precision != MCB200_F32 && precision != MCB200_F64))
        return MCB200_ERR_INVALID;
    if (!(scale_ref > 0) || !std::isfinite(scale_ref) || !std::isfinite(discount))
        return MCB200_ERR_INVALID;
    std::memset(plan, 0, sizeof *plan);
    plan->workload = workload;
    plan->precision = precision;
    plan->total_paths = n_paths;
    // one Philox block = six fp32 or four fp64 normals (device_math.cuh)
    plan->unit_paths = workload == MCB200_VANILLA ? (precision == MCB200_F32 ? 6 : 4) : 1;
    plan->total_units = (n_paths + plan->unit_paths - 1) / plan->unit_paths;
    plan->rounds = chunk_rounds(plan->total_units);
    plan->chunk_units = (uint64_t)kThreads * plan->rounds;
    plan->n_chunks = (plan->total_units + plan->chunk_units - 1) / plan->chunk_units;
    // fixed-point window: least significant bit 2^-80 of the job's money scale (a power of two
    // near max(spot, strike)), 160 bits wide
    const int e = std::ilogb(scale_ref) + 1;
    plan->scale_exp_sum = 80 - e;
    plan->scale_exp_sumsq = 80 - 2 * e;
    plan->discount = discount;
    return MCB200_OK;
}

JobGeometry job_geometry(const mcb200_plan_t &plan)
{
    JobGeometry g{};
    g.total_paths = plan.total_paths;
    g.chunk_units = plan.chunk_units;
    g.rounds = plan.rounds;
    g.scale_exp_sum = plan.scale_exp_sum;
    g.scale_exp_sumsq = plan.scale_exp_sumsq;
    return g;
}

Geometry make_geometry(const mcb200_plan_t &plan, uint64_t first_chunk, uint64_t n_chunks)
{
    Geometry g{};
    static_cast<JobGeometry &>(g) = job_geometry(plan);
    g.first_chunk = first_chunk;
    g.n_chunks = n_chunks;
    return g;
}

int check_range(const mcb200_plan_t *plan, uint64_t first_chunk, uint64_t n_chunks)
{
    if (!plan || first_chunk > plan->n_chunks || n_chunks > plan->n_chunks - first_chunk)
        return MCB200_ERR_ALIGNMENT;
    return MCB200_OK;
}

// ---- host-built jobs --------------------------------------------------------------------------

int make_vanilla_job(int precision, const mcb200_option_t *o, uint64_t seed, VanillaJob *job)
{
    if (!o || !finite_all({o->s, o->k, o->r, o->v, o->t}) || !(o->s > 0) || o->v < 0 || o->t < 0)
        return MCB200_ERR_INVALID;
    // S_T = S0 exp((r - v^2/2) T + v sqrt(T) z)   (DP/MonteCarloKernel.cu:69)
    (void)precision;   // natural-log units here; the launcher rescales for the kernel's exponential
    job->keys = make_keys(seed);
    job->a = std::log(o->s) + (o->r - 0.5 * o->v * o->v) * o->t;
    job->b = o->v * std::sqrt(o->t);
    job->k = o->k;
    // the table-driven exponential takes |exponent| <= 700 in natural-log units (normals reach |z| < 8.6)
    if (!(std::fabs(job->a) + 9.0 * std::fabs(job->b) < 700.0))
        return MCB200_ERR_INVALID;
    return MCB200_OK;
}

struct BasketTables {
    std::vector<double> factor, a, m;
    BasketJob job;
};

int make_basket_job(const mcb200_basket_t *o, uint64_t seed, BasketTables *t)
{
    if (!o || !o->s || !o->v || !o->p || !o->d || !o->w)
        return MCB200_ERR_INVALID;
    const int n = o->n;
    if (n < 1)
        return MCB200_ERR_INVALID;
    if (n > MCB200_MAX_ASSETS || n > basket_max_width())
        return MCB200_ERR_UNSUPPORTED;
    if (!finite_all({o->k, o->t, o->r}) || o->t < 0)
        return MCB200_ERR_INVALID;
    t->factor.assign((size_t)n * n, 0.0);
    t->a.assign(n, 0.0);
    t->m.assign(n, 0.0);
    bool full = false;
    const double sqrt_t = std::sqrt(o->t);
    for (int i = 0; i < n; i++) {
        if (!finite_all({o->s[i], o->v[i], o->d[i], o->w[i]}))
            return MCB200_ERR_INVALID;
        // S_i = s_i exp((r - v_i^2/2) T + v_i sqrt(T) (sum_j p_ij g_j + d_i))  (DP/MonteCarloKernel.cu:79-93)
        for (int j = 0; j < n; j++) {
            const double p = o->p[(size_t)i * n + j];
            if (!std::isfinite(p))
                return MCB200_ERR_INVALID;
            if (j > i && p != 0.0)
                full = true;
            t->factor[(size_t)i * n + j] = o->v[i] * sqrt_t * p;
        }
        t->a[i] = (o->r - 0.5 * o->v[i] * o->v[i]) * o->t + o->v[i] * sqrt_t * o->d[i];
        t->m[i] = o->w[i] * o->s[i];
        double reach = std::fabs(t->a[i]);
        for (int j = 0; j < n; j++)
            reach += 9.0 * std::fabs(t->factor[(size_t)i * n + j]);
        if (!(reach < 700.0))  // range of the table-driven exponential
            return MCB200_ERR_INVALID;
    }
    t->job.keys = make_keys(seed);
    t->job.n = n;
    t->job.full = full;
    t->job.factor = t->factor.data();
    t->job.a = t->a.data();
    t->job.m = t->m.data();
    t->job.k = o->k;
    return MCB200_OK;
}

struct CvaTables {
    std::vector<CvaDateHost> dates;
    CvaJob job;
};

// Remaining time at each exposure date.  grid_mode 0 follows the reference literally: dt = T / n
// and t -= dt in the WORKING precision, a date contributes when the rounded t is >= 0
// (DP/MonteCarloKernel.cu:233,249).
template <typename Real>
void reference_grid(double T, int n, std::vector<double> &tau, std::vector<int> &keep, double &dt_out)
{
    const Real dt = (Real)T / (Real)n;
    Real t = (Real)T;
    for (int j = 0; j < n; j++) {
        t -= dt;
        tau[j] = (double)t;
        keep[j] = t >= 0 ? 1 : 0;
    }
    dt_out = (double)dt;
}

int make_cva_job(int precision, const mcb200_cva_t *c, uint64_t seed, CvaTables *t)
{
    if (!c)
        return MCB200_ERR_INVALID;
    const mcb200_option_t &o = c->option;
    const int n = c->n_dates;
    if (!finite_all({o.s, o.k, o.r, o.v, o.t, c->def_int, c->lgd}) || !(o.s > 0) || !(o.k > 0) || !(o.v > 0) ||
        !(o.t > 0) || n < 1)
        return MCB200_ERR_INVALID;
    if (n > MCB200_MAX_DATES)
        return MCB200_ERR_UNSUPPORTED;
    std::vector<double> tau(n);
    std::vector<int> keep(n);
    double dt;
    if (c->grid_mode == 0) {
        if (precision == MCB200_F32)
            reference_grid<float>(o.t, n, tau, keep, dt);
        else
            reference_grid<double>(o.t, n, tau, keep, dt);
    } else {
        dt = o.t / n;
        for (int j = 0; j < n; j++) {
            tau[j] = o.t * (double)(n - 1 - j) / (double)n;
            keep[j] = 1;
        }
    }
    int kept = 0;
    while (kept < n && keep[kept])
        kept++;
    const double huge = precision == MCB200_F32 ? 1e30 : 1e300;
    t->dates.assign(kept > 0 ? kept : 1, CvaDateHost{});
    for (int j = 0; j < kept; j++) {
        CvaDateHost &d = t->dates[j];
        // dp_j = e^{-lambda t_{j-1}} - e^{-lambda t_j}   (DP/MonteCarloKernel.cu:248), times LGD (:259)
        d.w = c->lgd * (std::exp(-(dt * j) * c->def_int) - std::exp(-(dt * (j + 1)) * c->def_int));
        if (tau[j] > 0) {
            // d1 = (ln(s/K) + (r + v^2/2) tau) / (v sqrt(tau)), d2 = d1 - v sqrt(tau)   (:126-127)
            const double sig = o.v * std::sqrt(tau[j]);
            d.inv = 1.0 / sig;
            d.c1 = (o.r + 0.5 * o.v * o.v) * tau[j] * d.inv;
            d.sig = sig;
            d.kd = o.k * std::exp(-o.r * tau[j]);
        } else {
            // tau == 0: the reference's formulas degenerate to the intrinsic value
            // (log(s/K)/0 = +-inf, cnd -> 0 or 1); a huge finite slope does the same without NaNs
            d.inv = huge;
            d.c1 = 0;
            d.sig = 0;
            d.kd = o.k;
        }
    }
    // range of the table-driven exponential: |ln(s/K)| + |drift| T + 9 standard deviations of the log-spot at maturity
    // (the sum of n steps, not n times one step's reach: that rejected valid high-volatility jobs with many dates)
    if (!(std::fabs(std::log(o.s / o.k)) + std::fabs((o.r - 0.5 * o.v * o.v) * o.t) + 9.0 * o.v * std::sqrt(o.t) < 700.0))
        return MCB200_ERR_INVALID;
    t->job.keys = make_keys(seed);
    t->job.y0 = std::log(o.s / o.k);
    t->job.mu_dt = (o.r - 0.5 * o.v * o.v) * dt;  // geomBrownian, DP/MonteCarloKernel.cu:104-107
    t->job.sig_dt = o.v * std::sqrt(dt);
    t->job.k = o.k;
    t->job.n_dates = kept;
    t->job.dates = t->dates.data();
    return MCB200_OK;
}

double basket_scale(const mcb200_basket_t *o)
{
    double ref = std::fabs(o->k);
    double gross = 0;
    for (int i = 0; i < o->n; i++)
        gross += std::fabs(o->w[i] * o->s[i]);
    return gross > ref ? gross : ref;
}

// ---- launching ---------------------------------------------------------------------------------

struct AnyJob {
    int workload = 0;
    VanillaJob vanilla;
    BasketTables basket;
    CvaTables cva;
};

int blocks_per_sm(const mcb200_plan_t &plan, const AnyJob &job)
{
    switch (plan.workload) {
        case MCB200_VANILLA: return vanilla_blocks_per_sm(plan.precision);
        case MCB200_BASKET: return basket_blocks_per_sm(plan.precision, job.basket.job.n, job.basket.job.full);
        default: return cva_blocks_per_sm(plan.precision, job.cva.job.n_dates);
    }
}

// where a launch's result goes
struct Sink {
    unsigned long long *d_acc = nullptr;      // device accumulator block the launch adds its totals to, or nullptr
    unsigned long long *host_slot = nullptr;  // device address of a mapped host slot, or nullptr
    unsigned long long host_flag = 0;
    bool fused = false;                       // run the attached peer group's exchange in the kernel's tail
};

// a context's launches normally arrive on one stream; when the caller switches streams, the control blocks of
// launches still in flight on the old one must not be reused under them
void order_streams(mcb200_ctx *ctx, cudaStream_t stream)
{
    if (ctx->launched && ctx->last_stream != stream) {
        if (cudaStreamSynchronize(ctx->last_stream) != cudaSuccess)
            cudaGetLastError();  // the old stream is gone: nothing of it can still be running
    }
    ctx->last_stream = stream;
    ctx->launched = true;
}

// enqueue one shard on `stream` (device already current)
int enqueue(mcb200_ctx *ctx, const mcb200_plan_t &plan, const AnyJob &job, uint64_t first_chunk,
            uint64_t n_chunks, const Sink &sink, cudaStream_t stream)
{
    mcb200_peer *peer = sink.fused ? ctx->peer : nullptr;
    if (n_chunks == 0 && !peer && !sink.host_slot)
        return MCB200_OK;
    int per_sm = blocks_per_sm(plan, job);
    if (per_sm < 1)
        return fail(ctx, MCB200_ERR_CUDA, "kernel cannot be resident on this device (occupancy 0)");
    order_streams(ctx, stream);
    // a launch claims chunks through a 32-bit counter: longer shards go out as several launches (their totals add up:
    // integer limbs); only the last one carries the exchange / the host slot
    constexpr uint64_t kMaxChunks = 1ull << 30;
    if ((peer || sink.host_slot) && n_chunks > kMaxChunks)
        return fail(ctx, MCB200_ERR_UNSUPPORTED, "more than 2^30 chunks in one shard of a blocking or fused launch");
    do {
        const uint64_t part = n_chunks > kMaxChunks ? kMaxChunks : n_chunks;
        const bool final_part = part == n_chunks;
        uint64_t grid = (uint64_t)ctx->sm_count * per_sm;
        if (grid > part)
            grid = part;
        if (grid == 0)
            grid = 1;  // an empty shard still takes part in the fused combine / still answers its host slot
        Geometry g = make_geometry(plan, first_chunk, part);
        g.ctl = ctx->d_ctl + (ctx->launches % kCtlRing);
        if (final_part) {
            g.host_slot = sink.host_slot;
            g.host_flag = sink.host_flag;
        }
        if (peer && final_part) {
            g.peer.world = peer->world;
            g.peer.rank = peer->rank;
            g.peer.mode = peer->mode;
            g.peer.seq = peer->seq + 1;
            g.peer.timeout_ns = peer->timeout_ns;
            for (int r = 0; r < peer->world; r++)
                g.peer.mail[r] = peer->mail[r];
        }
        LaunchOptions opt;
        opt.overlap = ctx->overlap;
        cudaError_t e;
        switch (plan.workload) {
            case MCB200_VANILLA: e = vanilla_launch(plan.precision, job.vanilla, g, (int)grid, sink.d_acc, stream, opt); break;
            case MCB200_BASKET: e = basket_launch(plan.precision, job.basket.job, g, (int)grid, sink.d_acc, stream, opt); break;
            default: e = cva_launch(plan.precision, job.cva.job, g, (int)grid, sink.d_acc, stream, opt); break;
        }
        if (e != cudaSuccess)
            return fail_cuda(ctx, e, "kernel launch");
        ctx->launches++;
        if (peer && final_part)
            peer->seq++;  // only a launch that went out counts: a failed one must not desynchronise the ranks
        first_chunk += part;
        n_chunks -= part;
    } while (n_chunks > 0);
    return MCB200_OK;
}

inline void cpu_relax()
{
#if defined(__x86_64__) || defined(__i386__)
    _mm_pause();
#endif
}

// Wait for the 24 flagged halves of a host slot (written by a launch's last CTA over PCIe) and assemble the 12 words.
// No stream synchronisation: the slot IS the completion signal.  The stream is only queried now and then, so that a
// launch that died ends as an error instead of an endless poll.
int wait_slot(mcb200_ctx *ctx, cudaStream_t stream, const unsigned long long *slot, unsigned long long flag,
              uint64_t acc[MCB200_ACC_WORDS])
{
    const volatile unsigned long long *w = slot;
    bool drained = false;
    for (uint64_t spin = 1;; spin++) {
        int h = 0;
        unsigned long long v[kPeerHalves];
        for (; h < kPeerHalves; h++) {
            v[h] = w[h];
            if ((v[h] & 0xffffffff00000000ull) != flag)
                break;
        }
        if (h == kPeerHalves) {
            for (int i = 0; i < MCB200_ACC_WORDS; i++)
                acc[i] = (v[2 * i] & 0xffffffffull) | ((v[2 * i + 1] & 0xffffffffull) << 32);
            return MCB200_OK;
        }
        if (drained)
            return fail(ctx, MCB200_ERR_CUDA, "the launch finished without publishing its result");
        if ((spin & 0x3fff) == 0) {
            const cudaError_t e = cudaStreamQuery(stream);
            if (e == cudaSuccess)
                drained = true;  // everything has run: one more look at the slot, then give up
            else if (e != cudaErrorNotReady)
                return fail_cuda(ctx, e, "pricing kernel");
        }
        cpu_relax();
    }
}

// one-call pricing over n_ctx devices of this process
int price(mcb200_ctx **ctxs, int n_ctx, const mcb200_plan_t &plan, const AnyJob &job, mcb200_result_t *out)
{
    if (!ctxs || n_ctx < 1 || !out)
        return MCB200_ERR_INVALID;
    for (int i = 0; i < n_ctx; i++) {
        if (!ctxs[i])
            return MCB200_ERR_INVALID;
        for (int j = 0; j < i; j++)
            if (ctxs[j] == ctxs[i])
                return fail(ctxs[i], MCB200_ERR_INVALID, "the same context appears twice in the list");
    }
    // lock in one canonical order (by address), whatever order the caller listed the contexts in: two calls with
    // overlapping lists cannot deadlock each other
    std::vector<mcb200_ctx *> order(ctxs, ctxs + n_ctx);
    std::sort(order.begin(), order.end());
    std::vector<std::unique_lock<std::mutex>> locks;
    for (mcb200_ctx *c : order)
        locks.emplace_back(c->mu);
    // enqueue everywhere first, then wait: the devices run concurrently
    int status = MCB200_OK, enqueued = 0;
    std::vector<unsigned long long> flags((size_t)n_ctx);
    std::vector<int> slots((size_t)n_ctx);
    for (int i = 0; i < n_ctx && status == MCB200_OK; i++) {
        mcb200_ctx *ctx = ctxs[i];
        DeviceGuard guard(ctx->device);
        uint64_t first = 0, count = 0;
        if (guard.status() != cudaSuccess)
            status = fail_cuda(ctx, guard.status(), "cudaSetDevice");
        if (status == MCB200_OK)
            status = mcb200_shard_range(&plan, i, n_ctx, &first, &count);
        if (status == MCB200_OK && ctx->timing && cudaEventRecord(ctx->ev_begin, ctx->stream) != cudaSuccess)
            status = fail_cuda(ctx, cudaGetLastError(), "cudaEventRecord");
        if (status == MCB200_OK) {
            const uint64_t seq = ++ctx->results;
            slots[i] = (int)(seq % kHostRing);
            flags[i] = peer_flag(seq);
            Sink sink;
            sink.host_slot = ctx->d_slots + (size_t)slots[i] * kPeerSlotWords;
            sink.host_flag = flags[i];
            status = enqueue(ctx, plan, job, first, count, sink, ctx->stream);
        }
        if (status == MCB200_OK && ctx->timing && cudaEventRecord(ctx->ev_end, ctx->stream) != cudaSuccess)
            status = fail_cuda(ctx, cudaGetLastError(), "cudaEventRecord");
        if (status == MCB200_OK)
            enqueued = i + 1;
    }
    uint64_t total[MCB200_ACC_WORDS] = {0};
    double kernel_ms = 0;
    for (int i = 0; i < enqueued; i++) {
        mcb200_ctx *ctx = ctxs[i];
        DeviceGuard guard(ctx->device);
        if (status != MCB200_OK) {
            cudaStreamSynchronize(ctx->stream);  // an error on another device: do not leave kernels in flight behind it
            continue;
        }
        uint64_t part[MCB200_ACC_WORDS];
        status = wait_slot(ctx, ctx->stream, ctx->h_slots + (size_t)slots[i] * kPeerSlotWords, flags[i], part);
        if (status != MCB200_OK)
            continue;
        if (ctx->timing) {
            float ms = 0;
            if (cudaEventSynchronize(ctx->ev_end) == cudaSuccess && cudaEventElapsedTime(&ms, ctx->ev_begin, ctx->ev_end) == cudaSuccess) {
                if (ms > kernel_ms)
                    kernel_ms = ms;
            } else {
                cudaGetLastError();
            }
        }
        for (int w = 0; w < MCB200_ACC_WORDS; w++)
            total[w] += part[w];  // exact: integer limbs with 31 bits of headroom
    }
    if (status != MCB200_OK)
        return status;
    int st = mcb200_finalize(&plan, total, out);
    out->kernel_ms = kernel_ms;
    if (st != MCB200_OK)
        return fail(ctxs[0], st, st == MCB200_ERR_OVERFLOW ? "partial sum outside the fixed-point window or NaN"
                                                           : st == MCB200_ERR_PEER_TIMEOUT ? "a peer rank did not deliver its partial sums in time"
                                                                                            : "path count mismatch");
    return MCB200_OK;
}

// per-path values of [first_path, first_path + n_paths)
int paths(mcb200_ctx *ctx, const mcb200_plan_t &plan, const AnyJob &job, uint64_t first_path, uint64_t n_paths,
          void *out_host)
{
    if (!ctx || !out_host || n_paths == 0)
        return MCB200_ERR_INVALID;
    if (first_path % (uint64_t)plan.unit_paths)
        return fail(ctx, MCB200_ERR_ALIGNMENT, "first_path is not a multiple of the draw-unit size");
    std::lock_guard<std::mutex> lock(ctx->mu);
    DeviceGuard guard(ctx->device);
    MCB_CUDA(ctx, guard.status());
    const uint64_t first_unit = first_path / plan.unit_paths;
    const uint64_t n_units = (n_paths + plan.unit_paths - 1) / plan.unit_paths;
    const size_t elem = plan.precision == MCB200_F64 ? 8 : 4;
    void *d_out = nullptr;
    MCB_CUDA(ctx, cudaMalloc(&d_out, n_units * plan.unit_paths * elem));
    cudaError_t e;
    switch (plan.workload) {
        case MCB200_VANILLA: e = vanilla_paths(plan.precision, job.vanilla, first_unit, n_units, d_out, ctx->stream); break;
        case MCB200_BASKET: e = basket_paths(plan.precision, job.basket.job, first_unit, n_units, d_out, ctx->stream); break;
        default: e = cva_paths(plan.precision, job.cva.job, first_unit, n_units, d_out, ctx->stream); break;
    }
    if (e == cudaSuccess) {
        ctx->launches++;
        e = cudaMemcpyAsync(out_host, d_out, n_paths * elem, cudaMemcpyDeviceToHost, ctx->stream);
    }
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_out);
    if (e != cudaSuccess)
        return fail_cuda(ctx, e, "per-path kernel");
    return MCB200_OK;
}

}  // namespace

// ================================================================================================
// public C ABI
// ================================================================================================
extern "C" {

int mcb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int mcb200_create(mcb200_ctx **out, int device)
{
    if (!out)
        return MCB200_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return MCB200_ERR_NO_DEVICE;  // no CPU fallback: the engine is CUDA or nothing
    }
    if (device < 0 || device >= n)
        return MCB200_ERR_NO_DEVICE;
    mcb200_ctx *ctx = new (std::nothrow) mcb200_ctx;
    if (!ctx)
        return MCB200_ERR_INVALID;
    ctx->device = device;
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    const size_t ctl_bytes = sizeof(LaunchCtl) * kCtlRing;
    const size_t batch_acc_bytes = sizeof(unsigned long long) * kAccWords * kBatchMaxJobs * kCtlRing;
    const size_t slot_bytes = sizeof(unsigned long long) * kPeerSlotWords * kHostRing;
    bool ok = guard.status() == cudaSuccess && cudaGetDeviceProperties(&prop, device) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaMalloc(&ctx->d_acc, sizeof(unsigned long long) * kAccWords) == cudaSuccess &&
              cudaMallocHost(&ctx->h_acc, sizeof(unsigned long long) * kAccWords) == cudaSuccess &&
              cudaMalloc(&ctx->d_ctl, ctl_bytes) == cudaSuccess && cudaMemset(ctx->d_ctl, 0, ctl_bytes) == cudaSuccess &&
              cudaMalloc(&ctx->d_batch_acc, batch_acc_bytes) == cudaSuccess &&
              cudaMemset(ctx->d_batch_acc, 0, batch_acc_bytes) == cudaSuccess &&
              cudaHostAlloc(&ctx->h_slots, slot_bytes, cudaHostAllocMapped) == cudaSuccess &&
              cudaHostGetDevicePointer(&ctx->d_slots, ctx->h_slots, 0) == cudaSuccess &&
              cudaEventCreate(&ctx->ev_begin) == cudaSuccess && cudaEventCreate(&ctx->ev_end) == cudaSuccess &&
              cudaDeviceSynchronize() == cudaSuccess;
    if (!ok) {
        cudaGetLastError();
        mcb200_destroy(ctx);
        return MCB200_ERR_CUDA;
    }
    std::memset(ctx->h_slots, 0, slot_bytes);
    ctx->sm_count = prop.multiProcessorCount;
    const char *env = std::getenv("MCB200_OVERLAP");
    ctx->overlap = env && env[0] == '1';
    *out = ctx;
    return MCB200_OK;
}

int mcb200_destroy(mcb200_ctx *ctx)
{
    if (!ctx)
        return MCB200_OK;
    {
        DeviceGuard guard(ctx->device);
        if (ctx->stream)
            cudaStreamSynchronize(ctx->stream);
        if (ctx->launched && ctx->last_stream != ctx->stream && cudaStreamSynchronize(ctx->last_stream) != cudaSuccess)
            cudaGetLastError();
        if (ctx->ev_begin)
            cudaEventDestroy(ctx->ev_begin);
        if (ctx->ev_end)
            cudaEventDestroy(ctx->ev_end);
        if (ctx->d_acc)
            cudaFree(ctx->d_acc);
        if (ctx->h_acc)
            cudaFreeHost(ctx->h_acc);
        if (ctx->d_ctl)
            cudaFree(ctx->d_ctl);
        if (ctx->d_batch_acc)
            cudaFree(ctx->d_batch_acc);
        if (ctx->h_slots)
            cudaFreeHost(ctx->h_slots);
        if (ctx->h_batch_slots)
            cudaFreeHost(ctx->h_batch_slots);
        if (ctx->stream)
            cudaStreamDestroy(ctx->stream);
        cudaGetLastError();
    }
    delete ctx;
    return MCB200_OK;
}

int mcb200_device(const mcb200_ctx *ctx) { return ctx ? ctx->device : -1; }
int mcb200_sm_count(const mcb200_ctx *ctx) { return ctx ? ctx->sm_count : 0; }
uint64_t mcb200_launch_count(const mcb200_ctx *ctx) { return ctx ? ctx->launches : 0; }
const char *mcb200_last_error(const mcb200_ctx *ctx) { return ctx ? ctx->last_error.c_str() : ""; }

int mcb200_set_option(mcb200_ctx *ctx, int option, int value)
{
    if (!ctx)
        return MCB200_ERR_INVALID;
    std::lock_guard<std::mutex> lock(ctx->mu);
    switch (option) {
        case MCB200_OPT_TIMING: ctx->timing = value != 0; return MCB200_OK;
        case MCB200_OPT_OVERLAP: ctx->overlap = value != 0; return MCB200_OK;
        default: return MCB200_ERR_INVALID;
    }
}

int mcb200_get_option(const mcb200_ctx *ctx, int option)
{
    if (!ctx)
        return -1;
    switch (option) {
        case MCB200_OPT_TIMING: return ctx->timing ? 1 : 0;
        case MCB200_OPT_OVERLAP: return ctx->overlap ? 1 : 0;
        default: return -1;
    }
}

int mcb200_set_basket_engine(int engine)
{
    if (engine != MCB200_BASKET_TENSOR && engine != MCB200_BASKET_FFMA)
        return MCB200_ERR_INVALID;
    basket_engine_set(engine);
    return MCB200_OK;
}
int mcb200_get_basket_engine(void) { return basket_engine_get(); }

// ---- peer groups: the cross-GPU combine fused into the pricing kernel ----
int mcb200_peer_create(mcb200_ctx *ctx, int rank, int world, mcb200_peer **out, unsigned char handle[MCB200_PEER_HANDLE_BYTES])
{
    static_assert(sizeof(cudaIpcMemHandle_t) == MCB200_PEER_HANDLE_BYTES, "IPC handle size");
    if (!ctx || !out || !handle || world < 1 || world > kPeerMax || rank < 0 || rank >= world)
        return MCB200_ERR_INVALID;
    *out = nullptr;
    std::lock_guard<std::mutex> lock(ctx->mu);
    DeviceGuard guard(ctx->device);
    MCB_CUDA(ctx, guard.status());
    mcb200_peer *p = new (std::nothrow) mcb200_peer;
    if (!p)
        return MCB200_ERR_INVALID;
    p->ctx = ctx;
    p->rank = rank;
    p->world = world;
    cudaError_t e = cudaMalloc(&p->mailbox, kPeerMailboxBytes);
    if (e == cudaSuccess)
        e = cudaMemset(p->mailbox, 0, kPeerMailboxBytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess)
        e = cudaIpcGetMemHandle(&h, p->mailbox);
    if (e == cudaSuccess)
        e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        if (p->mailbox) cudaFree(p->mailbox);
        delete p;
        return fail_cuda(ctx, e, "mcb200_peer_create");
    }
    const char *env = std::getenv("MCB200_PEER_TIMEOUT_MS");
    if (env && std::atof(env) > 0)
        p->timeout_ns = (unsigned long long)(std::atof(env) * 1e6);
    std::memcpy(handle, &h, sizeof h);
    p->mail[rank] = p->mailbox;
    *out = p;
    return MCB200_OK;
}

int mcb200_peer_connect(mcb200_peer *p, const unsigned char *handles)
{
    if (!p || !handles || p->connected)
        return MCB200_ERR_INVALID;
    mcb200_ctx *ctx = p->ctx;
    std::lock_guard<std::mutex> lock(ctx->mu);
    DeviceGuard guard(ctx->device);
    MCB_CUDA(ctx, guard.status());
    for (int r = 0; r < p->world; r++) {
        if (r == p->rank)
            continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + (size_t)r * MCB200_PEER_HANDLE_BYTES, sizeof h);
        void *ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess)
            return fail_cuda(ctx, e, "cudaIpcOpenMemHandle (peer mailbox)");
        p->mail[r] = (unsigned long long *)ptr;
        p->opened[r] = true;
    }
    p->connected = true;
    return MCB200_OK;
}

int mcb200_peer_connect_local(mcb200_peer **group, int world)
{
    // all ranks live in this process: mailboxes are addressed directly (same device, or peer access between devices)
    if (!group || world < 1 || world > kPeerMax)
        return MCB200_ERR_INVALID;
    for (int r = 0; r < world; r++)
        if (!group[r] || group[r]->world != world || group[r]->rank != r || group[r]->connected)
            return MCB200_ERR_INVALID;
    for (int r = 0; r < world; r++) {
        mcb200_peer *p = group[r];
        DeviceGuard guard(p->ctx->device);
        MCB_CUDA(p->ctx, guard.status());
        for (int q = 0; q < world; q++) {
            const int other = group[q]->ctx->device;
            if (other != p->ctx->device) {
                int can = 0;
                MCB_CUDA(p->ctx, cudaDeviceCanAccessPeer(&can, p->ctx->device, other));
                if (!can)
                    return fail(p->ctx, MCB200_ERR_UNSUPPORTED, "no peer access between the devices of the group");
                cudaError_t e = cudaDeviceEnablePeerAccess(other, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return fail_cuda(p->ctx, e, "cudaDeviceEnablePeerAccess");
                cudaGetLastError();
            }
            p->mail[q] = group[q]->mailbox;
        }
        p->connected = true;
    }
    return MCB200_OK;
}

int mcb200_peer_attach(mcb200_ctx *ctx, mcb200_peer *peer)
{
    if (!ctx || (peer && (peer->ctx != ctx || !peer->connected)))
        return MCB200_ERR_INVALID;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->peer = peer;
    return MCB200_OK;
}

int mcb200_peer_set_mode(mcb200_peer *p, int mode)
{
    if (!p || (mode != MCB200_PEER_WAIT && mode != MCB200_PEER_PUSH))
        return MCB200_ERR_INVALID;
    std::lock_guard<std::mutex> lock(p->ctx->mu);
    p->mode = mode == MCB200_PEER_PUSH ? kPeerPush : kPeerWait;
    return MCB200_OK;
}

int mcb200_peer_set_timeout_ms(mcb200_peer *p, double ms)
{
    if (!p || !(ms > 0))
        return MCB200_ERR_INVALID;
    std::lock_guard<std::mutex> lock(p->ctx->mu);
    p->timeout_ns = (unsigned long long)(ms * 1e6);
    return MCB200_OK;
}

int mcb200_peer_pull(mcb200_peer *p, uint64_t *d_acc, void *stream)
{
    if (!p || !d_acc || !p->connected)
        return MCB200_ERR_INVALID;
    mcb200_ctx *ctx = p->ctx;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (p->seq == 0)
        return fail(ctx, MCB200_ERR_INVALID, "mcb200_peer_pull before the group's first launch");
    DeviceGuard guard(ctx->device);
    MCB_CUDA(ctx, guard.status());
    peer_pull_kernel<<<1, kThreads, 0, (cudaStream_t)stream>>>(p->mailbox, p->world, p->seq, p->timeout_ns,
                                                              (unsigned long long *)d_acc);
    MCB_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    return MCB200_OK;
}

int mcb200_peer_destroy(mcb200_peer *p)
{
    if (!p)
        return MCB200_OK;
    mcb200_ctx *ctx = p->ctx;
    {
        std::lock_guard<std::mutex> lock(ctx->mu);
        if (ctx->peer == p)
            ctx->peer = nullptr;
        DeviceGuard guard(ctx->device);
        cudaDeviceSynchronize();
        for (int r = 0; r < p->world; r++)
            if (p->opened[r])
                cudaIpcCloseMemHandle(p->mail[r]);
        if (p->mailbox) cudaFree(p->mailbox);
        cudaGetLastError();
    }
    delete p;
    return MCB200_OK;
}

const char *mcb200_strerror(int status)
{
    switch (status) {
        case MCB200_OK: return "success";
        case MCB200_ERR_INVALID: return "invalid argument";
        case MCB200_ERR_CUDA: return "CUDA runtime error";
        case MCB200_ERR_NO_DEVICE: return "no usable CUDA device (mcb200 has no CPU fallback)";
        case MCB200_ERR_OVERFLOW: return "partial sum outside the fixed-point window, or NaN";
        case MCB200_ERR_UNSUPPORTED: return "unsupported size";
        case MCB200_ERR_ALIGNMENT: return "range not aligned to the job's chunk / draw-unit grid";
        case MCB200_ERR_PEER_TIMEOUT: return "a peer rank did not deliver its partial sums in time (or the pull came too late)";
        default: return "unknown status";
    }
}

// ---- planning ----
int mcb200_plan_vanilla(int precision, const mcb200_option_t *opt, uint64_t n_paths, mcb200_plan_t *plan)
{
    if (!opt || !finite_all({opt->s, opt->k, opt->r, opt->t}))
        return MCB200_ERR_INVALID;
    const double ref = std::fabs(opt->s) > std::fabs(opt->k) ? std::fabs(opt->s) : std::fabs(opt->k);
    // price = e^{-rT} mean   (DP/MonteCarloKernel.cu:420)
    return fill_plan(MCB200_VANILLA, precision, n_paths, ref, std::exp(-opt->r * opt->t), plan);
}

int mcb200_plan_basket(int precision, const mcb200_basket_t *opt, uint64_t n_paths, mcb200_plan_t *plan)
{
    if (!opt || !opt->s || !opt->w || opt->n < 1 || !finite_all({opt->k, opt->r, opt->t}))
        return MCB200_ERR_INVALID;
    return fill_plan(MCB200_BASKET, precision, n_paths, basket_scale(opt), std::exp(-opt->r * opt->t), plan);
}

int mcb200_plan_cva(int precision, const mcb200_cva_t *cva, uint64_t n_paths, mcb200_plan_t *plan)
{
    if (!cva || !finite_all({cva->option.s, cva->option.k}))
        return MCB200_ERR_INVALID;
    const double s = std::fabs(cva->option.s), k = std::fabs(cva->option.k);
    // the CVA is not discounted   (DP/MonteCarloKernel.cu:466)
    return fill_plan(MCB200_CVA, precision, n_paths, s > k ? s : k, 1.0, plan);
}

int mcb200_shard_range(const mcb200_plan_t *plan, int rank, int world, uint64_t *first_chunk, uint64_t *n_chunks)
{
    if (!plan || !first_chunk || !n_chunks || world < 1 || rank < 0 || rank >= world)
        return MCB200_ERR_INVALID;
    const unsigned __int128 n = plan->n_chunks;
    const uint64_t lo = (uint64_t)(n * (unsigned)rank / (unsigned)world);
    const uint64_t hi = (uint64_t)(n * (unsigned)(rank + 1) / (unsigned)world);
    *first_chunk = lo;
    *n_chunks = hi - lo;
    return MCB200_OK;
}

// ---- closing: DP/MonteCarloKernel.cu:412-423 (pricing) and :459-469 (CVA) ----
static long double lanes_value(const uint64_t *lanes, int scale_exp)
{
    uint64_t limb[MCB200_LANES + 1], carry = 0;
    for (int i = 0; i < MCB200_LANES; i++) {
        const uint64_t x = lanes[i] + carry;
        limb[i] = x & 0xffffffffu;
        carry = x >> 32;
    }
    limb[MCB200_LANES] = carry;
    long double acc = 0;
    for (int i = MCB200_LANES; i >= 0; i--)
        acc = acc * 4294967296.0L + (long double)limb[i];
    return ldexpl(acc, -scale_exp);
}

int mcb200_finalize(const mcb200_plan_t *plan, const uint64_t acc[MCB200_ACC_WORDS], mcb200_result_t *out)
{
    if (!plan || !acc || !out)
        return MCB200_ERR_INVALID;
    std::memset(out, 0, sizeof *out);
    const long double sum = lanes_value(acc, plan->scale_exp_sum);
    const long double sumsq = lanes_value(acc + MCB200_LANES, plan->scale_exp_sumsq);
    const long double n = (long double)plan->total_paths;
    out->n_paths = acc[10];
    out->sum = (double)sum;
    out->sumsq = (double)sumsq;
    out->mean = (double)(sum / n);
    out->expected = (double)((long double)plan->discount * (sum / n));
    // s^2 = (n sum(x^2) - sum(x)^2) / (n (n - 1)); Confidence = 1.96 s / sqrt(n) on the UNdiscounted value
    long double var = (n * sumsq - sum * sum) / (n * (n - 1.0L));
    if (var < 0)
        var = 0;
    const long double sd = sqrtl(var);
    out->confidence = (double)(1.96L * sd / sqrtl(n));
    out->std_error = (double)((long double)plan->discount * sd / sqrtl(n));
    if ((acc[11] >> 32) != 0)
        return MCB200_ERR_PEER_TIMEOUT;  // a fused launch gave up waiting for a peer: no rank may trust these totals
    if (acc[11] != 0)
        return MCB200_ERR_OVERFLOW;
    if (acc[10] != plan->total_paths)
        return MCB200_ERR_INVALID;
    return MCB200_OK;
}

// ---- sharded launches ----
int mcb200_vanilla_launch(mcb200_ctx *ctx, const mcb200_plan_t *plan, const mcb200_option_t *opt, uint64_t seed,
                          uint64_t first_chunk, uint64_t n_chunks, uint64_t *d_acc, void *stream)
{
    if (!ctx || !plan || !d_acc || plan->workload != MCB200_VANILLA)
        return MCB200_ERR_INVALID;
    int st = check_range(plan, first_chunk, n_chunks);
    if (st != MCB200_OK)
        return st;
    AnyJob job;
    st = make_vanilla_job(plan->precision, opt, seed, &job.vanilla);
    if (st != MCB200_OK)
        return st;
    std::lock_guard<std::mutex> lock(ctx->mu);
    DeviceGuard guard(ctx->device);
    MCB_CUDA(ctx, guard.status());
    Sink sink;
    sink.d_acc = (unsigned long long *)d_acc;
    sink.fused = true;
    return enqueue(ctx, *plan, job, first_chunk, n_chunks, sink, (cudaStream_t)stream);
}

int mcb200_basket_launch(mcb200_ctx *ctx, const mcb200_plan_t *plan, const mcb200_basket_t *opt, uint64_t seed,
                         uint64_t first_chunk, uint64_t n_chunks, uint64_t *d_acc, void *stream)
{
    if (!ctx || !plan || !d_acc || plan->workload != MCB200_BASKET)
        return MCB200_ERR_INVALID;
    int st = check_range(plan, first_chunk, n_chunks);
    if (st != MCB200_OK)
        return st;
    AnyJob job;
    st = make_basket_job(opt, seed, &job.basket);
    if (st != MCB200_OK)
        return st;
    std::lock_guard<std::mutex> lock(ctx->mu);
    DeviceGuard guard(ctx->device);
    MCB_CUDA(ctx, guard.status());
    Sink sink;
    sink.d_acc = (unsigned long long *)d_acc;
    sink.fused = true;
    return enqueue(ctx, *plan, job, first_chunk, n_chunks, sink, (cudaStream_t)stream);
}

int mcb200_cva_launch(mcb200_ctx *ctx, const mcb200_plan_t *plan, const mcb200_cva_t *cva, uint64_t seed,
                      uint64_t first_chunk, uint64_t n_chunks, uint64_t *d_acc, void *stream)
{
    if (!ctx || !plan || !d_acc || plan->workload != MCB200_CVA)
        return MCB200_ERR_INVALID;
    int st = check_range(plan, first_chunk, n_chunks);
    if (st != MCB200_OK)
        return st;
    AnyJob job;
    st = make_cva_job(plan->precision, cva, seed, &job.cva);
    if (st != MCB200_OK)
        return st;
    std::lock_guard<std::mutex> lock(ctx->mu);
    DeviceGuard guard(ctx->device);
    MCB_CUDA(ctx, guard.status());
    Sink sink;
    sink.d_acc = (unsigned long long *)d_acc;
    sink.fused = true;
    return enqueue(ctx, *plan, job, first_chunk, n_chunks, sink, (cudaStream_t)stream);
}

// ---- one-call pricing ----
int mcb200_vanilla_multi(mcb200_ctx **ctxs, int n_ctx, int precision, const mcb200_option_t *opt,
                         uint64_t n_paths, uint64_t seed, mcb200_result_t *out)
{
    mcb200_plan_t plan;
    int st = mcb200_plan_vanilla(precision, opt, n_paths, &plan);
    if (st != MCB200_OK)
        return st;
    AnyJob job;
    st = make_vanilla_job(precision, opt, seed, &job.vanilla);
    if (st != MCB200_OK)
        return st;
    return price(ctxs, n_ctx, plan, job, out);
}

int mcb200_basket_multi(mcb200_ctx **ctxs, int n_ctx, int precision, const mcb200_basket_t *opt,
                        uint64_t n_paths, uint64_t seed, mcb200_result_t *out)
{
    mcb200_plan_t plan;
    int st = mcb200_plan_basket(precision, opt, n_paths, &plan);
    if (st != MCB200_OK)
        return st;
    AnyJob job;
    st = make_basket_job(opt, seed, &job.basket);
    if (st != MCB200_OK)
        return st;
    return price(ctxs, n_ctx, plan, job, out);
}

int mcb200_cva_multi(mcb200_ctx **ctxs, int n_ctx, int precision, const mcb200_cva_t *cva, uint64_t n_paths,
                     uint64_t seed, mcb200_result_t *out)
{
    mcb200_plan_t plan;
    int st = mcb200_plan_cva(precision, cva, n_paths, &plan);
    if (st != MCB200_OK)
        return st;
    AnyJob job;
    st = make_cva_job(precision, cva, seed, &job.cva);
    if (st != MCB200_OK)
        return st;
    return price(ctxs, n_ctx, plan, job, out);
}

int mcb200_vanilla(mcb200_ctx *ctx, int precision, const mcb200_option_t *opt, uint64_t n_paths, uint64_t seed,
                   mcb200_result_t *out)
{
    return mcb200_vanilla_multi(&ctx, 1, precision, opt, n_paths, seed, out);
}

int mcb200_basket(mcb200_ctx *ctx, int precision, const mcb200_basket_t *opt, uint64_t n_paths, uint64_t seed,
                  mcb200_result_t *out)
{
    return mcb200_basket_multi(&ctx, 1, precision, opt, n_paths, seed, out);
}

int mcb200_cva(mcb200_ctx *ctx, int precision, const mcb200_cva_t *cva, uint64_t n_paths, uint64_t seed,
               mcb200_result_t *out)
{
    return mcb200_cva_multi(&ctx, 1, precision, cva, n_paths, seed, out);
}

// ---- batched pricing: many jobs, few launches, one wait ----
// The reference's cvaOpt driver prices 5 grids x 4 thread counts one blocking call at a time, each
// with its own allocations, XORWOW seeding and device-to-host copy (double_precision/cvaOpt.cu:70-109).
// Here the European-call and CVA jobs of a batch that share a kernel (workload, precision) go out as ONE launch of
// the multi-job kernel (device_common.cuh, mc_accumulate_batch_kernel; up to kBatchMaxJobs jobs and, for the CVA,
// kCvaMaxDates exposure dates per launch), basket jobs as one launch each (their factor lives at fixed offsets of
// a __constant__ table); every launch overlaps the tail of the one before it, every job's result lands in its own
// slot of mapped host memory, and the host waits once.  A job's result is exactly what the one-call API returns
// for it (same chunks, same integer limbs).
int mcb200_price_batch(mcb200_ctx *ctx, int n_jobs, const mcb200_job_t *jobs, mcb200_result_t *out, int *status_out)
{
    if (!ctx || n_jobs < 1 || n_jobs > 65535 || !jobs || !out)
        return MCB200_ERR_INVALID;
    std::lock_guard<std::mutex> lock(ctx->mu);
    DeviceGuard guard(ctx->device);
    MCB_CUDA(ctx, guard.status());
    if ((size_t)n_jobs > ctx->batch_capacity) {
        if (ctx->h_batch_slots) {
            cudaStreamSynchronize(ctx->stream);
            cudaFreeHost(ctx->h_batch_slots);
        }
        ctx->h_batch_slots = ctx->d_batch_slots = nullptr;
        ctx->batch_capacity = 0;
        const size_t bytes = sizeof(unsigned long long) * kPeerSlotWords * (size_t)n_jobs;
        MCB_CUDA(ctx, cudaHostAlloc(&ctx->h_batch_slots, bytes, cudaHostAllocMapped));
        MCB_CUDA(ctx, cudaHostGetDevicePointer(&ctx->d_batch_slots, ctx->h_batch_slots, 0));
        std::memset(ctx->h_batch_slots, 0, bytes);
        ctx->batch_capacity = (size_t)n_jobs;
    }
    std::vector<mcb200_plan_t> plans((size_t)n_jobs);
    std::vector<AnyJob> built((size_t)n_jobs);
    std::vector<int> status((size_t)n_jobs, MCB200_OK);
    for (int i = 0; i < n_jobs; i++) {
        const mcb200_job_t &j = jobs[i];
        int st = MCB200_ERR_INVALID;
        if (j.params) {
            switch (j.workload) {
                case MCB200_VANILLA:
                    st = mcb200_plan_vanilla(j.precision, (const mcb200_option_t *)j.params, j.n_paths, &plans[i]);
                    if (st == MCB200_OK)
                        st = make_vanilla_job(j.precision, (const mcb200_option_t *)j.params, j.seed, &built[i].vanilla);
                    break;
                case MCB200_BASKET:
                    st = mcb200_plan_basket(j.precision, (const mcb200_basket_t *)j.params, j.n_paths, &plans[i]);
                    if (st == MCB200_OK)
                        st = make_basket_job((const mcb200_basket_t *)j.params, j.seed, &built[i].basket);
                    break;
                case MCB200_CVA:
                    st = mcb200_plan_cva(j.precision, (const mcb200_cva_t *)j.params, j.n_paths, &plans[i]);
                    if (st == MCB200_OK)
                        st = make_cva_job(j.precision, (const mcb200_cva_t *)j.params, j.seed, &built[i].cva);
                    break;
                default: break;
            }
        }
        if (st == MCB200_OK && plans[i].n_chunks > (1ull << 30))
            st = MCB200_ERR_UNSUPPORTED;
        status[i] = st;
    }
    order_streams(ctx, ctx->stream);
    const unsigned long long flag = peer_flag(++ctx->results);
    if (ctx->timing)
        MCB_CUDA(ctx, cudaEventRecord(ctx->ev_begin, ctx->stream));
    // consecutive launches of the batch are independent: let each start while its predecessor drains (one in
    // kCtlRing stays a full stream dependency, so the ring of control blocks can never be lapped)
    auto options = [&]() {
        LaunchOptions opt;
        opt.overlap = (ctx->launches % kCtlRing) != 0;
        return opt;
    };
    // ---- multi-job launches: European calls and CVAs grouped by kernel ----
    std::vector<char> done((size_t)n_jobs, 0);
    for (int workload : {MCB200_VANILLA, MCB200_CVA}) {
        for (int precision : {MCB200_F32, MCB200_F64}) {
            std::vector<int> members;
            for (int i = 0; i < n_jobs; i++)
                if (status[i] == MCB200_OK && jobs[i].workload == workload && jobs[i].precision == precision &&
                    !(workload == MCB200_CVA && built[i].cva.job.n_dates > kCvaMaxDates))   // a long grid has its own table and launch
                    members.push_back(i);
            if (members.size() < 2)
                continue;  // a lone job takes the one-job kernel below (its parameters sit in the constant bank)
            // heaviest chunks first: what is claimed last is then the cheapest
            auto chunk_cost = [&](int i) {
                return (double)plans[i].rounds * (workload == MCB200_CVA ? (double)std::max(built[i].cva.job.n_dates, 1) : 1.0);
            };
            std::stable_sort(members.begin(), members.end(), [&](int a, int b) { return chunk_cost(a) > chunk_cost(b); });
            size_t at = 0;
            while (at < members.size()) {
                BatchShape shape{};
                std::vector<VanillaJob> vjobs;
                std::vector<CvaJob> cjobs;
                uint64_t chunks = 0;
                int dates = 0;
                while (at < members.size() && shape.n_jobs < kBatchMaxJobs) {
                    const int i = members[at];
                    const int nd = workload == MCB200_CVA ? built[i].cva.job.n_dates : 0;
                    if (shape.n_jobs > 0 && (chunks + plans[i].n_chunks > (1ull << 30) || dates + nd > kCvaMaxDates))
                        break;
                    shape.geo[shape.n_jobs] = job_geometry(plans[i]);
                    shape.n_chunks[shape.n_jobs] = (unsigned int)plans[i].n_chunks;
                    shape.slot[shape.n_jobs] = (unsigned short)i;
                    shape.n_jobs++;
                    chunks += plans[i].n_chunks;
                    dates += nd;
                    if (workload == MCB200_CVA)
                        cjobs.push_back(built[i].cva.job);
                    else
                        vjobs.push_back(built[i].vanilla);
                    at++;
                }
                const int per_sm = workload == MCB200_CVA ? cva_batch_blocks_per_sm(precision) : vanilla_batch_blocks_per_sm(precision);
                uint64_t grid = (uint64_t)ctx->sm_count * (uint64_t)std::max(per_sm, 0);
                if (grid > chunks)
                    grid = chunks;
                cudaError_t e = cudaErrorLaunchOutOfResources;
                if (grid > 0) {
                    const int ring = (int)(ctx->launches % kCtlRing);
                    BatchTarget target;
                    target.ctl = ctx->d_ctl + ring;
                    target.d_acc = ctx->d_batch_acc + (size_t)ring * kBatchMaxJobs * kAccWords;
                    target.host_slots = ctx->d_batch_slots;
                    target.host_flag = flag;
                    e = workload == MCB200_CVA
                            ? cva_batch_launch(precision, shape, cjobs.data(), (int)grid, target, ctx->stream, options())
                            : vanilla_batch_launch(precision, shape, vjobs.data(), (int)grid, target, ctx->stream, options());
                }
                for (int k = 0; k < shape.n_jobs; k++) {
                    done[shape.slot[k]] = 1;
                    if (e != cudaSuccess)
                        status[shape.slot[k]] = fail_cuda(ctx, e, "multi-job kernel launch");
                }
                if (e == cudaSuccess)
                    ctx->launches++;
            }
        }
    }
    // ---- everything else: one launch per job ----
    for (int i = 0; i < n_jobs; i++) {
        if (status[i] != MCB200_OK || done[i])
            continue;
        Sink sink;
        sink.host_slot = ctx->d_batch_slots + (size_t)i * kPeerSlotWords;
        sink.host_flag = flag;
        const bool keep = ctx->overlap;
        ctx->overlap = options().overlap;
        status[i] = enqueue(ctx, plans[i], built[i], 0, plans[i].n_chunks, sink, ctx->stream);
        ctx->overlap = keep;
    }
    if (ctx->timing)
        MCB_CUDA(ctx, cudaEventRecord(ctx->ev_end, ctx->stream));
    float ms = 0;
    int worst = MCB200_OK;
    for (int i = 0; i < n_jobs; i++) {
        std::memset(&out[i], 0, sizeof out[i]);
        if (status[i] == MCB200_OK) {
            uint64_t acc[MCB200_ACC_WORDS];
            status[i] = wait_slot(ctx, ctx->stream, ctx->h_batch_slots + (size_t)i * kPeerSlotWords, flag, acc);
            if (status[i] == MCB200_OK)
                status[i] = mcb200_finalize(&plans[i], acc, &out[i]);
        }
    }
    if (ctx->timing && (cudaEventSynchronize(ctx->ev_end) != cudaSuccess || cudaEventElapsedTime(&ms, ctx->ev_begin, ctx->ev_end) != cudaSuccess)) {
        cudaGetLastError();
        ms = 0;
    }
    for (int i = 0; i < n_jobs; i++) {
        if (status[i] == MCB200_OK)
            out[i].kernel_ms = ms;  // device time of the whole batch
        if (status_out)
            status_out[i] = status[i];
        if (status[i] != MCB200_OK && worst == MCB200_OK)
            worst = status[i];
    }
    if (worst != MCB200_OK)
        return fail(ctx, worst, "at least one job of the batch failed; see the per-job status");
    return MCB200_OK;
}

// ---- per-path values ----
int mcb200_vanilla_paths(mcb200_ctx *ctx, int precision, const mcb200_option_t *opt, uint64_t seed,
                         uint64_t first_path, uint64_t n_paths, void *out_host)
{
    mcb200_plan_t plan;
    int st = mcb200_plan_vanilla(precision, opt, first_path + n_paths, &plan);
    if (st != MCB200_OK)
        return st;
    AnyJob job;
    st = make_vanilla_job(precision, opt, seed, &job.vanilla);
    if (st != MCB200_OK)
        return st;
    return paths(ctx, plan, job, first_path, n_paths, out_host);
}

int mcb200_basket_paths(mcb200_ctx *ctx, int precision, const mcb200_basket_t *opt, uint64_t seed,
                        uint64_t first_path, uint64_t n_paths, void *out_host)
{
    mcb200_plan_t plan;
    int st = mcb200_plan_basket(precision, opt, first_path + n_paths, &plan);
    if (st != MCB200_OK)
        return st;
    AnyJob job;
    st = make_basket_job(opt, seed, &job.basket);
    if (st != MCB200_OK)
        return st;
    return paths(ctx, plan, job, first_path, n_paths, out_host);
}

int mcb200_cva_paths(mcb200_ctx *ctx, int precision, const mcb200_cva_t *cva, uint64_t seed, uint64_t first_path,
                     uint64_t n_paths, void *out_host)
{
    mcb200_plan_t plan;
    int st = mcb200_plan_cva(precision, cva, first_path + n_paths, &plan);
    if (st != MCB200_OK)
        return st;
    AnyJob job;
    st = make_cva_job(precision, cva, seed, &job.cva);
    if (st != MCB200_OK)
        return st;
    return paths(ctx, plan, job, first_path, n_paths, out_host);
}

// ---- generator / reduction instrumentation ----
int mcb200_debug_philox(mcb200_ctx *ctx, uint64_t n, const uint32_t *ctr_host, const uint32_t key[2],
                        uint32_t *out_host)
{
    if (!ctx || !ctr_host || !key || !out_host || n == 0)
        return MCB200_ERR_INVALID;
    std::lock_guard<std::mutex> lock(ctx->mu);
    DeviceGuard guard(ctx->device);
    MCB_CUDA(ctx, guard.status());
    uint32_t *d_ctr = nullptr, *d_out = nullptr;
    MCB_CUDA(ctx, cudaMalloc(&d_ctr, n * 16));
    cudaError_t e = cudaMalloc(&d_out, n * 16);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(d_ctr, ctr_host, n * 16, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess)
        e = debug_philox(n, d_ctr, make_keys((uint64_t)key[0] | ((uint64_t)key[1] << 32)), d_out, ctx->stream);
    if (e == cudaSuccess) {
        ctx->launches++;
        e = cudaMemcpyAsync(out_host, d_out, n * 16, cudaMemcpyDeviceToHost, ctx->stream);
    }
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_ctr);
    cudaFree(d_out);
    if (e != cudaSuccess)
        return fail_cuda(ctx, e, "debug_philox");
    return MCB200_OK;
}

int mcb200_debug_normals(mcb200_ctx *ctx, int precision, uint64_t n, const uint32_t *ctr_host,
                         const uint32_t key[2], void *out_host)
{
    if (!ctx || !ctr_host || !key || !out_host || n == 0 || (precision != MCB200_F32 && precision != MCB200_F64))
        return MCB200_ERR_INVALID;
    std::lock_guard<std::mutex> lock(ctx->mu);
    DeviceGuard guard(ctx->device);
    MCB_CUDA(ctx, guard.status());
    uint32_t *d_ctr = nullptr;
    void *d_out = nullptr;
    MCB_CUDA(ctx, cudaMalloc(&d_ctr, n * 16));
    const size_t out_bytes = n * (precision == MCB200_F64 ? 4 * 8 : 6 * 4);  // 4 fp64 / 6 fp32 normals per counter
    cudaError_t e = cudaMalloc(&d_out, out_bytes);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(d_ctr, ctr_host, n * 16, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess)
        e = debug_normals(precision, n, d_ctr, make_keys((uint64_t)key[0] | ((uint64_t)key[1] << 32)), d_out,
                          ctx->stream);
    if (e == cudaSuccess) {
        ctx->launches++;
        e = cudaMemcpyAsync(out_host, d_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream);
    }
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_ctr);
    cudaFree(d_out);
    if (e != cudaSuccess)
        return fail_cuda(ctx, e, "debug_normals");
    return MCB200_OK;
}

int mcb200_debug_math64(mcb200_ctx *ctx, int fn, uint64_t n, const double *in_host, double *out_host)
{
    if (!ctx || !in_host || !out_host || n == 0 || fn < 0 || fn > 8)
        return MCB200_ERR_INVALID;
    std::lock_guard<std::mutex> lock(ctx->mu);
    DeviceGuard guard(ctx->device);
    MCB_CUDA(ctx, guard.status());
    double *d_in = nullptr, *d_out = nullptr;
    MCB_CUDA(ctx, cudaMalloc(&d_in, n * sizeof(double)));
    cudaError_t e = cudaMalloc(&d_out, 2 * n * sizeof(double));
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(d_in, in_host, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess)
        e = debug_math64(fn, n, d_in, d_out, ctx->stream);
    if (e == cudaSuccess) {
        ctx->launches++;
        e = cudaMemcpyAsync(out_host, d_out, 2 * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    }
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_in);
    cudaFree(d_out);
    if (e != cudaSuccess)
        return fail_cuda(ctx, e, "debug_math64");
    return MCB200_OK;
}

int mcb200_debug_reduce(mcb200_ctx *ctx, const double *values_host, uint64_t n_valid, int unit_paths, int rounds,
                        int accumulate_in_float, int scale_exp_sum, int scale_exp_sumsq,
                        uint64_t acc_host[MCB200_ACC_WORDS])
{
    if (!ctx || !values_host || !acc_host || unit_paths < 1 || rounds < 1 || rounds > 64 ||
        n_valid > (uint64_t)kThreads * rounds * unit_paths)
        return MCB200_ERR_INVALID;
    std::lock_guard<std::mutex> lock(ctx->mu);
    DeviceGuard guard(ctx->device);
    MCB_CUDA(ctx, guard.status());
    double *d_values = nullptr;
    const size_t bytes = (size_t)(n_valid > 0 ? n_valid : 1) * sizeof(double);
    MCB_CUDA(ctx, cudaMalloc(&d_values, bytes));
    cudaError_t e = cudaMemcpyAsync(d_values, values_host, n_valid * sizeof(double), cudaMemcpyHostToDevice,
                                    ctx->stream);
    if (e == cudaSuccess)
        e = cudaMemsetAsync(ctx->d_acc, 0, sizeof(unsigned long long) * kAccWords, ctx->stream);
    if (e == cudaSuccess)
        e = debug_reduce(d_values, n_valid, unit_paths, rounds, accumulate_in_float != 0, scale_exp_sum,
                         scale_exp_sumsq, ctx->d_acc, ctx->stream);
    if (e == cudaSuccess) {
        ctx->launches++;
        e = cudaMemcpyAsync(ctx->h_acc, ctx->d_acc, sizeof(unsigned long long) * kAccWords, cudaMemcpyDeviceToHost,
                            ctx->stream);
    }
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_values);
    if (e != cudaSuccess)
        return fail_cuda(ctx, e, "debug_reduce");
    std::memcpy(acc_host, ctx->h_acc, sizeof(uint64_t) * MCB200_ACC_WORDS);
    return MCB200_OK;
}

}  // extern "C"
