// dropin.cpp -- the reference's three GPU entry points on top of libmcb200.
//
// Built once per precision (and per basket width N, a compile-time macro of the reference header)
// into libmcb200_dp[_nN].so / libmcb200_sp[_nN].so; exports exactly the symbols
//   OptionValue dev_vanillaOpt     (OptionData *,      int numBlocks, int numThreads, int sims)  DP/MonteCarloKernel.cu:500
//   OptionValue dev_basketOpt      (MultiOptionData *, int numBlocks, int numThreads, int sims)  DP/MonteCarloKernel.cu:483
//   OptionValue dev_cvaEquityOption(CVA *,             int numBlocks, int numThreads, int sims)  DP/MonteCarloKernel.cu:517
// with the reference's semantics: input structs are read-only, n = numBlocks * (sims / numBlocks)
// paths are simulated (:491,508,524), Expected is discounted for pricing and plain for CVA, Confidence
// is 1.96 s / sqrt(n) of the undiscounted value (:420-423, :466-469), and any failure prints a
// message and exit(1)s (DP/MonteCarlo.h:22-30).  numThreads no longer shapes the launch.
//
// Differences a caller can observe: no timing chatter on stdout (set MCB200_VERBOSE=1 for one line
// per call), results do not depend on (numBlocks, numThreads) beyond n, and the context (stream,
// buffers) is created on first use and kept.  Environment: MCB200_DEVICE (default 0),
// MCB200_GPUS (default 1: number of devices to spread paths over, starting at MCB200_DEVICE),
// MCB200_SEED (default below; the reference's seeds are fixed too, DP/MonteCarloKernel.cu:289).
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "../../include/MonteCarlo.h"
#include "../../include/mcb200.h"

namespace {

#ifdef MCB200_SINGLE
constexpr int kPrecision = MCB200_F32;
#else
constexpr int kPrecision = MCB200_F64;
#endif

constexpr uint64_t kDefaultSeed = 0x6d63623230300001ull;  // "mcb200" 0001

struct Global {
    std::mutex mu;
    std::vector<mcb200_ctx *> ctxs;
    uint64_t seed = kDefaultSeed;
    bool verbose = false;
    bool ready = false;
};

Global &global()
{
    static Global g;
    return g;
}

[[noreturn]] void die(const char *where, int status, mcb200_ctx *ctx)
{
    std::fprintf(stderr, "mcb200: %s failed: %s%s%s\n", where, mcb200_strerror(status),
                 ctx && mcb200_last_error(ctx)[0] ? " -- " : "", ctx ? mcb200_last_error(ctx) : "");
    std::exit(1);
}

Global &engine()
{
    Global &g = global();
    std::lock_guard<std::mutex> lock(g.mu);
    if (g.ready)
        return g;
    const char *dev = std::getenv("MCB200_DEVICE");
    const char *gpus = std::getenv("MCB200_GPUS");
    const char *seed = std::getenv("MCB200_SEED");
    const char *verbose = std::getenv("MCB200_VERBOSE");
    const int first = dev ? std::atoi(dev) : 0;
    const int count = gpus ? std::atoi(gpus) : 1;
    if (seed)
        g.seed = std::strtoull(seed, nullptr, 0);
    g.verbose = verbose && verbose[0] && verbose[0] != '0';
    for (int i = 0; i < (count < 1 ? 1 : count); i++) {
        mcb200_ctx *ctx = nullptr;
        const int st = mcb200_create(&ctx, first + i);
        if (st != MCB200_OK)
            die("mcb200_create", st, nullptr);
        if (g.verbose)
            mcb200_set_option(ctx, MCB200_OPT_TIMING, 1);  // the kernel time printed below
        g.ctxs.push_back(ctx);
    }
    g.ready = true;
    return g;
}

// n = numBlocks * (sims / numBlocks), the reference's integer arithmetic
uint64_t simulated_paths(int numBlocks, int sims)
{
    if (numBlocks <= 0 || sims <= 0) {
        std::fprintf(stderr, "mcb200: numBlocks and sims must be positive (got %d, %d)\n", numBlocks, sims);
        std::exit(1);
    }
    return (uint64_t)numBlocks * (uint64_t)(sims / numBlocks);
}

OptionValue to_value(const mcb200_result_t &r, const char *what, bool verbose)
{
    if (verbose)
        std::printf("mcb200 %s: n=%llu expected=%.9g confidence=%.3g kernel=%.3f ms\n", what,
                    (unsigned long long)r.n_paths, r.expected, r.confidence, r.kernel_ms);
    OptionValue v;
    v.Expected = (mc_real)r.expected;
    v.Confidence = (mc_real)r.confidence;
    return v;
}

}  // namespace

extern "C" {

OptionValue dev_vanillaOpt(OptionData *opt, int numBlocks, int numThreads, int sims)
{
    (void)numThreads;
    Global &g = engine();
    const uint64_t n = simulated_paths(numBlocks, sims);
    if (n == 0 || !opt)
        die("dev_vanillaOpt", MCB200_ERR_INVALID, nullptr);
    const mcb200_option_t o = {(double)opt->s, (double)opt->k, (double)opt->r, (double)opt->v, (double)opt->t};
    mcb200_result_t r;
    const int st = mcb200_vanilla_multi(g.ctxs.data(), (int)g.ctxs.size(), kPrecision, &o, n, g.seed, &r);
    if (st != MCB200_OK)
        die("dev_vanillaOpt", st, g.ctxs[0]);
    return to_value(r, "vanilla", g.verbose);
}

OptionValue dev_basketOpt(MultiOptionData *option, int numBlocks, int numThreads, int sims)
{
    (void)numThreads;
    Global &g = engine();
    const uint64_t n = simulated_paths(numBlocks, sims);
    if (n == 0 || !option)
        die("dev_basketOpt", MCB200_ERR_INVALID, nullptr);
    double s[N], v[N], p[N * N], d[N], w[N];
    for (int i = 0; i < N; i++) {
        s[i] = (double)option->s[i];
        v[i] = (double)option->v[i];
        d[i] = (double)option->d[i];
        w[i] = (double)option->w[i];
        for (int j = 0; j < N; j++)
            p[i * N + j] = (double)option->p[i][j];
    }
    const mcb200_basket_t b = {N, s, v, p, d, w, (double)option->k, (double)option->t, (double)option->r};
    mcb200_result_t r;
    const int st = mcb200_basket_multi(g.ctxs.data(), (int)g.ctxs.size(), kPrecision, &b, n, g.seed, &r);
    if (st != MCB200_OK)
        die("dev_basketOpt", st, g.ctxs[0]);
    return to_value(r, "basket", g.verbose);
}

OptionValue dev_cvaEquityOption(CVA *cva, int numBlocks, int numThreads, int sims)
{
    (void)numThreads;
    Global &g = engine();
    const uint64_t n = simulated_paths(numBlocks, sims);
    if (n == 0 || !cva)
        die("dev_cvaEquityOption", MCB200_ERR_INVALID, nullptr);
    mcb200_cva_t c;
    c.def_int = (double)cva->defInt;
    c.lgd = (double)cva->lgd;
    c.option = {(double)cva->option.s, (double)cva->option.k, (double)cva->option.r, (double)cva->option.v,
                (double)cva->option.t};
    c.n_dates = cva->n;
    c.grid_mode = 0;  // the reference's time grid
    mcb200_result_t r;
    const int st = mcb200_cva_multi(g.ctxs.data(), (int)g.ctxs.size(), kPrecision, &c, n, g.seed, &r);
    if (st != MCB200_OK)
        die("dev_cvaEquityOption", st, g.ctxs[0]);
    return to_value(r, "cva", g.verbose);
}

}  // extern "C"
