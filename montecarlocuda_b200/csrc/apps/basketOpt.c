/* basketOpt.c -- basket call on N correlated underlyings: CPU Monte Carlo vs GPU Monte Carlo.
 * Non-interactive counterpart of the reference driver double_precision/basketOpt.cu:27-182 (N is a
 * compile-time macro there and here: mcb200_basketOpt_dp, _dp_n10, _dp_n64, ...).
 *   mcb200_basketOpt_dp [--sims N] [--cpu-sims N] [--rho 0.3] [--reference-data] [--gpus G] [--seed S] [--no-cpu]
 * Default market data: s = 100, w = 1/N, v = 0.3/0.2 alternating (the reference's getRandomSigma),
 * k = 100, r = 0.048790164, t = 1, equicorrelation rho.  --reference-data uses the reference driver's
 * own matrices instead (all -0.5 at N = 3, +-0.5 by column parity otherwise: basketOpt.cu:42-68,160-177),
 * which are singular / indefinite; mcb200_chol reports that, Chol silently zeroes columns. */
#include "cli_common.h"

int main(int argc, char **argv)
{
    if (arg_flag(argc, argv, "--help")) {
        puts("usage: basketOpt [--sims N] [--cpu-sims N] [--rho R] [--reference-data] [--gpus G] [--seed S] [--no-cpu]");
        return 0;
    }
    forward_gpus(argc, argv);
    MultiOptionData option;
    const double rho = arg_num(argc, argv, "--rho", 0.3);
    const int ref_data = arg_flag(argc, argv, "--reference-data");
    for (int i = 0; i < N; i++) {
        option.s[i] = 100;
        option.w[i] = (mc_real)1 / N;
        option.d[i] = 0;
        option.v[i] = (mc_real)(N == 3 ? (i == 1 ? 0.3 : 0.2) : (i % 2 == 0 ? 0.3 : 0.2));
        for (int j = 0; j < N; j++) {
            double c = i == j ? 1.0 : rho;
            if (ref_data && i != j)
                c = N == 3 ? -0.5 : ((i > j ? i : j) % 2 == 0 ? 0.5 : -0.5);
            option.p[i][j] = (mc_real)c;
        }
    }
    option.k = 100;
    option.r = (mc_real)0.048790164;
    option.t = 1;
    const int sims = (int)arg_num(argc, argv, "--sims", 8 * 131072);
    const int cpu_sims = (int)arg_num(argc, argv, "--cpu-sims", sims > 1 << 20 ? 1 << 20 : sims);
    const int blocks = (int)arg_num(argc, argv, "--blocks", 512), threads = (int)arg_num(argc, argv, "--threads", 128);

    printf("Basket Option Pricing (%s precision, %d underlyings)\n", PRECISION_NAME, N);
    if (N < 7)
        printMultiOpt(&option);
    static double c64[N * N], a64[N * N];
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++)
            c64[i * N + j] = (double)option.p[i][j];
    const int bad = mcb200_chol(N, c64, a64);
    if (bad)
        printf("warning: correlation matrix is not positive definite (pivot %d): factor is rank-deficient\n", bad);
    static mc_real factor[N][N];
    Chol(option.p, factor);                                       /* basketOpt.cu:96-99: p <- Cholesky factor */
    memcpy(option.p, factor, sizeof factor);

    double cpu_ms = 0;
    OptionValue cpu = {0, 0};
    if (!arg_flag(argc, argv, "--no-cpu")) {
        double t0 = now_ms();
        cpu = host_basketOpt(&option, cpu_sims);
        cpu_ms = now_ms() - t0;
        printf("cpu_sims %d\ncpu_price %f\ncpu_confidence %f\ncpu_time_ms %f\n", cpu_sims, (double)cpu.Expected, (double)cpu.Confidence, cpu_ms);
    }
    dev_basketOpt(&option, blocks, threads, blocks);
    double t0 = now_ms();
    OptionValue gpu = dev_basketOpt(&option, blocks, threads, sims);
    double gpu_ms = now_ms() - t0;
    printf("gpu_sims %d\ngpu_price %f\ngpu_confidence %f\ngpu_time_ms %f\n", blocks * (sims / blocks), (double)gpu.Expected,
           (double)gpu.Confidence, gpu_ms);
    if (cpu_ms > 0) {
        printf("difference_gpu_cpu %f\n", fabs((double)gpu.Expected - (double)cpu.Expected));
        printf("speedup_per_path %.2f\n", (cpu_ms / cpu_sims) / (gpu_ms / (blocks * (double)(sims / blocks))));
    }
    return 0;
}
