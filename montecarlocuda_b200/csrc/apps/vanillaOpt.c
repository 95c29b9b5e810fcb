/* vanillaOpt.c -- European call: Black-Scholes, CPU Monte Carlo, GPU Monte Carlo.
 * Non-interactive counterpart of the reference driver double_precision/vanillaOpt.cu:28-107.
 *   mcb200_vanillaOpt_dp [--sims 1048576] [--cpu-sims 1048576] [--blocks 512] [--threads 128]
 *                        [--s 100 --k 100 --r 0.048790 --v 0.2 --t 1] [--gpus G] [--seed S] [--no-cpu] */
#include "cli_common.h"

int main(int argc, char **argv)
{
    if (arg_flag(argc, argv, "--help")) {
        puts("usage: vanillaOpt [--sims N] [--cpu-sims N] [--blocks B] [--threads T] [--s --k --r --v --t] [--gpus G] [--seed S] [--no-cpu]");
        return 0;
    }
    forward_gpus(argc, argv);
    OptionData option;
    option.s = (mc_real)arg_num(argc, argv, "--s", 100);      /* reference defaults: vanillaOpt.cu:22-26 */
    option.k = (mc_real)arg_num(argc, argv, "--k", 100);
    option.r = (mc_real)arg_num(argc, argv, "--r", 0.048790);
    option.v = (mc_real)arg_num(argc, argv, "--v", 0.2);
    option.t = (mc_real)arg_num(argc, argv, "--t", 1);
    const int sims = (int)arg_num(argc, argv, "--sims", 8 * 131072);
    const int cpu_sims = (int)arg_num(argc, argv, "--cpu-sims", sims);
    const int blocks = (int)arg_num(argc, argv, "--blocks", 512), threads = (int)arg_num(argc, argv, "--threads", 128);

    printf("Vanilla Option Pricing (%s precision)\n", PRECISION_NAME);
    printOption(option);
    const double bs = (double)host_bsCall(option);
    printf("\nblack_scholes_price %f\n", bs);

    double cpu_ms = 0;
    if (!arg_flag(argc, argv, "--no-cpu")) {
        double t0 = now_ms();
        OptionValue cpu = host_vanillaOpt(option, cpu_sims);
        cpu_ms = now_ms() - t0;
        printf("cpu_sims %d\ncpu_price %f\ncpu_confidence %f\ncpu_difference_from_bs %f\ncpu_time_ms %f\n", cpu_sims, (double)cpu.Expected,
               (double)cpu.Confidence, fabs((double)cpu.Expected - bs), cpu_ms);
    }
    dev_vanillaOpt(&option, blocks, threads, blocks);            /* warm-up: context creation is not pricing time */
    double t0 = now_ms();
    OptionValue gpu = dev_vanillaOpt(&option, blocks, threads, sims);
    double gpu_ms = now_ms() - t0;
    printf("gpu_sims %d\ngpu_price %f\ngpu_confidence %f\ngpu_difference_from_bs %f\ngpu_time_ms %f\n", blocks * (sims / blocks),
           (double)gpu.Expected, (double)gpu.Confidence, fabs((double)gpu.Expected - bs), gpu_ms);
    if (cpu_ms > 0)
        printf("speedup_per_path %.2f\n", (cpu_ms / cpu_sims) / (gpu_ms / (blocks * (double)(sims / blocks))));
    return 0;
}
