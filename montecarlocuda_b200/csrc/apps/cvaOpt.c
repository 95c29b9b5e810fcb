/* cvaOpt.c -- CVA of a European call: sweep over exposure-grid sizes, GPU (and optionally CPU).
 * Non-interactive counterpart of the reference driver double_precision/cvaOpt.cu:30-111, which runs
 * grids {25, 50, 75, 250, 500} x thread counts {128..1024} at 131072 paths with the CPU leg commented out.
 *   mcb200_cvaOpt_dp [--sims 131072] [--grids 25,50,75,250,500] [--cpu] [--cpu-sims N] [--gpus G] [--seed S]
 *                    [--lambda 0.03 --recovery 0.4 --s 100 --k 100 --r 0.05 --v 0.2 --t 1] */
#include "cli_common.h"

int main(int argc, char **argv)
{
    if (arg_flag(argc, argv, "--help")) {
        puts("usage: cvaOpt [--sims N] [--grids a,b,c] [--cpu] [--cpu-sims N] [--gpus G] [--seed S] [--lambda L --recovery R --s --k --r --v --t]");
        return 0;
    }
    forward_gpus(argc, argv);
    CVA cva;
    cva.defInt = (mc_real)arg_num(argc, argv, "--lambda", 0.03);   /* reference defaults: cvaOpt.cu:22-28 */
    cva.lgd = (mc_real)(1 - arg_num(argc, argv, "--recovery", 0.4));
    cva.ns = 1;
    cva.option.s = (mc_real)arg_num(argc, argv, "--s", 100);
    cva.option.k = (mc_real)arg_num(argc, argv, "--k", 100);
    cva.option.r = (mc_real)arg_num(argc, argv, "--r", 0.05);
    cva.option.v = (mc_real)arg_num(argc, argv, "--v", 0.2);
    cva.option.t = (mc_real)arg_num(argc, argv, "--t", 1);
    const int sims = (int)arg_num(argc, argv, "--sims", 131072);
    const int cpu_sims = (int)arg_num(argc, argv, "--cpu-sims", sims > 131072 ? 131072 : sims);
    const int with_cpu = arg_flag(argc, argv, "--cpu");
    char grids[256];
    snprintf(grids, sizeof grids, "%s", arg_str(argc, argv, "--grids", "25,50,75,250,500"));

    printf("CVA of an European call Option (%s precision)\nDefault intensity %.2f, LGD %.2f\n", PRECISION_NAME, (double)cva.defInt,
           (double)cva.lgd);
    printOption(cva.option);
    cva.n = 1;
    dev_cvaEquityOption(&cva, 1024, 128, 1024);                    /* warm-up */
    for (char *tok = strtok(grids, ","); tok; tok = strtok(NULL, ",")) {
        cva.n = atoi(tok);
        if (cva.n < 1)
            continue;
        printf("\nexposure_dates %d\nloop_iterations %lld\n", cva.n, (long long)cva.n * sims);
        double cpu_ms = 0;
        if (with_cpu) {
            double t0 = now_ms();
            OptionValue cpu = host_cvaEquityOption(&cva, cpu_sims);
            cpu_ms = now_ms() - t0;
            printf("cpu_sims %d\ncpu_cva %f\ncpu_confidence %f\ncpu_time_ms %f\n", cpu_sims, (double)cpu.Expected, (double)cpu.Confidence, cpu_ms);
        }
        double t0 = now_ms();
        OptionValue gpu = dev_cvaEquityOption(&cva, 1024, 128, sims);
        double gpu_ms = now_ms() - t0;
        printf("gpu_sims %d\ngpu_cva %f\ngpu_confidence %f\ngpu_time_ms %f\n", 1024 * (sims / 1024), (double)gpu.Expected,
               (double)gpu.Confidence, gpu_ms);
        if (cpu_ms > 0)
            printf("speedup_per_path %.2f\n", (cpu_ms / cpu_sims) / (gpu_ms / (1024 * (double)(sims / 1024))));
    }
    return 0;
}
