/* cli_common.h -- shared bits of the three non-interactive drivers.
 * The reference drivers (double_precision/vanillaOpt.cu, basketOpt.cu, cvaOpt.cu) hard-code the
 * market data, read a path multiplier with scanf and time CPU and GPU with cudaEvents; these take
 * everything from argv, use the same API (MonteCarlo.h: host_* and dev_*) and print the same fields
 * (price, 95 % half-width, difference, time, speed-up), one "key value" pair per line. */
#ifndef MCB200_CLI_COMMON_H_
#define MCB200_CLI_COMMON_H_

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "MonteCarlo.h"
#include "mcb200.h"

static double now_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

/* --name value lookup; returns fallback when absent */
static const char *arg_str(int argc, char **argv, const char *name, const char *fallback)
{
    for (int i = 1; i + 1 < argc; i++)
        if (strcmp(argv[i], name) == 0)
            return argv[i + 1];
    return fallback;
}
static double arg_num(int argc, char **argv, const char *name, double fallback)
{
    const char *s = arg_str(argc, argv, name, NULL);
    return s ? strtod(s, NULL) : fallback;
}
static int arg_flag(int argc, char **argv, const char *name)
{
    for (int i = 1; i < argc; i++)
        if (strcmp(argv[i], name) == 0)
            return 1;
    return 0;
}

/* --gpus G is forwarded to the drop-in library through its environment knob (dropin.cpp) */
static void forward_gpus(int argc, char **argv)
{
    const char *g = arg_str(argc, argv, "--gpus", NULL);
    if (g)
        setenv("MCB200_GPUS", g, 1);
    const char *s = arg_str(argc, argv, "--seed", NULL);
    if (s)
        setenv("MCB200_SEED", s, 1);
}

#ifdef MCB200_SINGLE
#define PRECISION_NAME "single"
#else
#define PRECISION_NAME "double"
#endif

#endif
