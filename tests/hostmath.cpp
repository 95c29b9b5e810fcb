// hostmath.cpp -- host build of montecarlocuda_b200/csrc/device_math64.cuh (MCB_HOST_MATH) exposed
// through a C ABI so tests/test_device_math64.py can measure its accuracy against libm without a
// GPU.  Test infrastructure; the product uses the same header compiled for the device.
#define MCB_HOST_MATH
#include "../montecarlocuda_b200/csrc/device_math64.cuh"

using namespace mcb::hostmath;

static Tables64 g_tables;
static bool g_ready = false;
static const Tables64 &tables()
{
    if (!g_ready) {
        for (int i = 0; i < 256; i++)
            g_tables.fill(i);
        g_ready = true;
    }
    return g_tables;
}
// the exponent table of k ln u (what a kernel's sub-block fills per job)
static LogScale64 log_scale(double k_ln2)
{
    LogScale64 s;
    for (int i = 0; i < 64; i++)
        s.fill(i, k_ln2);
    return s;
}

extern "C" {
void hm_sincos20(const uint32_t *k, double *cs, double *sn, long n)
{
    for (long i = 0; i < n; i++) sincos_turn20(k[i], cs[i], sn[i], tables());
}
void hm_neg2log(const double *u, double *out, long n)
{
    const LogScale64 s = log_scale(-2.0 * 0x1.62e42fefa39efp-1);
    for (long i = 0; i < n; i++) out[i] = neg2log_unit(u[i], tables(), s);
}
void hm_scaled_log(const double *u, double k, double k_ln2, double *out, long n)
{
    const LogScale64 s = log_scale(k_ln2);
    for (long i = 0; i < n; i++) out[i] = scaled_log_unit(u[i], tables(), k, s);
}
void hm_exp_units(const double *y, double *out, long n) { for (long i = 0; i < n; i++) out[i] = exp_units(y[i], tables()); }
void hm_exp_units_late(const double *y, double *out, long n) { for (long i = 0; i < n; i++) out[i] = exp_units<true>(y[i], tables()); }
void hm_sqrt(const double *x, double *out, long n) { for (long i = 0; i < n; i++) out[i] = sqrt_pos(x[i]); }
void hm_sqrt_short(const double *x, double *out, long n) { for (long i = 0; i < n; i++) out[i] = sqrt_pos<true>(x[i]); }
void hm_rcp(const double *x, double *out, long n) { for (long i = 0; i < n; i++) out[i] = rcp_newton(x[i]); }
void hm_exp(const double *x, double *out, long n) { for (long i = 0; i < n; i++) out[i] = exp_tab(x[i], tables()); }
void hm_sincos(const uint32_t *k_hi, const uint32_t *k_lo, double *cs, double *sn, long n)
{
    for (long i = 0; i < n; i++) sincos_turn(k_hi[i], k_lo[i], cs[i], sn[i]);
}
}
