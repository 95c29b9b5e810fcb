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
        std::memcpy(g_tables.log_tab, kLogTable, sizeof kLogTable);
        std::memcpy(g_tables.exp_tab, kExpTable, sizeof kExpTable);
        std::memcpy(g_tables.turn_hi, kTurnHiTable, sizeof kTurnHiTable);
        std::memcpy(g_tables.turn_lo, kTurnLoTable, sizeof kTurnLoTable);
        g_ready = true;
    }
    return g_tables;
}

extern "C" {
void hm_sincos20(const uint32_t *k, double *cs, double *sn, long n)
{
    for (long i = 0; i < n; i++) sincos_turn20(k[i], cs[i], sn[i], tables());
}
void hm_neg2log(const double *u, double *out, long n) { for (long i = 0; i < n; i++) out[i] = neg2log_unit(u[i], tables()); }
void hm_scaled_log(const double *u, double k, double k_ln2, double *out, long n)
{
    for (long i = 0; i < n; i++) out[i] = scaled_log_unit(u[i], tables(), k, k_ln2);
}
void hm_sqrt(const double *x, double *out, long n) { for (long i = 0; i < n; i++) out[i] = sqrt_pos(x[i]); }
void hm_sqrt_short(const double *x, double *out, long n) { for (long i = 0; i < n; i++) out[i] = sqrt_pos<true>(x[i]); }
void hm_rcp(const double *x, double *out, long n) { for (long i = 0; i < n; i++) out[i] = rcp_newton(x[i]); }
void hm_exp(const double *x, double *out, long n) { for (long i = 0; i < n; i++) out[i] = exp_tab(x[i], tables()); }
void hm_sincos(const uint32_t *k_hi, const uint32_t *k_lo, double *cs, double *sn, long n)
{
    for (long i = 0; i < n; i++) sincos_turn(k_hi[i], k_lo[i], cs[i], sn[i]);
}
}
