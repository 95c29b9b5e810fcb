/*
 * ref_shim.c -- seed control for the reference CPU path built into oracle/_ref/.
 * TEST INFRASTRUCTURE ONLY (see mc_oracle.h).
 *
 * The reference host estimators call srand((unsigned)time(NULL)) on every call
 * (DP/MonteCarloHost.c:189,237), so two runs never agree.  The _ref libraries are linked
 * with -Wl,--wrap=time, which routes those calls here; ref_set_time() pins the value.
 * Nothing else of the reference is altered.
 */
#include <time.h>

static time_t g_pinned = 20180206; /* the reference's creation date, an arbitrary constant */
static int g_auto_advance = 0;

time_t __wrap_time(time_t *out)
{
    time_t v = g_pinned;
    if (g_auto_advance)
        g_pinned++;
    if (out)
        *out = v;
    return v;
}

void ref_set_time(long value, int auto_advance)
{
    g_pinned = (time_t)value;
    g_auto_advance = auto_advance;
}
