/* oracle/fixtime.c -- TEST INFRASTRUCTURE ONLY.
 *
 * LD_PRELOAD shim that pins time() so that the reference's `srand(time(NULL))`
 * (reference src/libfm/libfm.cpp:123-124; `-seed` is ignored there) becomes a
 * fixed, caller-chosen seed:   FAKE_TIME=42 LD_PRELOAD=oracle/_ref/fixtime.so libFM ...
 */
#include <stdlib.h>
#include <time.h>

time_t time(time_t *t) {
    const char *s = getenv("FAKE_TIME");
    time_t v = s ? (time_t)atol(s) : (time_t)42;
    if (t) *t = v;
    return v;
}
