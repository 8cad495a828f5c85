// Compares cpu-ray-tracer_b200/csrc/rt_glibc_math.cuh (compiled for the host) with the system libm bit for bit.
//   glibc_math_check STRIDE PAIRS
// expf / acosf / atanf: every STRIDE-th of the 2^32 float bit patterns (STRIDE = 1 is exhaustive);
// atan2f: a lattice of special values, PAIRS pseudo-random bit-pattern pairs, PAIRS pairs of moderate magnitude and PAIRS
// components (-D.z, D.x) of normalised directions (what GetSkyColor passes, file_scene.cpp:142-154).
// NaN results compare equal when both are NaN.  Prints one line per function; exit code 1 on any mismatch.
#include <stdio.h>
#include <stdlib.h>
#include "rt_glibc_math.cuh"

static int same(float a, float b) { return rt_gm_f2u(a) == rt_gm_f2u(b) || (a != a && b != b); }
static uint64_t rng(uint64_t* s) { *s ^= *s << 13; *s ^= *s >> 7; *s ^= *s << 17; return *s; }

#define CHECK1(NAME, MINE, LIBM) \
    static long check_##NAME(uint64_t stride) { \
        long bad = 0; \
        _Pragma("omp parallel for reduction(+:bad) schedule(static)") \
        for (uint64_t i = 0; i < (1ull << 32); i += stride) { \
            const float x = rt_gm_u2f((uint32_t)i); \
            const float a = MINE(x), b = LIBM(x); \
            if (!same(a, b)) { if (bad < 5) fprintf(stderr, #NAME "(%a) = %a, libm %a\n", x, a, b); bad++; } \
        } \
        return bad; }
CHECK1(expf, rt_glibc_expf, expf)
CHECK1(acosf, rt_glibc_acosf, acosf)
CHECK1(atanf, rt_glibc_atanf, atanf)

static long pair(float y, float x)
{
    const float a = rt_glibc_atan2f(y, x), b = atan2f(y, x);
    if (same(a, b)) return 0;
    fprintf(stderr, "atan2f(%a, %a) = %a, libm %a\n", y, x, a, b);
    return 1;
}

int main(int argc, char** argv)
{
    const uint64_t stride = argc > 1 ? strtoull(argv[1], 0, 10) : 4099;
    const long pairs = argc > 2 ? atol(argv[2]) : 2000000;
    long total = 0, bad;
    bad = check_expf(stride), total += bad, printf("expf   stride %llu mismatches %ld\n", (unsigned long long)stride, bad);
    bad = check_acosf(stride), total += bad, printf("acosf  stride %llu mismatches %ld\n", (unsigned long long)stride, bad);
    bad = check_atanf(stride), total += bad, printf("atanf  stride %llu mismatches %ld\n", (unsigned long long)stride, bad);
    static const uint32_t special[] = { 0x00000000u, 0x80000000u, 0x00000001u, 0x80000001u, 0x007fffffu, 0x00800000u, 0x3f800000u, 0xbf800000u,
        0x3f000000u, 0x3effffffu, 0x3ee00000u, 0x3f300000u, 0x3f980000u, 0x401c0000u, 0x4c000000u, 0x4bffffffu, 0x31000000u, 0x30ffffffu,
        0x7f7fffffu, 0xff7fffffu, 0x7f800000u, 0xff800000u, 0x7fc00000u, 0xffc00000u, 0x5e800000u, 0xde800000u, 0x4c000001u, 0x30800000u, 0xb0800000u, 0x3e000000u, 0xbf000000u, 0x40000000u, 0xc0400000u, 0x21000000u, 0x40490fdbu, 0x3fc90fdbu };
    const int ns = sizeof(special) / sizeof(special[0]);
    bad = 0;
    for (int i = 0; i < ns; i++) for (int j = 0; j < ns; j++) bad += pair(rt_gm_u2f(special[i]), rt_gm_u2f(special[j]));
    long bad2 = 0;
    #pragma omp parallel for reduction(+:bad2) schedule(static)
    for (long i = 0; i < pairs; i++)
    {
        uint64_t s = 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1);
        rng(&s);
        const uint64_t r = rng(&s);
        bad2 += pair(rt_gm_u2f((uint32_t)r), rt_gm_u2f((uint32_t)(r >> 32)));                       // any bit patterns
        const uint64_t q = rng(&s);                                                                // exponents within 2^-8 .. 2^8
        bad2 += pair(rt_gm_u2f(((uint32_t)q & 0x87ffffffu) | 0x3b800000u), rt_gm_u2f(((uint32_t)(q >> 32) & 0x87ffffffu) | 0x3b800000u));
        float d[3];                                                                                // unit directions
        for (int k = 0; k < 3; k++) d[k] = (float)(rng(&s) >> 40) * (1.0f / 8388608.0f) - 1.0f;
        const float inv = 1.0f / sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        bad2 += pair(-(d[2] * inv), d[0] * inv);
        const float ya = rt_glibc_acosf(-(d[1] * inv)), yb = acosf(-(d[1] * inv));
        if (!same(ya, yb)) bad2++;
    }
    bad += bad2, total += bad;
    printf("atan2f lattice %d x %d + 3 x %ld pairs mismatches %ld\n", ns, ns, pairs, bad);
    return total ? 1 : 0;
}

// the same file built as a shared library (tests/test_glibc_math.py): the host libm over arrays, for the device comparison
void libm_eval(int fn, const float* a, const float* b, float* out, size_t n)
{
    for (size_t i = 0; i < n; i++) out[i] = fn == 0 ? expf(a[i]) : fn == 1 ? acosf(a[i]) : atan2f(a[i], b[i]);
}
void restated_eval(int fn, const float* a, const float* b, float* out, size_t n)
{
    for (size_t i = 0; i < n; i++) out[i] = fn == 0 ? rt_glibc_expf(a[i]) : fn == 1 ? rt_glibc_acosf(a[i]) : rt_glibc_atan2f(a[i], b[i]);
}
