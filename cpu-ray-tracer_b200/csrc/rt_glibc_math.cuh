// glibc 2.39 (x86-64) single-precision expf / acosf / atanf / atan2f restated so that host and device
// return the SAME BITS as the libm the reference's CPU build links against on this image.
//
// Why: the reference's integrators call exp() on floats for Beer's law (renderer.cpp:76-80) and atan2f / acosf for the
// skydome lookup (file_scene.cpp:142-154).  CUDA's expf / atan2f / acosf differ from glibc's by up to 2 ulp, which moved
// a few radiance values in the last bits and, rarely, a sky texel.  With these restatements the device computes
// every arithmetic step of those routines in the same order and precision as the host library.
//
// Algorithms (glibc is not vendored under /root/reference: it is the system libm, Ubuntu GLIBC 2.39-0ubuntu8.5):
//   expf   sysdeps/ieee754/flt-32/e_expf.c (Szabolcs Nagy's exp2f-table routine, N = 32, double arithmetic), in the
//          contraction pattern of the ifunc variant __expf_fma (sysdeps/x86_64/fpu/multiarch/e_expf-fma.c) that x86-64
//          hosts with FMA + AVX2 select: read off the disassembly of libm.so.6 (five fused operations, marked below).
//   acosf  sysdeps/ieee754/flt-32/e_acosf.c   (fdlibm, float arithmetic, no contraction)
//   atanf  sysdeps/ieee754/flt-32/s_atanf.c   (fdlibm)
//   atan2f sysdeps/ieee754/flt-32/e_atan2f.c  (fdlibm)
// Pinned by tests/test_glibc_math.py: compiled for the host and compared with libm bit for bit (exhaustively over all
// 2^32 arguments for expf and acosf and atanf with RT_GLIBC_MATH_EXHAUSTIVE=1, strided otherwise; atan2f on a lattice
// of argument pairs plus the directions the sky lookup produces).
//
// The file is plain C/C++: every operation is written out; the builds use -fmad=false (device) and -ffp-contract=off
// (host), so nothing is fused except the explicit rt_gm_fma calls.
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define RT_GM_FN __host__ __device__ __forceinline__
#else
#define RT_GM_FN static inline
#endif

#if defined(__CUDA_ARCH__)
#define RT_GM_TABLE static __device__ const
RT_GM_FN double rt_gm_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
RT_GM_FN float rt_gm_sqrtf(float x) { return __fsqrt_rn(x); }
RT_GM_FN float rt_gm_fmaf(float a, float b, float c) { return __fmaf_rn(a, b, c); }
RT_GM_FN uint32_t rt_gm_f2u(float f) { return __float_as_uint(f); }
RT_GM_FN float rt_gm_u2f(uint32_t u) { return __uint_as_float(u); }
RT_GM_FN uint64_t rt_gm_d2u(double d) { return (uint64_t)__double_as_longlong(d); }
RT_GM_FN double rt_gm_u2d(uint64_t u) { return __longlong_as_double((long long)u); }
#else
#define RT_GM_TABLE static const
RT_GM_FN double rt_gm_fma(double a, double b, double c) { return __builtin_fma(a, b, c); }
RT_GM_FN float rt_gm_sqrtf(float x) { return __builtin_sqrtf(x); }
RT_GM_FN float rt_gm_fmaf(float a, float b, float c) { return __builtin_fmaf(a, b, c); }
RT_GM_FN uint32_t rt_gm_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
RT_GM_FN float rt_gm_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
RT_GM_FN uint64_t rt_gm_d2u(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
RT_GM_FN double rt_gm_u2d(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
#endif

// __exp2f_data.tab: bits(2^(i/32)) - (i << 47), i = 0..31 (regenerated from the definition and compared with libm's .rodata)
RT_GM_TABLE uint64_t rt_gm_exp2f_tab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};

RT_GM_FN float rt_glibc_expf(float x)
{
    const uint32_t ux = rt_gm_f2u(x);
    const uint32_t abstop = (ux >> 20) & 0x7ff;
    if (abstop >= 0x42b) // |x| >= 88 or NaN
    {
        if (ux == 0xff800000u) return 0.0f;
        if (abstop >= 0x7f8) return x + x;
        if (x > 0x1.62e42ep6f) return rt_gm_u2f(0x7f800000u); // __math_oflowf
        if (x < -0x1.9fe368p6f) return 0.0f;                  // __math_uflowf
        if (x < -0x1.9d1d9ep6f) return rt_gm_u2f(1u);         // __math_may_uflowf: 0x1.4p-75f * 0x1.4p-75f
    }
    const double xd = (double)x;
    const double InvLn2N = 0x1.71547652b82fep+5, Shift = 0x1.8p+52;
    double kd = rt_gm_fma(InvLn2N, xd, Shift);       // fused in __expf_fma
    const uint64_t ki = rt_gm_d2u(kd);
    kd = kd - Shift;
    const double r = rt_gm_fma(InvLn2N, xd, -kd);    // fused
    const uint64_t t = rt_gm_exp2f_tab[ki & 31] + (ki << 47);
    const double s = rt_gm_u2d(t);
    const double z = rt_gm_fma(0x1.c6af84b912394p-20, r, 0x1.ebfce50fac4f3p-13); // fused
    const double r2 = r * r;
    double y = rt_gm_fma(0x1.62e42ff0c52d6p-6, r, 1.0); // fused
    y = rt_gm_fma(z, r2, y);                            // fused
    y = y * s;
    return (float)y;
}

// acosf / atanf / atan2f below keep glibc's arithmetic (each result is produced by the same sequence of float operations
// as in e_acosf.c / s_atanf.c / e_atan2f.c) but the range cases of the C sources are folded into selects around ONE
// polynomial and ONE division, so that the lanes of a warp that look up the sky in different octants do not serialise
// (the straight transcription compiled to 744 SASS instructions with 16 division sites).
RT_GM_FN float rt_glibc_acosf(float x)
{
    const float one = 1.0f, pi = rt_gm_u2f(0x40490fdau), pio2_hi = rt_gm_u2f(0x3fc90fdau), pio2_lo = rt_gm_u2f(0x33a22168u);
    const float pS0 = rt_gm_u2f(0x3e2aaaabu), pS1 = rt_gm_u2f(0xbea6b090u), pS2 = rt_gm_u2f(0x3e4e0aa8u), pS3 = rt_gm_u2f(0xbd241146u),
                pS4 = rt_gm_u2f(0x3a4f7f04u), pS5 = rt_gm_u2f(0x3811ef08u), qS1 = rt_gm_u2f(0xc019d139u), qS2 = rt_gm_u2f(0x4001572du),
                qS3 = rt_gm_u2f(0xbf303361u), qS4 = rt_gm_u2f(0x3d9dc62eu);
    const int32_t hx = (int32_t)rt_gm_f2u(x);
    const int32_t ix = hx & 0x7fffffff;
    if (ix >= 0x3f800000) // |x| >= 1 or NaN
    {
        if (ix == 0x3f800000) return hx > 0 ? 0.0f : pi + 2.0f * pio2_lo;
        return rt_gm_u2f(0x7fc00000u); // (x - x) / (x - x): NaN
    }
    if (ix <= 0x32800000) return pio2_hi + pio2_lo; // |x| <= 2^-26
    const int small = ix < 0x3f000000;              // |x| < 0.5
    const float ax = rt_gm_u2f((uint32_t)ix);
    const float z = small ? x * x : (one - ax) * 0.5f; // x < -0.5: (one + x) * 0.5 is the same operation
    const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    const float q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    const float r = p / q;
    const float s = rt_gm_sqrtf(z);
    const float res_small = pio2_hi - (x - (pio2_lo - x * r));
    const float res_neg = pi - 2.0f * (s + (r * s - pio2_lo));
    if (!small && hx > 0) // x > 0.5: the one case with a second division (directions below the horizon for the sky lookup)
    {
        const float df = rt_gm_u2f(rt_gm_f2u(s) & 0xfffff000u);
        const float c = (z - df * df) / (s + df);
        return 2.0f * (df + (r * s + c));
    }
    return small ? res_small : res_neg;
}

RT_GM_FN float rt_glibc_atanf(float x)
{
    const float aT0 = rt_gm_u2f(0x3eaaaaabu), aT1 = rt_gm_u2f(0xbe4ccccdu), aT2 = rt_gm_u2f(0x3e124925u), aT3 = rt_gm_u2f(0xbde38e38u),
                aT4 = rt_gm_u2f(0x3dba2e6eu), aT5 = rt_gm_u2f(0xbd9d8795u), aT6 = rt_gm_u2f(0x3d886b35u), aT7 = rt_gm_u2f(0xbd6ef16bu),
                aT8 = rt_gm_u2f(0x3d4bda59u), aT9 = rt_gm_u2f(0xbd15a221u), aT10 = rt_gm_u2f(0x3c8569d7u);
    const float one = 1.0f;
    const int32_t hx = (int32_t)rt_gm_f2u(x);
    const int32_t ix = hx & 0x7fffffff;
    if (ix >= 0x4c000000) // |x| >= 2^25
    {
        if (ix > 0x7f800000) return x + x;
        const float h = rt_gm_u2f(0x3fc90fdau), l = rt_gm_u2f(0x33a22168u);
        return hx > 0 ? h + l : -h - l;
    }
    if (ix < 0x31000000) return x; // |x| < 2^-29
    // argument reduction: id = -1 (|x| < 7/16, x itself: x / 1 is exact), 0 (< 11/16), 1 (< 19/16), 2 (< 39/16), 3
    const float ax = rt_gm_u2f((uint32_t)ix);
    const int c0 = ix < 0x3ee00000, c1 = ix < 0x3f300000, c2 = ix < 0x3f980000, c3 = ix < 0x401c0000;
    const float num = c0 ? x : c1 ? 2.0f * ax - one : c2 ? ax - one : c3 ? ax - 1.5f : -1.0f;
    const float den = c0 ? one : c1 ? 2.0f + ax : c2 ? ax + one : c3 ? one + 1.5f * ax : ax;
    const float hi = rt_gm_u2f(c1 ? 0x3eed6338u : c2 ? 0x3f490fdau : c3 ? 0x3f7b985eu : 0x3fc90fdau);
    const float lo = rt_gm_u2f(c1 ? 0x31ac3769u : c2 ? 0x33222168u : c3 ? 0x33140fb4u : 0x33a22168u);
    const float xr = num / den;
    const float z = xr * xr;
    const float w = z * z;
    const float s1 = z * (aT0 + w * (aT2 + w * (aT4 + w * (aT6 + w * (aT8 + w * aT10)))));
    const float s2 = w * (aT1 + w * (aT3 + w * (aT5 + w * (aT7 + w * aT9))));
    const float t = xr * (s1 + s2);
    const float zz = hi - ((t - lo) - xr);
    return c0 ? xr - t : hx < 0 ? -zz : zz;
}

RT_GM_FN float rt_glibc_atan2f(float y, float x)
{
    const float tiny = 1.0e-30f, pi_o_4 = rt_gm_u2f(0x3f490fdbu), pi_o_2 = rt_gm_u2f(0x3fc90fdbu), pi = rt_gm_u2f(0x40490fdbu),
                pi_lo = rt_gm_u2f(0xb3bbbd2eu);
    const int32_t hx = (int32_t)rt_gm_f2u(x), hy = (int32_t)rt_gm_f2u(y);
    const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2); // 2 * sign(x) + sign(y)
    // zeros, infinities and NaNs: exactly the cases of e_atan2f.c, in its order.  Its x == 1 shortcut (return atanf(y)) needs
    // no branch: y / 1 is y, atanf is odd operation by operation, and m is 0 or 1 there, so the general path returns the same bits
    // once the |y / x| > 2^60 constant is kept away from it (checked by the lattice of tests/tools/glibc_math_check.c).
    if (ix > 0x7f800000 || iy > 0x7f800000 || iy == 0 || ix == 0 || ix == 0x7f800000 || iy == 0x7f800000)
    {
        if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
        if (iy == 0) return m < 2 ? y : m == 2 ? pi + tiny : -pi - tiny;
        if (ix == 0) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
        if (ix == 0x7f800000)
        {
            if (iy == 0x7f800000) return m == 0 ? pi_o_4 + tiny : m == 1 ? -pi_o_4 - tiny : m == 2 ? 3.0f * pi_o_4 + tiny : -3.0f * pi_o_4 - tiny;
            return m == 0 ? 0.0f : m == 1 ? -0.0f : m == 2 ? pi + tiny : -pi - tiny;
        }
        return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
    }
    const int32_t k = (iy - ix) >> 23;
    float z;
    if (k > 60 && hx != 0x3f800000) z = pi_o_2 + 0.5f * pi_lo;
    else if (hx < 0 && k < -60) z = 0.0f;
    else z = rt_glibc_atanf(rt_gm_u2f(rt_gm_f2u(y / x) & 0x7fffffffu));
    const float zl = z - pi_lo;
    return m == 0 ? z : m == 1 ? rt_gm_u2f(rt_gm_f2u(z) ^ 0x80000000u) : m == 2 ? pi - zl : zl - pi;
}
