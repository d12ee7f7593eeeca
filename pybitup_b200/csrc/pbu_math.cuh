// Lean double-precision exp and log for the ODE forward models (host + device, so that they can be checked on the
// CPU: tools/test_math.cu).  CUDA's exp()/log() cover the full IEEE domain with special cases; the Arrhenius model only
// needs exp on [-700, 700] and log on normal numbers in (0, 2), which takes 17 and 22 FP64 instructions instead of
// about 65 each, and the model is nothing but six of these per RK4 step.
//   pbu_exp : x = k ln2 + r, |r| <= ln2/2;  exp(r) = 1 + r + r^2 P9(r)  (near-minimax, 1.7e-17 relative), scaled by 2^k
//   pbu_log : u = 2^e m, m in [sqrt(1/2), sqrt(2));  s = (m-1)/(m+1);  log m = 2s + s^3 Q6(s^2)  (4.5e-18 relative)
// Measured against libm on 2e7 random arguments (tools/test_math.cu): max error < 1 ulp for both.
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>
#include "pbu_math_tables.h"

#ifdef __CUDACC__
#define PBU_HD __host__ __device__ __forceinline__
#else
#define PBU_HD inline
#endif

namespace pbu {

PBU_HD int32_t hi_word(double x) {
#ifdef __CUDA_ARCH__
    return __double2hiint(x);
#else
    uint64_t b;
    memcpy(&b, &x, 8);
    return (int32_t)(b >> 32);
#endif
}
PBU_HD double with_hi_word(double x, int32_t hi) {
#ifdef __CUDA_ARCH__
    return __hiloint2double(hi, __double2loint(x));
#else
    uint64_t b;
    memcpy(&b, &x, 8);
    b = (b & 0xffffffffull) | ((uint64_t)(uint32_t)hi << 32);
    memcpy(&x, &b, 8);
    return x;
#endif
}

// exp(x) for -700 <= x <= 700 (0 below); no inf / nan handling
PBU_HD double pbu_exp(double x) {
    const bool tiny = x < -700.0;                                // branch-free: the result is selected at the end
    x = tiny ? 0.0 : x;
    const double SHIFT = 6755399441055744.0;                     // 1.5 * 2^52: rounds to nearest integer in the low bits
    const double t = fma(x, 1.4426950408889634, SHIFT);
    const double kf = t - SHIFT;
    const int32_t k = (int32_t)(uint32_t)(
#ifdef __CUDA_ARCH__
        __double2loint(t)
#else
        [](double v) { uint64_t b; memcpy(&b, &v, 8); return (int32_t)(uint32_t)b; }(t)
#endif
    );
    double r = fma(kf, -0.6931471803691238, x);                  // ln2 split: hi has 32 significant bits, kf*hi exact
    r = fma(kf, -1.9082149292705877e-10, r);
    double p = 2.5100375832561234e-08;
    p = fma(p, r, 2.7620075879983367e-07);
    p = fma(p, r, 2.7557268480310024e-06);
    p = fma(p, r, 2.4801521322368692e-05);
    p = fma(p, r, 0.00019841269863040545);
    p = fma(p, r, 0.0013888888917196719);
    p = fma(p, r, 0.008333333333330065);
    p = fma(p, r, 0.041666666666624164);
    p = fma(p, r, 0.16666666666666669);
    p = fma(p, r, 0.5000000000000001);
    const double v = 1.0 + fma(r * r, p, r);
    const double res = with_hi_word(v, hi_word(v) + (k << 20));   // * 2^k on the exponent field (k >= -1010 here)
    return tiny ? 0.0 : res;
}

// log(u) for normal u > 0 (no zero / denormal / inf / nan handling)
PBU_HD double pbu_log(double u) {
    int32_t hi = hi_word(u);
    int32_t e = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;                          // m in [1, 2)
    // m in [sqrt(1/2), sqrt(2)): halve the mantissas above sqrt(2) by lowering the exponent field (branch-free);
    // 0x6a09e is the top 20 mantissa bits of sqrt(2), so the cut sits within 2^-20 of it, harmless for the series
    const int32_t big = ((hi & 0x000fffff) > 0x6a09e) ? 1 : 0;
    hi -= big << 20;
    e += big;
    const double m = with_hi_word(u, hi);
    const double f = m - 1.0, g = m + 1.0;
    // s = f / g: reciprocal seed refined by two Newton steps, then one residual correction of the quotient
#ifdef __CUDA_ARCH__
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(g));
#else
    double y = (double)(1.0f / (float)g);
#endif
    double err = fma(-g, y, 1.0);
    y = fma(y, err, y);
    err = fma(-g, y, 1.0);
    y = fma(y, err, y);
    double s = f * y;
    s = fma(fma(-s, g, f), y, s);
    const double z = s * s;
    double q = 0.14616585424888623;
    q = fma(q, z, 0.15331710618210773);
    q = fma(q, z, 0.18182889455674947);
    q = fma(q, z, 0.22222211130259878);
    q = fma(q, z, 0.2857142862600327);
    q = fma(q, z, 0.39999999999899444);
    q = fma(q, z, 0.666666666666667);
    const double ef = (double)e;
    // e ln2_hi + (2s + (s z q + e ln2_lo)): e*ln2_hi is exact (32-bit hi, |e| < 2^11)
    const double lo = fma(s * z, q, ef * 1.9082149292705877e-10);
    return fma(ef, 0.6931471803691238, fma(2.0, s, lo));
}

// ------------------------------------------------------------------------------------------------
// Table-driven base-2 variants: 2^y in 8 and log2(u) in 7 FP64 instructions (against 17 and 22 above), paid for with
// one shared-memory load each.  The FP64 pipe is what bounds the ODE models, the load/store pipe is idle there.  Base 2
// makes both argument reductions exact and removes the ln2 hi/lo bookkeeping; u^n = 2^(n log2 u) needs nothing else.
//   pbu_exp2_tab : y = k + j/512 + r, |r| <= 2^-10 (r = y - (512k+j)/512 EXACTLY, one FMA);
//                  2^y = 2^k T[j] (1 + r (b1 + b2 r + b3 r^2 + b4 r^3)),  T[j] = 2^(j/512),  b_i = ln2^i / i!
//   pbu_log2_tab : u = 2^e m, m in [1, 2), i = top 8 mantissa bits;  r = m c_i - 1, |r| <= 2^-9 (one FMA);
//                  log2 u = e + (-log2 c_i + r (a1 + a2 r + .. + a5 r^4)),  a_i = (-1)^(i+1) / (i ln2)
// Tables: pbu_math_tables.h (generated by tools/gen_math_tables.py), 512 + 2*256 doubles (8 KB).  On the device the kernels
// stage them in shared memory once per block (Model::block_init), on a 4 KB boundary of the shared window, and pass the
// pointer; on the host the static copy.
// Checked against libm in tools/test_math.cu: exp2 <= 1 ulp; log2 <= 1 ulp of max(|log2 u|, 1) -- the error that
// matters for 2^(n log2 u) is the absolute one.
#define PBU_MATH_TABLE_DOUBLES (512 + 512)
#define PBU_LOG2_TABLE_OFFSET 512
static const double pbu_math_tables_host[PBU_MATH_TABLE_DOUBLES] = {PBU_EXP2_TABLE, PBU_LOG2_TABLE};
#ifdef __CUDACC__
static __device__ const double pbu_math_tables_dev[PBU_MATH_TABLE_DOUBLES] = {PBU_EXP2_TABLE, PBU_LOG2_TABLE};
#endif

// Polynomial coefficients.  On the device they are read as constant-bank operands of the DFMAs (a __constant__ array
// the compiler cannot fold): written as literals ptxas rebuilds every 64-bit constant with two UMOVs inside the loop,
// about 30 instruction-issue slots per RK4 step, and on this kernel every issued instruction is a cycle the FP64 pipe
// idles (an FP64 instruction holds the issue port for 2 cycles, everything else for 1; ncu: issue slots add up to the
// elapsed cycles).
#define PBU_K_B4 9.6181291076284772e-03                          // ln2^4/24
#define PBU_K_B3 5.5504108664821580e-02                          // ln2^3/6
#define PBU_K_B2 2.4022650695910072e-01                          // ln2^2/2
#define PBU_K_B1 6.9314718055994531e-01                          // ln2
#define PBU_K_A5 2.8853900817779266e-01                          // 1/(5 ln2)
#define PBU_K_A4 -3.6067376022224085e-01                         // -1/(4 ln2)
#define PBU_K_A3 4.8089834696298783e-01                          // 1/(3 ln2)
#define PBU_K_A2 -7.2134752044448170e-01                         // -1/(2 ln2)
#define PBU_K_A1 1.4426950408889634e+00                          // 1/ln2
#ifdef __CUDACC__
static __constant__ double pbu_kc[9] = {PBU_K_B4, PBU_K_B3, PBU_K_B2, PBU_K_B1, PBU_K_A5, PBU_K_A4, PBU_K_A3, PBU_K_A2, PBU_K_A1};
#endif
#ifdef __CUDA_ARCH__
#define PBU_KC(i, lit) pbu_kc[i]
#else
#define PBU_KC(i, lit) (lit)
#endif

#ifdef __CUDA_ARCH__
// 32-bit shared-window address of the tables, which the caller has placed on a 4 KB boundary of that window
// (ArrheniusModel::tables).  The mov hides the known-zero low bits from the front end, which would otherwise turn
// base | offset back into an add; ptxas folds the mov and merges mask and OR into one LOP3.
__device__ __forceinline__ uint32_t pbu_tab_base(const double *tab) {
    uint32_t b;
    asm("mov.u32 %0, %1;" : "=r"(b) : "r"((uint32_t)__cvta_generic_to_shared(tab)));
    return b;
}
#endif

// 2^y, hi word of the result forced to 0 where !keep (the caller's own range test rides on the selection that is
// already there).  tab: the PBU_MATH_TABLE_DOUBLES doubles above (exp2 table first).  Domain y < 1024; y <= -1022 (and
// -inf) returns a denormal (< 2.3e-308) instead of the exact tail: the hi word of y is clamped as an unsigned integer
// (one instruction; negative doubles order by magnitude), which lands in k = -1022 or -1023 and leaves at most the
// mantissa bits.  The FP64 pipe sees the eight arithmetic instructions only.  NaN with the sign bit clear propagates.
PBU_HD double pbu_exp2_tab_sel(double y, const double *tab, bool keep) {
    const uint32_t yh = (uint32_t)hi_word(y);
    y = with_hi_word(y, (int32_t)(yh < 0xC08FF000u ? yh : 0xC08FF000u));        // y >= -1022.001
    const double SHIFT = 6755399441055744.0;                     // 1.5 * 2^52
    const double t = fma(y, 512.0, SHIFT);
    const double kf = t - SHIFT;                                 // 512 k + j as a double
    const int32_t m = (int32_t)(uint32_t)(
#ifdef __CUDA_ARCH__
        __double2loint(t)
#else
        [](double v) { uint64_t b; memcpy(&b, &v, 8); return (int32_t)(uint32_t)b; }(t)
#endif
    );
    const double r = fma(kf, -0.001953125, y);                   // exact
    double p = fma(r, PBU_KC(0, PBU_K_B4), PBU_KC(1, PBU_K_B3));
    p = fma(p, r, PBU_KC(2, PBU_K_B2));
    p = fma(p, r, PBU_KC(3, PBU_K_B1));
    const double v = p * r;                                      // 2^r - 1, remainder (r ln2)^5/120 < 1.3e-18
#ifdef __CUDA_ARCH__
    double T;                                                    // tab is 4 KB aligned in the shared window: base | offset
    asm("ld.shared.f64 %0, [%1];" : "=d"(T) : "r"(pbu_tab_base(tab) | (((uint32_t)m << 3) & 0xFF8u)));      // is one LOP3
#else
    const double T = tab[m & 511];
#endif
    const double e = fma(T, v, T);
    int32_t rh;                                                  // * 2^k on the exponent field (k >= -1023 here)
#ifdef __CUDA_ARCH__
    asm("{\n\t.reg .s32 k;\n\tshr.s32 k, %1, 9;\n\tmad.lo.s32 %0, k, 1048576, %2;\n\t}" : "=r"(rh) : "r"(m), "r"(hi_word(e)));
#else
    rh = hi_word(e) + (m >> 9) * 1048576;
#endif
    return with_hi_word(e, keep ? rh : 0);
}
PBU_HD double pbu_exp2_tab(double y, const double *tab) { return pbu_exp2_tab_sel(y, tab, true); }

// true for positive, normal, finite u: the arguments pbu_log2_tab accepts (an integer test)
PBU_HD bool pbu_log_arg_ok(double u) { return (uint32_t)(hi_word(u) - 0x00100000) < 0x7FE00000u; }

PBU_HD double pbu_log2_tab(double u, const double *tab) {
    const int32_t hi = hi_word(u);
    const int32_t e = (hi >> 20) - 1023;
    const double m = with_hi_word(u, (hi & 0x000FFFFF) | 0x3FF00000);
#ifdef __CUDA_ARCH__
    double c, nlc;                                               // one 16-byte shared-memory load, address as in exp2
    asm("ld.shared.v2.f64 {%0, %1}, [%2+4096];" : "=d"(c), "=d"(nlc)                 // 4096 = 8 * PBU_LOG2_TABLE_OFFSET
        : "r"(pbu_tab_base(tab) | (((uint32_t)hi >> 8) & 0xFF0u)));
#else
    const int32_t i = (hi >> 12) & 255;
    const double c = tab[PBU_LOG2_TABLE_OFFSET + 2 * i], nlc = tab[PBU_LOG2_TABLE_OFFSET + 2 * i + 1];
#endif
    const double r = fma(m, c, -1.0);
    double q = fma(r, PBU_KC(4, PBU_K_A5), PBU_KC(5, PBU_K_A4));     // remainder r^6 / (6 ln2) < 1.4e-17
    q = fma(q, r, PBU_KC(6, PBU_K_A3));
    q = fma(q, r, PBU_KC(7, PBU_K_A2));
    q = fma(q, r, PBU_KC(8, PBU_K_A1));
    return fma(r, q, nlc) + (double)e;                           // small part first: one rounding at full size
}

}  // namespace pbu
