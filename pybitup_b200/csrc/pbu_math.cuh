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

// exp(x) over the whole domain, for the acceptance ratios (mha.py:200-207, :547-556): +inf from 708 (numpy overflows to
// inf from 709.78: alpha = 1 either way), 0 below -708 (numpy: 3e-308 and denormals -- below any uniform and
// invisible in a sum), NaN for NaN.  The degree-10 polynomial of pbu_exp with its coefficients read by the FMAs from the constant bank: CUDA's
// exp() builds each of its constants with two UMOVs inside the caller's loop, 22 instruction-issue slots per call.
#ifdef __CUDACC__
static __constant__ double pbu_exp_kc[14] = {1.4426950408889634, -0.6931471803691238, -1.9082149292705877e-10, 2.5100375832561234e-08,
                                              2.7620075879983367e-07, 2.7557268480310024e-06, 2.4801521322368692e-05, 0.00019841269863040545,
                                              0.0013888888917196719, 0.008333333333330065, 0.041666666666624164, 0.16666666666666669,
                                              0.5000000000000001, 0.0};
#endif
#ifdef __CUDA_ARCH__
#define PBU_EXP_KC(i, lit) pbu_exp_kc[i]
#else
#define PBU_EXP_KC(i, lit) (lit)
#endif
PBU_HD double pbu_exp_full(double x) {
    const double SHIFT = 6755399441055744.0;                     // 1.5 * 2^52
    const double t = fma(x, PBU_EXP_KC(0, 1.4426950408889634), SHIFT);
    const double kf = t - SHIFT;
    const int32_t k = (int32_t)(uint32_t)(
#ifdef __CUDA_ARCH__
        __double2loint(t)
#else
        [](double v) { uint64_t b; memcpy(&b, &v, 8); return (int32_t)(uint32_t)b; }(t)
#endif
    );
    double r = fma(kf, PBU_EXP_KC(1, -0.6931471803691238), x);
    r = fma(kf, PBU_EXP_KC(2, -1.9082149292705877e-10), r);
    double p = fma(PBU_EXP_KC(3, 2.5100375832561234e-08), r, PBU_EXP_KC(4, 2.7620075879983367e-07));
    p = fma(p, r, PBU_EXP_KC(5, 2.7557268480310024e-06));
    p = fma(p, r, PBU_EXP_KC(6, 2.4801521322368692e-05));
    p = fma(p, r, PBU_EXP_KC(7, 0.00019841269863040545));
    p = fma(p, r, PBU_EXP_KC(8, 0.0013888888917196719));
    p = fma(p, r, PBU_EXP_KC(9, 0.008333333333330065));
    p = fma(p, r, PBU_EXP_KC(10, 0.041666666666624164));
    p = fma(p, r, PBU_EXP_KC(11, 0.16666666666666669));
    p = fma(p, r, PBU_EXP_KC(12, 0.5000000000000001));
    const double v = 1.0 + fma(r * r, p, r);                     // in [0.70, 1.42]
    double res = with_hi_word(v, hi_word(v) + (k << 20));        // * 2^k on the exponent field: k in [-1021, 1023] for |x| < 708
    // |x| >= 708 (one integer compare on the hi word; inf and NaN land here too): the rare path
    const int32_t hx = hi_word(x);
    if (((uint32_t)hx & 0x7FFFFFFFu) >= 0x40862000u) res = (x != x) ? x : ((hx < 0) ? 0.0 : (double)INFINITY);
    return res;
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
// Table-driven base-2 variants: 2^y in 7 and log2(u) in 6 FP64 instructions (against 17 and 22 above), paid for with
// one shared-memory load each.  The FP64 pipe is what bounds the ODE models, the load/store pipe is idle there.  Base 2
// makes both argument reductions exact and removes the ln2 hi/lo bookkeeping; u^n = 2^(n log2 u) needs nothing else.
//   pbu_exp2_tab : y = k + j/1024 + r, |r| <= 2^-11 (r = y - (1024k+j)/1024 EXACTLY, one FMA);
//                  2^y = 2^k T[j] (1 + r (B1 + r (B2 + r B3))),  T[j] = 2^(j/1024)
//   pbu_log2_tab : u = 2^e m, m in [1, 2), i = top 9 mantissa bits;  r = m c_i - 1 = u (2^-e c_i) - 1, |r| <= 2^-10 (one FMA on
//                  u itself: the table value is scaled by 2^-e with one integer instruction, u is never taken apart);
//                  log2 u = e + (-log2 c_i + r (A1 + r (A2 + r (A3 + r A4))))
// The coefficients are Taylor series whose last term is economised over the reduced range (tools/gen_math_tables.py): a
// cubic and a quartic where the plain series need one degree more.  Tables: pbu_math_tables.h, 1024 + 2*512 doubles (16 KB).
// On the device the kernels stage them in shared memory once per block (Model::block_init), on an 8 KB boundary of the
// shared window, and pass the pointer; on the host the static copy.
// Checked against libm in tools/test_math.cu: exp2 <= 2 ulp; log2 <= 1 ulp of max(|log2 u|, 1) -- the error that
// matters for 2^(n log2 u) is the absolute one.
//
// Conflict-free replicas (the evaluation kernel of the propagation, where every lane of a warp looks up a different entry:
// ncu showed the shared-memory pipe at 95 % of its wavefront rate, 60 % of the wavefronts bank conflicts).  An 8-byte load
// is served per half warp, a 16-byte load per quarter warp, 128 bytes at a time: with 16 interleaved copies of the exp2
// table (entry j of copy r at byte 128 j + 8 r, lane l reads copy l mod 16) and 8 copies of the log2 table (entry i of
// copy r at 128 i + 16 r, copy l mod 8) the lanes served together always sit in different banks, whatever their indices.
// 128 KB + 64 KB of shared memory: one block of 1024 threads per SM (the kernel needs fewer than 64 registers).
#define PBU_EXP2_REPLICAS 16
#define PBU_LOG2_REPLICAS 8
#define PBU_REP_EXP2_BYTES (8 * PBU_EXP2_ENTRIES * PBU_EXP2_REPLICAS)
#define PBU_REP_LOG2_BYTES (16 * PBU_LOG2_ENTRIES * PBU_LOG2_REPLICAS)
#define PBU_REP_TABLE_BYTES (PBU_REP_EXP2_BYTES + PBU_REP_LOG2_BYTES + 128)      // + the leading coefficients (PBU_MATH_EXTRA_DOUBLES)
#define PBU_MATH_TABLE_DOUBLES (PBU_EXP2_ENTRIES + 2 * PBU_LOG2_ENTRIES)
#define PBU_LOG2_TABLE_OFFSET PBU_EXP2_ENTRIES
static const double pbu_math_tables_host[PBU_MATH_TABLE_DOUBLES] = {PBU_EXP2_TABLE, PBU_LOG2_TABLE};
#ifdef __CUDACC__
static __device__ const double pbu_math_tables_dev[PBU_MATH_TABLE_DOUBLES] = {PBU_EXP2_TABLE, PBU_LOG2_TABLE};
#endif

// Polynomial coefficients.  On the device they are read as constant-bank operands of the DFMAs (a __constant__ array
// the compiler cannot fold): written as literals ptxas rebuilds every 64-bit constant with two UMOVs inside the loop,
// about 30 instruction-issue slots per RK4 step, and on this kernel every issued instruction is a cycle the FP64 pipe
// idles (an FP64 instruction holds the issue port for 2 cycles, everything else for 1; ncu: issue slots add up to the
// elapsed cycles).
#ifdef __CUDACC__
// exp2 coefficients in the scaled variable r' = PBU_EXP2_ENTRIES r (exact power-of-two scalings)
#define PBU_K_B1S (PBU_K_B1 / PBU_EXP2_ENTRIES)
#define PBU_K_B2S (PBU_K_B2 / ((double)PBU_EXP2_ENTRIES * PBU_EXP2_ENTRIES))
#define PBU_K_B3S (PBU_K_B3 / ((double)PBU_EXP2_ENTRIES * PBU_EXP2_ENTRIES * PBU_EXP2_ENTRIES))
static __constant__ double pbu_kc[8] = {0.0, PBU_K_B1S, PBU_K_B2S, PBU_K_B3S, PBU_K_A1, PBU_K_A2, PBU_K_A3, PBU_K_A4};
#endif
#ifdef __CUDA_ARCH__
#define PBU_KC(i, lit) pbu_kc[i]
#else
#define PBU_KC(i, lit) (lit)
#endif

#ifdef __CUDA_ARCH__
// 32-bit shared-window address of the tables, which the caller has placed on an 8 KB boundary of that window
// (ArrheniusModel::tables).  The mov hides the known-zero low bits from the front end, which would otherwise turn
// base | offset back into an add; ptxas folds the mov and merges mask and OR into one LOP3.
__device__ __forceinline__ uint32_t pbu_tab_base(const double *tab) {
    uint32_t b;
    asm("mov.u32 %0, %1;" : "=r"(b) : "r"((uint32_t)__cvta_generic_to_shared(tab)));
    return b;
}
#endif

// The first FMA of each polynomial has TWO constant operands and a DFMA reads only one from the constant bank: left to
// itself ptxas reloads the other into a register inside the caller's loop, every trip (it rematerialises constant-bank
// loads at will, through inline-asm moves as well).  The caller therefore keeps the two leading coefficients (and its own
// -1/6) behind the tables in SHARED memory, loads them once per model evaluation and passes them in: a shared-memory
// load is not something ptxas repeats.
#define PBU_MATH_EXTRA_DOUBLES 4
#ifdef __CUDACC__
static __device__ const double pbu_math_extra_dev[PBU_MATH_EXTRA_DOUBLES] = {PBU_K_A4, PBU_K_B3S, -1.0 / 6.0, 0.0};
#endif

// The exp2 table as the kernels stage it in shared memory: entry j with j 2^10 taken off its hi word.  The scale 2^k then goes
// onto the entry with ONE multiply-add on the integer m = 1024 k + j the reduction already has (m 2^10 + hi = k 2^20 + hi(T_j)),
// instead of a shift and a multiply-add; the entry itself is never used unscaled.
#ifdef __CUDACC__
__device__ __forceinline__ double pbu_exp2_staged(int j) {
    const double T = pbu_math_tables_dev[j];
    return __hiloint2double(__double2hiint(T) - (j << 10), __double2loint(T));
}
#endif

// 2^y for y < 1024, taking ys = PBU_EXP2_ENTRIES * y (the caller keeps its exponents in that unit: scaling by a power of two is exact,
// and the reduction becomes three plain additions with immediate operands -- no constant held in a register).  The result is
// forced to (almost) zero where y <= -1022 or where uh, the HI WORD of the caller's log argument u, is not that of a positive
// normal number (negative, zero, denormal: pbu_log2_tab returned garbage for it): both tests ride on ONE selection of the
// table value's hi word, which leaves a tiny denormal instead of the exact tail below 2^-1022 (-inf included).  The scale
// 2^k goes onto the TABLE value before the last FMA, so the result needs no fix-up.  The FP64 pipe sees the seven arithmetic
// instructions only.  NaN with the sign bit clear propagates.
// REP: `rep_addr` is the 32-bit shared-window address of THIS LANE's copy of entry 0 in the replicated table (see
// "Conflict-free replicas" below); `tab` is unused.
template <bool REP>
PBU_HD double pbu_exp2_tab_scaled_sel_t(double ys, const double *tab, uint32_t rep_addr, int32_t uh, double kB3) {
    const double SHIFT = 6755399441055744.0;                     // 1.5 * 2^52
    const double t = ys + SHIFT;
    const double kf = t - SHIFT;                                 // 1024 k + j as a double
    const int32_t m = (int32_t)(uint32_t)(
#ifdef __CUDA_ARCH__
        __double2loint(t)
#else
        [](double v) { uint64_t b; memcpy(&b, &v, 8); return (int32_t)(uint32_t)b; }(t)
#endif
    );
    const double r = ys - kf;                                    // exact; |r| <= 1/2, i.e. 2^-11 in units of y
    double p = fma(r, kB3, PBU_KC(2, PBU_K_B2S));                // kB3 = PBU_K_B3S, held in a register by the caller (see pbu_lead_coefficients)
    p = fma(p, r, PBU_KC(1, PBU_K_B1S));
    const double v = p * r;                                      // 2^(r/1024) - 1 (no constant term: 2^0 is exactly 1)
#ifdef __CUDA_ARCH__
    double T;
    if constexpr (REP) {                                         // entry j of the lane's replica: rep_addr + 128 j (mask, then one multiply-add)
        asm("{\n\t.reg .u32 j, a;\n\t"
            "and.b32 j, %1, %3;\n\t"
            "mad.lo.u32 a, j, %4, %2;\n\t"
            "ld.shared.f64 %0, [a];\n\t}"
            : "=d"(T) : "r"(m), "r"(rep_addr), "n"(PBU_EXP2_ENTRIES - 1), "n"(8 * PBU_EXP2_REPLICAS));
    } else {                                                     // tab is 8 KB aligned in the shared window: base | offset
        asm("ld.shared.f64 %0, [%1];" : "=d"(T) : "r"(pbu_tab_base(tab) | (((uint32_t)m << 3) & (8u * PBU_EXP2_ENTRIES - 8u))));      // is one LOP3
    }
    // hi word of T 2^k where kept, 0 otherwise: two compares into one predicate, one multiply-add, one select
    double Ts;
    asm("{\n\t.reg .pred p, q;\n\t.reg .s32 lo, hi;\n\t"
        "setp.ge.s32 q, %3, 0x00100000;\n\t"                      // u positive and normal
        "setp.lt.and.u32 p, %2, 0xC12FF000, q;\n\t"               // ys > -1022 * 1024 (negative doubles order by magnitude as unsigned integers)
        "mov.b64 {lo, hi}, %1;\n\t"
        "mad.lo.s32 hi, %4, 1024, hi;\n\t"                      // (1024 k + j) 2^10 + (hi(T_j) - j 2^10) = k 2^20 + hi(T_j): see pbu_exp2_staged
        "selp.b32 hi, hi, 0, p;\n\t"
        "mov.b64 %0, {lo, hi};\n\t}"
        : "=d"(Ts) : "d"(T), "r"(hi_word(ys)), "r"(uh), "r"(m));
#else
    const double T = tab[m & (PBU_EXP2_ENTRIES - 1)];
    const bool keep = (uh >= 0x00100000) && ((uint32_t)hi_word(ys) < 0xC12FF000u);
    const double Ts = with_hi_word(T, keep ? hi_word(T) + (m >> 10) * 1048576 : 0);
#endif
    return fma(Ts, v, Ts);
}
PBU_HD double pbu_exp2_tab_scaled_sel(double ys, const double *tab, int32_t uh, double kB3) {
    return pbu_exp2_tab_scaled_sel_t<false>(ys, tab, 0u, uh, kB3);
}
PBU_HD double pbu_exp2_tab(double y, const double *tab) { return pbu_exp2_tab_scaled_sel(y * (double)PBU_EXP2_ENTRIES, tab, 0x3FF00000, PBU_K_B3S); }

// log2(u) for normal 0 < u < 2^1000 (anything else: finite garbage, to be discarded by the caller -- pbu_exp2_tab_scaled_sel
// does it from the hi word of u)
template <bool REP>
PBU_HD double pbu_log2_tab_t(double u, const double *tab, uint32_t rep_addr, double kA4) {
    const int32_t hi = hi_word(u);
    const int32_t e = (hi >> 20) - 1023;
#ifdef __CUDA_ARCH__
    double cs, nlc;                                              // one 16-byte shared-memory load, address as in exp2
    if constexpr (REP) {                                         // rep_addr + 128 i, i = the top 9 mantissa bits: mask in place, then hi(x 2^28) = x >> 4 with the add
        asm("{\n\t.reg .u32 i, a;\n\t"
            "and.b32 i, %2, %4;\n\t"
            "mad.hi.u32 a, i, %5, %3;\n\t"
            "ld.shared.v2.f64 {%0, %1}, [a];\n\t}"
            : "=d"(cs), "=d"(nlc) : "r"(hi), "r"(rep_addr), "n"((PBU_LOG2_ENTRIES - 1) << 11), "n"((16 * PBU_LOG2_REPLICAS) << 21));
    } else {
        asm("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(cs), "=d"(nlc)
            : "r"(pbu_tab_base(tab) | (((uint32_t)hi >> 7) & (16u * PBU_LOG2_ENTRIES - 16u))), "n"(8 * PBU_LOG2_TABLE_OFFSET));
    }
    // c 2^-e: the exponent field of u comes off the exponent field of c, in place (one three-input integer add)
    asm("{\n\t.reg .s32 lo, hi, t;\n\t"
        "mov.b64 {lo, hi}, %0;\n\t"
        "and.b32 t, %1, 0x7FF00000;\n\t"
        "sub.s32 hi, hi, t;\n\t"
        "add.s32 hi, hi, 0x3FF00000;\n\t"
        "mov.b64 %0, {lo, hi};\n\t}"
        : "+d"(cs) : "r"(hi));
#else
    const int32_t i = (hi >> 11) & (PBU_LOG2_ENTRIES - 1);
    const double c = tab[PBU_LOG2_TABLE_OFFSET + 2 * i], nlc = tab[PBU_LOG2_TABLE_OFFSET + 2 * i + 1];
    const double cs = with_hi_word(c, hi_word(c) - (hi & 0x7FF00000) + 0x3FF00000);
#endif
    const double r = fma(u, cs, -1.0);
    double q = fma(r, kA4, PBU_KC(6, PBU_K_A3));                 // economised: see the header of this section
    q = fma(q, r, PBU_KC(5, PBU_K_A2));
    q = fma(q, r, PBU_KC(4, PBU_K_A1));
    return fma(r, q, nlc) + (double)e;                           // small part first: one rounding at full size
}
PBU_HD double pbu_log2_tab(double u, const double *tab, double kA4 = PBU_K_A4) { return pbu_log2_tab_t<false>(u, tab, 0u, kA4); }

}  // namespace pbu
