// Host check of pybitup_b200/csrc/pbu_math.cuh against libm (run: nvcc -O2 -o /tmp/test_math tools/test_math.cu && /tmp/test_math)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <random>
#include "../pybitup_b200/csrc/pbu_math.cuh"

static double ulp_err(double got, double ref) {
    if (ref == 0.0) return got == 0.0 ? 0.0 : 1e300;
    int e;
    frexp(ref, &e);
    return fabs(got - ref) / ldexp(1.0, e - 53);
}

int main() {
    std::mt19937_64 rng(1);
    std::uniform_real_distribution<double> ue(-700.0, 700.0), us(-60.0, 20.0), u01(0.0, 1.0);
    double mef = 0, me = 0, ml = 0, mp = 0, met = 0, mlt = 0, mpt = 0, mst = 0;
    long bad = 0;
    const double *tab = pbu::pbu_math_tables_host;
    for (long i = 0; i < 20000000; ++i) {
        double x = (i & 1) ? ue(rng) : us(rng);
        me = fmax(me, ulp_err(pbu::pbu_exp(x), exp(x)));
        mef = fmax(mef, ulp_err(pbu::pbu_exp_full(x), exp(x)));
        double u = (i % 3 == 0) ? u01(rng) : ((i % 3 == 1) ? exp(-40.0 * u01(rng)) : 1.0 - 1e-3 * u01(rng) * u01(rng));
        if (u <= 0.0) continue;
        const double l = pbu::pbu_log(u), lr = log(u);
        const double el = ulp_err(l, lr);
        ml = fmax(ml, el);
        const double n = 0.5 + 2.5 * u01(rng);
        const double pw = pbu::pbu_exp(n * l), pr = pow(u, n);
        const double rel = pr > 0 ? fabs(pw - pr) / pr : 0.0;
        mp = fmax(mp, rel);
        if (el > 2.0) ++bad;
        // table-driven variants: exp in ulp; log in ulp of max(|log u|, 1) (absolute accuracy is what exp(n log u) needs)
        met = fmax(met, ulp_err(pbu::pbu_exp2_tab(x, tab), exp2(x)));
        const double lt = pbu::pbu_log2_tab(u, tab), lr2 = log2(u);
        mlt = fmax(mlt, fabs(lt - lr2) / ldexp(1.0, ilogb(fmax(fabs(lr2), 1.0)) - 52));
        const double pwt = pbu::pbu_exp2_tab(n * lt, tab);
        mpt = fmax(mpt, pr > 0 ? fabs(pwt - pr) / pr : 0.0);
        // one stage of the Arrhenius model: 2^(y + n log2 u) against a long double reference
        const double ys = us(rng);
        const long double ref = exp2l((long double)ys + (long double)n * log2l((long double)u));
        const double st = pbu::pbu_exp2_tab(fma(n, lt, ys), tab);
        mst = fmax(mst, ref > 0 ? (double)(fabsl((long double)st - ref) / ref) : 0.0);
    }
    printf("max ulp error: exp %.3f  log %.3f   max relative error of exp(n log u) vs pow: %.3e   log errors > 2 ulp: %ld\n", me, ml, mp, bad);
    printf("edge: exp(0)=%.17g log(1)=%.17g exp(-701)=%g log(0.5)=%.17g (ref %.17g)\n", pbu::pbu_exp(0.0), pbu::pbu_log(1.0), pbu::pbu_exp(-701.0),
           pbu::pbu_log(0.5), log(0.5));
    printf("table-driven, base 2: max ulp error exp2 %.3f  log2 (ulp of max(|log2|, 1)) %.3f   2^(n log2 u) vs pow: %.3e\n", met, mlt, mpt);
    printf("stage 2^(y + n log2 u), y in [-60, 20]: max relative error %.3e\n", mst);
    printf("edge: exp2_tab(0)=%.17g log2_tab(1)=%.17g log2_tab(1-2^-53)=%.17g (ref %.17g) log2_tab(0.5)=%.17g exp2_tab(-1022.5)=%g\n",
           pbu::pbu_exp2_tab(0.0, tab), pbu::pbu_log2_tab(1.0, tab), pbu::pbu_log2_tab(1.0 - ldexp(1.0, -53), tab),
           log2(1.0 - ldexp(1.0, -53)), pbu::pbu_log2_tab(0.5, tab), pbu::pbu_exp2_tab(-1022.5, tab));
    printf("pbu_exp_full: max ulp error %.3f; exp(-inf)=%g exp(inf)=%g exp(nan)=%g exp(708.5)=%g exp(-708.5)=%g exp(707.9)=%.17g (ref %.17g) exp(-707.9)=%.17g (ref %.17g) exp(0)=%.17g\n", mef,
           pbu::pbu_exp_full(-INFINITY), pbu::pbu_exp_full(INFINITY), pbu::pbu_exp_full(NAN), pbu::pbu_exp_full(708.5), pbu::pbu_exp_full(-708.5),
           pbu::pbu_exp_full(707.9), exp(707.9), pbu::pbu_exp_full(-707.9), exp(-707.9), pbu::pbu_exp_full(0.0));
    if (!(mef < 2.0 && pbu::pbu_exp_full(-INFINITY) == 0.0 && std::isinf(pbu::pbu_exp_full(INFINITY)) && std::isnan(pbu::pbu_exp_full(NAN)) && pbu::pbu_exp_full(0.0) == 1.0)) return 1;
    return (me < 2.0 && ml < 2.5 && mp < 2e-14 && met < 2.5 && mlt < 2.5 && mpt < 2e-14 && mst < 3e-14) ? 0 : 1;
}
