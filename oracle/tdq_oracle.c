/*
 * tdq_oracle.c -- plain C restatement of the elementwise arithmetic of the explicit Runge-Kutta hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/ode_oracle.py): nothing under torchdiffeq_b200/ links or loads this
 * file.  It exists so that the order of roundings the CUDA kernels must reproduce is also written down in
 * the simplest possible form -- scalar loops, one rounding per operation, no contraction
 * (compile with -ffp-contract=off) -- and is itself checked bitwise against the torch-CPU oracle, which is
 * pinned to the reference's golden vectors (tests/test_oracle_c.py).
 *
 * Each function cites the reference lines (torchdiffeq/_impl/) it follows.
 *     gcc -O2 -ffp-contract=off -fPIC -shared -o oracle/_build/libtdq_oracle.so oracle/tdq_oracle.c -lm
 */
#include <math.h>
#include <stddef.h>

#define DEF_COMBINE(NAME, T)                                                                              \
    /* rk_common.py:79 / :85 / :89: out = y0 + sum_j k_j*c_j (y0 == NULL: the bare sum, the error estimate) */ \
    void NAME(T *out, const T *y0, const T *const *k, const T *c, int nk, size_t n) {                     \
        for (size_t i = 0; i < n; ++i) {                                                                  \
            T acc = k[0][i] * c[0];                                                                       \
            for (int j = 1; j < nk; ++j) acc = acc + k[j][i] * c[j];                                      \
            out[i] = y0 ? y0[i] + acc : acc;                                                              \
        }                                                                                                 \
    }
DEF_COMBINE(orc_combine_f32, float)
DEF_COMBINE(orc_combine_f64, double)

#define DEF_SUMSQ(NAME, T, ABS)                                                                           \
    /* misc.py:80-82 + :22-23: sum((err/tol)^2), tol = atol + rtol*max(|y0|,|y1|) in T; float64 sum */   \
    double NAME(const T *err, const T *y0, const T *y1, double rtol, double atol, size_t n) {             \
        const T rt = (T)rtol, at = (T)atol;                                                               \
        double s = 0.0;                                                                                   \
        for (size_t i = 0; i < n; ++i) {                                                                  \
            const T a = ABS(y0[i]), b = ABS(y1[i]);                                                       \
            const T m = a > b ? a : b;                                                                    \
            const T tol = at + rt * m;                                                                    \
            const T q = err[i] / tol;                                                                     \
            s += (double)(q * q);                                                                         \
        }                                                                                                 \
        return s;                                                                                         \
    }
DEF_SUMSQ(orc_error_sumsq_f32, float, fabsf)
DEF_SUMSQ(orc_error_sumsq_f64, double, fabs)

#define DEF_FIT(NAME, T)                                                                                  \
    /* interp.py:17-22 with y_mid from rk_common.py:366; coefficients e,d,c,b,a */                        \
    void NAME(T *e, T *d, T *c, T *b, T *a, const T *y0, const T *y1, const T *ymid, const T *f0,         \
              const T *f1, T dt, size_t n) {                                                              \
        const T two_dt = (T)2 * dt;                                                                       \
        for (size_t i = 0; i < n; ++i) {                                                                  \
            a[i] = two_dt * (f1[i] - f0[i]) - (T)8 * (y1[i] + y0[i]) + (T)16 * ymid[i];                   \
            b[i] = dt * ((T)5 * f0[i] - (T)3 * f1[i]) + (T)18 * y0[i] + (T)14 * y1[i] - (T)32 * ymid[i];  \
            c[i] = dt * (f1[i] - (T)4 * f0[i]) - (T)11 * y0[i] - (T)5 * y1[i] + (T)16 * ymid[i];          \
            d[i] = dt * f0[i];                                                                            \
            e[i] = y0[i];                                                                                 \
        }                                                                                                 \
    }
DEF_FIT(orc_interp_fit_f32, float)
DEF_FIT(orc_interp_fit_f64, double)

#define DEF_EVAL(NAME, T)                                                                                 \
    /* interp.py:39-46: x in float64 then cast; running powers, not Horner */                            \
    void NAME(T *out, const T *e, const T *d, const T *c, const T *b, const T *a, double t0, double t1,   \
              double t, size_t n) {                                                                       \
        const T x = (T)((t - t0) / (t1 - t0));                                                            \
        for (size_t i = 0; i < n; ++i) {                                                                  \
            T total = e[i] + x * d[i];                                                                    \
            T xp = x * x;                                                                                 \
            total = total + xp * c[i];                                                                    \
            xp = xp * x;                                                                                  \
            total = total + xp * b[i];                                                                    \
            xp = xp * x;                                                                                  \
            total = total + xp * a[i];                                                                    \
            out[i] = total;                                                                               \
        }                                                                                                 \
    }
DEF_EVAL(orc_interp_eval_f32, float)
DEF_EVAL(orc_interp_eval_f64, double)

#define DEF_RK4(NAME, T)                                                                                  \
    /* rk_common.py:110-118 + solvers.py:115; which = 1..4 as in include/tdq.h */                        \
    void NAME(int which, T *out, const T *y0, const T *k1, const T *k2, const T *k3, const T *k4, T dt,   \
              size_t n) {                                                                                 \
        const T third = (T)(1.0 / 3.0);                                                                   \
        for (size_t i = 0; i < n; ++i) {                                                                  \
            if (which == 1) out[i] = y0[i] + dt * k1[i] * third;                                          \
            else if (which == 2) out[i] = y0[i] + dt * (k2[i] - k1[i] * third);                           \
            else if (which == 3) out[i] = y0[i] + dt * (k1[i] - k2[i] + k3[i]);                           \
            else out[i] = y0[i] + (k1[i] + (T)3 * (k2[i] + k3[i]) + k4[i]) * dt * (T)0.125;               \
        }                                                                                                 \
    }
DEF_RK4(orc_rk4_stage_f32, float)
DEF_RK4(orc_rk4_stage_f64, double)

/* misc.py:85-95 followed by the clamp of rk_common.py:359 (all float64) */
double orc_optimal_step(double last, double ratio, double safety, double ifactor, double dfactor, int order,
                        double min_step, double max_step) {
    double next;
    if (ratio == 0.0) {
        next = last * ifactor;
    } else {
        const double df = ratio < 1.0 ? 1.0 : dfactor;
        const double cand = safety / pow(ratio, 1.0 / (double)order);
        const double inner = (cand != cand) ? cand : (cand > df ? cand : df);
        const double factor = (inner != inner) ? inner : (ifactor < inner ? ifactor : inner);
        next = last * factor;
    }
    if (next == next) next = next < min_step ? min_step : (next > max_step ? max_step : next);
    return next;
}
