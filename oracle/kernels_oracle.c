/*
 * oracle/kernels_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C CPU restatement of the reference's kernel-matrix builders
 * (/root/reference/gp/ext/gaussian_c.pyx and periodic_c.pyx).  Each function
 * cites the reference lines it follows.  The element formulas are evaluated
 * with libm exp/sin/cos in the same operation order as the reference so that
 * the result agrees with the reference's compiled Cython to the last bit or
 * two; tests/test_oracle.py pins it against oracle/_ref (the reference's own
 * .pyx compiled in place) and against the golden vectors in tests/golden/.
 *
 * Parity status: PINNED (see tests/test_oracle.py and tests/golden/).
 *
 * Build: gcc -O2 -shared -fPIC oracle/kernels_oracle.c -o oracle/libgporacle.so -lm
 */
#include <math.h>
#include <stdint.h>

/* gaussian_c.pyx:14-15 / periodic_c.pyx:14-15 */
static double sqrt_2_div_pi(void) { return sqrt(2.0 / M_PI); }
/* MIN = log(exp2(minexp + 4)), minexp = -1022  (gaussian_c.pyx:15) */
static double min_log(void) { return log(exp2(-1022.0 + 4.0)); }

double gpo_min_log(void) { return min_log(); }

/* ---------------- Gaussian (gaussian_c.pyx) ---------------- */

/* slice ids: 0 K (:18-36) 1 dK_dh (:51-69) 2 dK_dw (:72-92) 3 d2K_dhdh (:95-113)
 *            4 d2K_dhdw (:116-136) 5 d2K_dwdh (:139-140) 6 d2K_dwdw (:143-164) */
void gpo_gaussian_slice(int slice, double *out, const double *x1, int64_t n1,
                        const double *x2, int64_t n2, double h, double w)
{
    const double S = sqrt_2_div_pi(), MIN = min_log();
    const double h2 = h * h, w2 = w * w;
    const double c1 = -0.5 / w2;
    double c2 = 0, c3 = 0, c4 = 0;
    switch (slice) {
    case 0: c2 = 0.5 * S * h2 / w; break;                       /* :27 */
    case 1: c2 = S * h / w; break;                              /* :61 */
    case 2: c2 = 0.5 * S * h2 / w2; c3 = 0.5 * S * h2 / pow(w, 4); break;   /* :83-84 */
    case 3: c2 = S / w; break;                                  /* :105 */
    case 4: case 5: c2 = S * h / w2; c3 = S * h / pow(w, 4); break;         /* :127-128 */
    case 6: c2 = S * h2 / pow(w, 3); c3 = 2.5 * S * h2 / pow(w, 5);
            c4 = 0.5 * S * h2 / pow(w, 7); break;               /* :154-156 */
    }
    for (int64_t i = 0; i < n1; i++)
        for (int64_t j = 0; j < n2; j++) {
            const double d = x1[i] - x2[j];
            const double d2 = d * d;
            const double e = c1 * d2;
            double v;
            if (e < MIN) v = 0.0;                               /* :33-34 and siblings */
            else switch (slice) {
                case 0: case 1: case 3: v = c2 * exp(e); break;
                case 2: case 4: case 5: v = exp(e) * (c3 * d2 - c2); break;
                default: v = exp(e) * (c4 * (d2 * d2) - c3 * d2 + c2); break;
            }
            out[i * n2 + j] = v;
        }
}

/* ---------------- Periodic (periodic_c.pyx) ---------------- */

/* slice ids: 0 K (:18-30) 1 dK_dh (:53-65) 2 dK_dw (:68-80) 3 dK_dp (:83-96)
 *            4..12 = hessian row-major over (h,w,p): hh :99-111, hw :114-126,
 *            hp :129-142, wh :145-157, ww :160-172, wp :175-188, ph :191-204,
 *            pw :207-220, pp :223-235 */
void gpo_periodic_slice(int slice, double *out, const double *x1, int64_t n1,
                        const double *x2, int64_t n2, double h, double w, double p)
{
    const double h2 = h * h, w2 = w * w, p2 = p * p;
    for (int64_t i = 0; i < n1; i++)
        for (int64_t j = 0; j < n2; j++) {
            const double d = x1[i] - x2[j];
            const double u = 0.5 * d / p;
            const double s = sin(u), c = cos(u);
            const double E = exp(-2.0 * (s * s) / w2);
            double v;
            switch (slice) {
            case 0: v = h2 * E; break;
            case 1: v = 2.0 * h * E; break;
            case 2: v = 4.0 * h2 * E * (s * s) / pow(w, 3); break;
            case 3: v = 2.0 * d * h2 * E * s * c / (p2 * w2); break;
            case 4: v = 2.0 * E; break;
            case 5: case 7: v = 8.0 * h * E * (s * s) / pow(w, 3); break;
            case 6: case 10: v = 4.0 * d * h * E * s * c / (p2 * w2); break;
            case 8: v = -12.0 * h2 * E * (s * s) / pow(w, 4)
                        + 16.0 * h2 * E * pow(s, 4) / pow(w, 6); break;
            case 9: case 11: v = -4.0 * d * h2 * E * s * c / (p2 * pow(w, 3))
                        + 8.0 * d * h2 * E * pow(s, 3) * c / (p2 * pow(w, 5)); break;
            default: v = (d * d) * h2 * E * (s * s) / (pow(p, 4) * w2)
                        - 1.0 * (d * d) * h2 * E * (c * c) / (pow(p, 4) * w2)
                        + 4.0 * (d * d) * h2 * E * (s * s) * (c * c) / (pow(p, 4) * pow(w, 4))
                        - 4.0 * d * h2 * E * s * c / (pow(p, 3) * w2); break;
            }
            out[i * n2 + j] = v;
        }
}
