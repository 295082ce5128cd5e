/*
 * dddm_oracle.c — CPU restatement (plain C, double precision) of the DDDM hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker or as the timed CPU baseline.  The shipped path is the CUDA
 * library in ddm_b200/csrc and it fails loudly when that library is missing.
 *
 * Parity pinning: the reference (edluyuan/ddm) has no tests or golden vectors of its own
 * (SURVEY.md §4).  This oracle is pinned against outputs of the reference's own Python
 * functions run in the build container: tests/golden/ (npz files), produced by
 * tests/golden/make_golden.py (imports /root/reference/dddm with a stub matplotlib), and
 * checked by tests/test_oracle_golden.py.
 *
 * Each function cites the reference lines it restates (paths relative to /root/reference).
 * All inputs are double arrays (callers widen fp32/bf16 exactly), all accumulation is
 * double, pairs are visited in a fixed order, so results are deterministic.
 */
#include <math.h>
#include <stddef.h>

#define ORACLE_EPS_POW 1e-12 /* dddm/losses.py:14,24 — added inside the power, beta != 2 only */
#define ORACLE_EPS_W 1e-12   /* dddm/losses.py:33-34 */
#define ORACLE_EPS_BR 1e-8   /* dddm/schedules.py:47 */

static double sqdist(const double *a, const double *b, long D) {
    double s = 0.0;
    for (long k = 0; k < D; ++k) {
        double d = a[k] - b[k];
        s += d * d;
    }
    return s;
}

/* value of the beta-power term: dddm/losses.py:11-14 and :21-24.
 * beta == 2.0 (exact compare, as the reference does) drops both the epsilon and the pow. */
static double pow_term(double d2, double beta) {
    if (beta == 2.0) return d2;
    return pow(d2 + ORACLE_EPS_POW, 0.5 * beta);
}

/* d/d(d2) of pow_term */
static double pow_term_deriv(double d2, double beta) {
    if (beta == 2.0) return 1.0;
    return 0.5 * beta * pow(d2 + ORACLE_EPS_POW, 0.5 * beta - 1.0);
}

/*
 * generalized_energy_terms — dddm/losses.py:5-25.
 *   xhat [B,m,D], x0 [B,D]  ->  conf = mean_{b,i} f(|x0_b - xhat_bi|^2)
 *                               inter = mean_{b,i,j!=i} f(|xhat_bi - xhat_bj|^2)
 * `lam` is not a parameter here because the reference never reads it (losses.py:6).
 * Optionally returns per-row sums (row_conf[b] = sum_i, row_inter[b] = sum_{i!=j}).
 */
void oracle_energy_terms(const double *xhat, const double *x0, long B, long m, long D, double beta,
                         double *conf, double *inter, double *row_conf, double *row_inter) {
    double csum = 0.0, isum = 0.0;
    for (long b = 0; b < B; ++b) {
        const double *xb = xhat + b * m * D;
        const double *cb = x0 + b * D;
        double rc = 0.0, ri = 0.0;
        for (long i = 0; i < m; ++i) rc += pow_term(sqdist(cb, xb + i * D, D), beta);
        for (long i = 0; i < m; ++i)
            for (long j = 0; j < m; ++j)
                if (i != j) ri += pow_term(sqdist(xb + i * D, xb + j * D, D), beta);
        if (row_conf) row_conf[b] = rc;
        if (row_inter) row_inter[b] = ri;
        csum += rc;
        isum += ri;
    }
    *conf = csum / (double)(B * m);
    *inter = isum / (double)(B * m * (m - 1));
}

/*
 * Closed-form gradient of  g_conf*conf + g_inter*inter  w.r.t. xhat (and x0 when grad_x0 != NULL).
 * Derived from dddm/losses.py:10-24 (what autograd replays at train_cifar10_dit.py:166);
 * SURVEY.md §8(a) states the same formula.  Ordered pairs (i,j),(j,i) both appear in the
 * off-diagonal mean, hence the factor 2 on the interaction part.
 */
void oracle_energy_terms_grad(const double *xhat, const double *x0, long B, long m, long D, double beta,
                              double g_conf, double g_inter, double *grad_xhat, double *grad_x0) {
    const double sc = g_conf / (double)(B * m);
    const double si = g_inter / (double)(B * m * (m - 1));
    for (long b = 0; b < B; ++b) {
        const double *xb = xhat + b * m * D;
        const double *cb = x0 + b * D;
        double *gb = grad_xhat + b * m * D;
        for (long k = 0; k < m * D; ++k) gb[k] = 0.0;
        if (grad_x0)
            for (long k = 0; k < D; ++k) grad_x0[b * D + k] = 0.0;
        for (long i = 0; i < m; ++i) {
            /* d2 = |x0 - xhat_i|^2 ; d(d2)/dxhat_i = 2 (xhat_i - x0) */
            double a = 2.0 * sc * pow_term_deriv(sqdist(cb, xb + i * D, D), beta);
            for (long k = 0; k < D; ++k) {
                double d = xb[i * D + k] - cb[k];
                gb[i * D + k] += a * d;
                if (grad_x0) grad_x0[b * D + k] -= a * d;
            }
        }
        for (long i = 0; i < m; ++i)
            for (long j = i + 1; j < m; ++j) {
                double c = 2.0 * (2.0 * si) * pow_term_deriv(sqdist(xb + i * D, xb + j * D, D), beta);
                for (long k = 0; k < D; ++k) {
                    double d = xb[i * D + k] - xb[j * D + k];
                    gb[i * D + k] += c * d;
                    gb[j * D + k] -= c * d;
                }
            }
    }
}

/* sigmoid_weight — dddm/losses.py:28-35 with alpha_sigma of dddm/schedules.py:5-14. */
void oracle_sigmoid_weight(const double *t, long B, double bias, double *w) {
    for (long b = 0; b < B; ++b) {
        double a = 1.0 - t[b], s = t[b];
        double ratio = (a * a) / (s * s + ORACLE_EPS_W);
        double z = log(ratio + ORACLE_EPS_W) - bias;
        w[b] = 1.0 / (1.0 + exp(-z));
    }
}

/*
 * The loss of distributional_training_step — dddm/training.py:84-85:
 *   loss = mean_b w(t_b) * (conf - lam/(2(m-1)) * inter)
 * and its gradient w.r.t. xhat.  `weight` is the batch mean (possibly a global, all-reduced
 * mean: SURVEY.md §8(e)).  out = {loss, conf, inter}.
 */
void oracle_energy_loss(const double *xhat, const double *x0, long B, long m, long D, double beta,
                        double lam, double weight, double *out, double *grad_xhat) {
    double conf, inter;
    oracle_energy_terms(xhat, x0, B, m, D, beta, &conf, &inter, NULL, NULL);
    const double c = lam / (2.0 * (double)(m - 1));
    out[0] = weight * (conf - c * inter);
    out[1] = conf;
    out[2] = inter;
    if (grad_xhat) oracle_energy_terms_grad(xhat, x0, B, m, D, beta, weight, -weight * c, grad_xhat, NULL);
}

/*
 * forward_marginal_sample — dddm/schedules.py:17-25: x_t = (1-t) x0 + t eps, t per leading row,
 * followed by the m-fold row expansion of dddm/training.py:70 when m > 0 (xt_rep [B*m, D]).
 */
void oracle_forward_marginal(const double *x0, const double *t, const double *eps, long B, long D,
                             long m, double *xt, double *xt_rep) {
    for (long b = 0; b < B; ++b)
        for (long k = 0; k < D; ++k) {
            double v = (1.0 - t[b]) * x0[b * D + k] + t[b] * eps[b * D + k];
            if (xt) xt[b * D + k] = v;
            if (xt_rep)
                for (long i = 0; i < m; ++i) xt_rep[(b * m + i) * D + k] = v;
        }
}

/*
 * gaussian_bridge_mu_sigma coefficients — dddm/schedules.py:45-77.
 * coef = {c_xt, c_x0, std}:  mu = c_xt * xt + c_x0 * x0hat.
 */
void oracle_bridge_coeffs(double s, double t, double eps_churn, double *coef) {
    double a_s = 1.0 - s, a_t = 1.0 - t;
    double rho = s / (t + ORACLE_EPS_BR);
    double ar = a_t / (a_s + ORACLE_EPS_BR);
    double e2 = eps_churn * eps_churn;
    double r11 = ar * rho, r12 = ar * rho * rho;
    coef[0] = e2 * r12 + (1.0 - e2) * rho;
    coef[1] = a_s * (1.0 - e2 * r12 - (1.0 - e2) * r11);
    double inner = e2 * r11 + (1.0 - e2);
    double one_minus = 1.0 - inner * inner;
    if (one_minus < 0.0) one_minus = 0.0;
    double var = s * s * one_minus;
    if (var < 0.0) var = 0.0;
    coef[2] = sqrt(var);
}

/*
 * One Algorithm-2 update — dddm/sampling.py:29-31: x <- mu(s,t,x0hat,x) + std * z.
 * s, t have one entry (st_is_vector == 0) or N entries (per-sample times).
 */
void oracle_bridge_step(const double *x, const double *x0hat, const double *z, const double *s,
                        const double *t, int st_is_vector, double eps_churn, long N, long D, double *x_out,
                        double *mu_out) {
    double coef[3];
    if (!st_is_vector) oracle_bridge_coeffs(s[0], t[0], eps_churn, coef);
    for (long n = 0; n < N; ++n) {
        if (st_is_vector) oracle_bridge_coeffs(s[n], t[n], eps_churn, coef);
        for (long k = 0; k < D; ++k) {
            double mu = coef[0] * x[n * D + k] + coef[1] * x0hat[n * D + k];
            if (mu_out) mu_out[n * D + k] = mu;
            if (x_out) x_out[n * D + k] = mu + coef[2] * (z ? z[n * D + k] : 0.0);
        }
    }
}
