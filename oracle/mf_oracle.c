/*
 * CPU oracle for the funk-SVD / RSVD SGD loop.  TEST INFRASTRUCTURE ONLY
 * (see oracle/__init__.py): linked by tests/, smoke() and bench.py's CPU
 * baseline leg, never by the product.
 *
 * Plain-C restatement of the strictly sequential per-rating update the
 * reference runs in Cython:
 *   SVD.fit   util/matrix_factorization.pyx:104-155  (loop body :135-151)
 *   RSVD.fit  util/matrix_factorization.pyx:22-66    (loop body :45-61)
 * float64 throughout, one rating at a time, update t+1 sees update t.
 * The caller supplies the initial tables (the reference draws them from the
 * global numpy RNG, :38-39 / :124-125 -- the RNG is not re-implemented here).
 *
 * Pinned bit-for-bit against the reference's own compiled extension
 * (oracle/_ref, built by oracle/build_ref.py) in tests/test_oracle_golden.py.
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (no FMA contraction, so the
 * rounding sequence is the one the Cython-generated C performs).
 */
#include <stdint.h>
#include <stddef.h>

/* SVD.fit -- util/matrix_factorization.pyx:132-151.
 * pu [U,D], qi [I,D], bu [U], bi [I] are updated in place.
 * global_mean must already be 0 when !biased (:127-130).
 * Returns the sum of squared errors seen during the LAST epoch (not a
 * reference quantity; a convenience for loss-trajectory checks). */
double mf_oracle_svd_fit(int64_t n, const int32_t *users, const int32_t *items, const double *ratings,
                         int D, int n_epochs, int biased,
                         double lr_bu, double lr_bi, double lr_pu, double lr_qi,
                         double reg_bu, double reg_bi, double reg_pu, double reg_qi,
                         double global_mean,
                         double *pu, double *qi, double *bu, double *bi)
{
    double sse = 0.0;
    for (int epoch = 0; epoch < n_epochs; ++epoch) {
        sse = 0.0;
        for (int64_t t = 0; t < n; ++t) {
            double *p = pu + (size_t)users[t] * D;
            double *q = qi + (size_t)items[t] * D;
            double r = ratings[t];
            double dot = 0.0;                                   /* :137-139 */
            for (int f = 0; f < D; ++f) dot += q[f] * p[f];
            double err = r - (global_mean + bu[users[t]] + bi[items[t]] + dot);   /* :140 */
            sse += err * err;
            if (biased) {                                       /* :143-145 */
                bu[users[t]] += lr_bu * (err - reg_bu * bu[users[t]]);
                bi[items[t]] += lr_bi * (err - reg_bi * bi[items[t]]);
            }
            for (int f = 0; f < D; ++f) {                       /* :147-151 */
                double puf = p[f], qif = q[f];
                p[f] += lr_pu * (err * qif - reg_pu * puf);
                q[f] += lr_qi * (err * puf - reg_qi * qif);
            }
        }
    }
    return sse;
}

/* RSVD.fit -- util/matrix_factorization.pyx:41-61.
 * ui [U,D], vj [I,D], ci [U], dj [I] updated in place.  version 2 couples the
 * two biases through reg2*(c_i + d_j - mu) using their PRE-update values
 * (:51-55); the prediction has no global mean term (:49). */
double mf_oracle_rsvd_fit(int64_t n, const int32_t *users, const int32_t *items, const double *ratings,
                          int D, int n_epochs, int version,
                          double lr, double reg, double reg2, double global_mean,
                          double *ui, double *vj, double *ci, double *dj)
{
    double sse = 0.0;
    for (int epoch = 0; epoch < n_epochs; ++epoch) {
        sse = 0.0;
        for (int64_t t = 0; t < n; ++t) {
            int32_t i = users[t], j = items[t];
            double *p = ui + (size_t)i * D;
            double *q = vj + (size_t)j * D;
            double dot = 0.0;                                   /* :46-48 */
            for (int k = 0; k < D; ++k) dot += p[k] * q[k];
            double err = ratings[t] - (ci[i] + dj[j] + dot);    /* :49 */
            sse += err * err;
            if (version == 2) {                                 /* :51-55 */
                double cii = ci[i], djj = dj[j];
                ci[i] += lr * (err - reg2 * (cii + djj - global_mean));
                dj[j] += lr * (err - reg2 * (cii + djj - global_mean));
            }
            for (int k = 0; k < D; ++k) {                       /* :57-61 */
                double uik = p[k], vjk = q[k];
                p[k] += lr * (err * vjk - reg * uik);
                q[k] += lr * (err * uik - reg * vjk);
            }
        }
    }
    return sse;
}

/* predict -- util/matrix_factorization.pyx:157-167 (SVD) / :68-78 (RSVD).
 * Returns 0 and writes *est, or -1 / -2 for an invalid user / item code
 * (the reference raises ValueError('Invalid user code' / 'Invalid item code')).
 * with_bias: SVD.biased, or RSVD.version == 2;  mu: SVD.global_mean (0 for RSVD). */
int mf_oracle_predict(int64_t u, int64_t i, int64_t U, int64_t I, int D, int with_bias, double mu,
                      const double *pu, const double *qi, const double *bu, const double *bi, double *est)
{
    if (u >= U) return -1;
    if (i >= I) return -2;
    double dot = 0.0;
    for (int f = 0; f < D; ++f) dot += qi[(size_t)i * D + f] * pu[(size_t)u * D + f];
    *est = with_bias ? mu + bu[u] + bi[i] + dot : dot;
    return 0;
}

/* SVDpp.fit -- util/matrix_factorization.pyx:193-271 (loop body :238-263).  Strictly sequential: every rating of user u
 * updates pu[u], qi[i], both biases AND the implicit-feedback row yj[j] of EVERY item j the user rated (:259-263), so
 * two ratings of different users conflict whenever the users share an item.
 * ur_ptr [U+1], ur_idx: the items each user rated, in the order of their first appearance in the training frame
 * (`ur[u].append((i, r))`, :231-234 -- list order is what the sums below follow).  Operation order as in the .pyx:
 * u_impl_fdb[f] accumulates yj[j, f] / sqrt_Iu over j in list order (:245-247); the factor loop reads puf, qif BEFORE
 * updating (:254-258); yj[j, f] uses the OLD qif (:261-263).  Returns the last epoch's sum of squared errors. */
double mf_oracle_svdpp_fit(int64_t n, const int32_t *users, const int32_t *items, const double *ratings,
                           int D, int n_epochs,
                           double lr_bu, double lr_bi, double lr_pu, double lr_qi, double lr_yj,
                           double reg_bu, double reg_bi, double reg_pu, double reg_qi, double reg_yj,
                           double global_mean, const int64_t *ur_ptr, const int32_t *ur_idx,
                           double *pu, double *qi, double *yj, double *bu, double *bi, double *u_impl_fdb)
{
    double sse = 0.0;
    for (int epoch = 0; epoch < n_epochs; ++epoch) {
        sse = 0.0;
        for (int64_t t = 0; t < n; ++t) {
            const int32_t u = users[t], i = items[t];
            const int32_t *Iu = ur_idx + ur_ptr[u];
            const int64_t nI = ur_ptr[u + 1] - ur_ptr[u];
            const double sqrt_Iu = __builtin_sqrt((double)nI);                    /* :241 */
            for (int f = 0; f < D; ++f) u_impl_fdb[f] = 0.0;                      /* :243 */
            for (int64_t k = 0; k < nI; ++k)                                      /* :244-246 */
                for (int f = 0; f < D; ++f) u_impl_fdb[f] += yj[(size_t)Iu[k] * D + f] / sqrt_Iu;
            double dot = 0.0;                                                     /* :248-250 */
            for (int f = 0; f < D; ++f) dot += qi[(size_t)i * D + f] * (pu[(size_t)u * D + f] + u_impl_fdb[f]);
            const double err = ratings[t] - (global_mean + bu[u] + bi[i] + dot);  /* :252 */
            sse += err * err;
            bu[u] += lr_bu * (err - reg_bu * bu[u]);                              /* :255-256 */
            bi[i] += lr_bi * (err - reg_bi * bi[i]);
            for (int f = 0; f < D; ++f) {                                         /* :259-265 */
                const double puf = pu[(size_t)u * D + f], qif = qi[(size_t)i * D + f];
                pu[(size_t)u * D + f] += lr_pu * (err * qif - reg_pu * puf);
                qi[(size_t)i * D + f] += lr_qi * (err * (puf + u_impl_fdb[f]) - reg_qi * qif);
                for (int64_t k = 0; k < nI; ++k) {
                    double *y = yj + (size_t)Iu[k] * D + f;
                    *y += lr_yj * (err * qif / sqrt_Iu - reg_yj * *y);
                }
            }
        }
    }
    return sse;
}

