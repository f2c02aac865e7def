/*
 * oracle_backend.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Implements the include/medgp_cuda.h entry points the host classes call, on top of the FP64
 * CPU oracle, so that tests/ can exercise the HOST logic (optimiser steppers, prior
 * adjustment, file formats, front-ends) in the GPU-less build container.  It is linked only
 * into the test executables under oracle/_build/ (oracle/Makefile: host_on_oracle); the
 * shipped binaries in medgp_b200/host/ link libmedgp_cuda.so and have no CPU path.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/medgp_cuda.h"

int medgp_oracle_nlml_grad(int Q, int D, int R, double pi, int n, const int32_t *meta, const float *x,
                           const float *y, const double *theta, int want_grad, int grad_mode,
                           double *nlml, double *grad, int *status);
int medgp_oracle_predict(int Q, int D, int R, double pi, int n, const int32_t *meta, const float *x,
                         const float *y, const double *theta, int m, const int32_t *meta_star,
                         const float *x_star, double *mean, double *var, int *status);
void medgp_oracle_force_fail(int attempts);
int medgp_oracle_factors(int Q, int D, int R, double pi, int n, const int32_t *meta, const float *x, const float *y,
                         const double *theta, double *alpha, double *Linv, double *nlml, int *status);

typedef struct { int n; int32_t *meta; float *x, *y; } series_t;
struct medgp_ctx { int Q, D, R, P; double pi; series_t *s; int ns, cap; };

int medgp_cuda_create(medgp_ctx **out, int device, size_t workspace_bytes)
{
    (void)device; (void)workspace_bytes;
    *out = (medgp_ctx *)calloc(1, sizeof(medgp_ctx));
    /* same test hook as the CUDA library: the first attempts of every factorisation count as failed */
    if (getenv("MEDGP_FORCE_FAIL")) medgp_oracle_force_fail(atoi(getenv("MEDGP_FORCE_FAIL")));
    return MEDGP_OK;
}
void medgp_cuda_destroy(medgp_ctx *c) { if (c) { medgp_cuda_clear_series(c); free(c->s); free(c); } }
const char *medgp_cuda_last_error(const medgp_ctx *c) { (void)c; return "oracle backend (tests only)"; }
int medgp_cuda_model(medgp_ctx *c, int Q, int D, int R, double pi)
{
    c->Q = Q; c->D = D; c->R = R; c->pi = pi; c->P = D + Q * (D * R + 2 + D);
    return MEDGP_OK;
}
int medgp_cuda_num_hyp(const medgp_ctx *c) { return c->P; }
int medgp_cuda_add_series(medgp_ctx *c, int n, const int32_t *meta, const float *x, const float *y, int *id)
{
    int slot = -1;
    for (int i = 0; i < c->ns; i++) if (c->s[i].n == 0) { slot = i; break; }
    if (slot < 0) {
        if (c->ns == c->cap) { c->cap = c->cap ? 2 * c->cap : 64; c->s = (series_t *)realloc(c->s, sizeof(series_t) * c->cap); }
        slot = c->ns++;
    }
    series_t *s = &c->s[slot];
    s->n = n;
    s->meta = (int32_t *)malloc(sizeof(int32_t) * n); memcpy(s->meta, meta, sizeof(int32_t) * n);
    s->x = (float *)malloc(sizeof(float) * n); memcpy(s->x, x, sizeof(float) * n);
    s->y = (float *)malloc(sizeof(float) * n); memcpy(s->y, y, sizeof(float) * n);
    *id = slot;
    return MEDGP_OK;
}
int medgp_cuda_add_series_batch(medgp_ctx *c, int count, const int *n, const int32_t *meta, const float *x,
                                const float *y, int order, int *ids)
{
    size_t pos = 0;
    (void)order;
    for (int b = 0; b < count; b++) {
        int rc = medgp_cuda_add_series(c, n[b], meta + pos, x + pos, y + pos, &ids[b]);
        if (rc) return rc;
        pos += (size_t)n[b];
    }
    return MEDGP_OK;
}
int medgp_cuda_free_series(medgp_ctx *c, int id)
{
    if (id < 0 || id >= c->ns || c->s[id].n == 0) return MEDGP_ERR_ARG;
    free(c->s[id].meta); free(c->s[id].x); free(c->s[id].y);
    memset(&c->s[id], 0, sizeof(series_t));
    return MEDGP_OK;
}
int medgp_cuda_export_factors(medgp_ctx *c, int sid, const double *theta, float *alpha, float *Linv, double *nlml, int *status)
{
    const series_t *s = &c->s[sid];
    const size_t n = (size_t)s->n;
    double *a = (double *)malloc(sizeof(double) * n), *X = (double *)malloc(sizeof(double) * n * n);
    int rc = medgp_oracle_factors(c->Q, c->D, c->R, c->pi, s->n, s->meta, s->x, s->y, theta, a, X, nlml, status);
    if (rc == 0) {
        for (size_t i = 0; i < n; i++) alpha[i] = (float)a[i];
        for (size_t i = 0; i < n * n; i++) Linv[i] = (float)X[i];
    } else {
        *nlml = NAN;
    }
    free(a); free(X);
    return MEDGP_OK;
}
int medgp_cuda_free_series_batch(medgp_ctx *c, int count, const int *ids)
{
    for (int b = 0; b < count; b++) medgp_cuda_free_series(c, ids[b]);
    return MEDGP_OK;
}
int medgp_cuda_clear_series(medgp_ctx *c)
{
    for (int i = 0; i < c->ns; i++) if (c->s[i].n) medgp_cuda_free_series(c, i);
    c->ns = 0;
    return MEDGP_OK;
}
int medgp_cuda_nlml_grad(medgp_ctx *c, int batch, const int *sid, const double *theta, int want_grad,
                         double *nlml, double *grad, int *status)
{
    for (int b = 0; b < batch; b++) {
        const series_t *s = &c->s[sid[b]];
        int rc = medgp_oracle_nlml_grad(c->Q, c->D, c->R, c->pi, s->n, s->meta, s->x, s->y, theta + (size_t)b * c->P,
                                        want_grad, 0, &nlml[b], want_grad ? grad + (size_t)b * c->P : NULL, &status[b]);
        if (rc != 0) nlml[b] = NAN;
    }
    return MEDGP_OK;
}
int medgp_cuda_predict(medgp_ctx *c, int batch, const int *sid, const double *theta, const int *off,
                       const int32_t *meta_star, const float *x_star, double *mean, double *var, int *status)
{
    for (int b = 0; b < batch; b++) {
        const series_t *s = &c->s[sid[b]];
        medgp_oracle_predict(c->Q, c->D, c->R, c->pi, s->n, s->meta, s->x, s->y, theta + (size_t)b * c->P,
                             off[b + 1] - off[b], meta_star + off[b], x_star + off[b], mean + off[b], var + off[b],
                             &status[b]);
    }
    return MEDGP_OK;
}
int medgp_cuda_add_series_ordered(medgp_ctx *c, int n, const int32_t *meta, const float *x, const float *y,
                                  int order, int *id)
{
    (void)order;  /* the oracle keeps the caller's order; results do not depend on it */
    return medgp_cuda_add_series(c, n, meta, x, y, id);
}
/* The reference's loop restated literally (main_one_test.cpp:269-444 without updates): one fit per
 * observation on "earlier points + the other points of the same time stamp".  The CUDA library
 * gets the same numbers from one factorisation per series. */
int medgp_cuda_predict_online(medgp_ctx *c, int batch, const int *sid, const double *theta, double *mean,
                              double *var, int *status)
{
    size_t o = 0;
    for (int b = 0; b < batch; b++) {
        const series_t *s = &c->s[sid[b]];
        const int n = s->n;
        int32_t *tm = (int32_t *)malloc(sizeof(int32_t) * n);
        float *tx = (float *)malloc(sizeof(float) * n), *ty = (float *)malloc(sizeof(float) * n);
        status[b] = 0;
        for (int j = 0; j < n; j++) {
            int m = 0;
            for (int pass = 0; pass < 2; pass++)
                for (int i = 0; i < n; i++)
                    if (pass == 0 ? s->x[i] < s->x[j] : (s->x[i] == s->x[j] && i != j)) {
                        tm[m] = s->meta[i]; tx[m] = s->x[i]; ty[m] = s->y[i]; m++;
                    }
            int st = 0;
            if (m == 0) {  /* prior: zero mean, k** + sigma^2, from a far-away dummy point */
                const float far = s->x[j] + 1e9f;
                medgp_oracle_predict(c->Q, c->D, c->R, c->pi, 1, &s->meta[j], &far, &s->y[j], theta + (size_t)b * c->P, 1,
                                     &s->meta[j], &s->x[j], &mean[o + j], &var[o + j], &st);
                mean[o + j] = 0.0;
            } else {
                medgp_oracle_predict(c->Q, c->D, c->R, c->pi, m, tm, tx, ty, theta + (size_t)b * c->P, 1, &s->meta[j],
                                     &s->x[j], &mean[o + j], &var[o + j], &st);
            }
            if (st != 0) status[b] = -1;  /* no jitter on this path */
        }
        free(tm); free(tx); free(ty);
        o += (size_t)n;
    }
    return MEDGP_OK;
}
int medgp_cuda_sync(medgp_ctx *c) { (void)c; return MEDGP_OK; }
int medgp_cuda_host_alloc(medgp_ctx *c, size_t bytes, void **p) { (void)c; *p = malloc(bytes ? bytes : 1); return *p ? MEDGP_OK : MEDGP_ERR_NOMEM; }
int medgp_cuda_host_free(medgp_ctx *c, void *p) { (void)c; free(p); return MEDGP_OK; }
