/* oracle/svbfm_oracle.c -- TEST INFRASTRUCTURE ONLY (see svbfm_oracle.h).
 *
 * Sequential CPU restatement of the reference's hot path. It deliberately keeps the reference's
 * data structures (per-row caches e,q / t,q,z; per-column ascending sweep; CSC built by a counting
 * transpose) and expression order so that, compiled without FMA contraction, it reproduces the
 * reference binary's printed numbers for the same seed. The product (CUDA) restructures all of this;
 * this file is what the product is checked against.
 *
 * Citations are to the reference tree, e.g. "vb.h:577" = src/libfm/src/fm_learn_vb.h line 577.
 *   vb.h    = src/libfm/src/fm_learn_vb.h            vbs.h  = src/libfm/src/fm_learn_vb_simultaneous.h
 *   vbo.h   = src/libfm/src/fm_learn_vb_online.h     vbos.h = src/libfm/src/fm_learn_vb_online_simultaneous.h
 *   mcmc.h  = src/libfm/src/fm_learn_mcmc.h          mcmcs.h= src/libfm/src/fm_learn_mcmc_simultaneous.h
 *   Data.h  = src/libfm/src/Data.h   random.h/matrix.h/fmatrix.h = src/util/...   libfm.cpp = src/libfm/libfm.cpp
 */
#define _GNU_SOURCE
#include "svbfm_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

/* ------------------------------------------------------------------ RNG (random.h:118-176) */
double orc_ran_uniform(void) { return rand() / ((double)RAND_MAX + 1); }        /* random.h:174-176 */

double orc_ran_gaussian(void) {                                                   /* random.h:150-164 (Leva) */
    double u, v, x, y, Q;
    do {
        do { u = orc_ran_uniform(); } while (u == 0.0);
        v = 1.7156 * (orc_ran_uniform() - 0.5);
        x = u - 0.449871;
        y = fabs(v) + 0.386595;
        Q = x * x + y * (0.19600 * y - 0.25472 * x);
        if (Q < 0.27597) break;
    } while ((Q > 0.27846) || ((v * v) > (-4.0 * u * u * log(u))));
    return v / u;
}

static double ran_gaussian_ms(double mean, double stdev) {                        /* random.h:166-172 */
    if ((stdev == 0.0) || isnan(stdev)) return mean;
    return mean + stdev * orc_ran_gaussian();
}

double orc_ran_gamma(double alpha) {                                              /* random.h:118-144 */
    if (alpha < 1.0) {
        double u;
        do { u = orc_ran_uniform(); } while (u == 0.0);
        return orc_ran_gamma(alpha + 1.0) * pow(u, 1.0 / alpha);
    } else {
        double d, c, x, v, u;
        d = alpha - 1.0 / 3.0;
        c = 1.0 / sqrt(9.0 * d);
        do {
            do { x = orc_ran_gaussian(); v = 1.0 + c * x; } while (v <= 0.0);
            v = v * v * v;
            u = orc_ran_uniform();
        } while ((u >= (1.0 - 0.0331 * (x * x) * (x * x))) && (log(u) >= (0.5 * x * x + d * (1.0 - v + log(v)))));
        return d * v;
    }
}
static double ran_gamma_ab(double alpha, double beta) { return orc_ran_gamma(alpha) / beta; } /* random.h:146-148 */

/* ---- classification (-task c): the reference's own erf approximation, cdf and truncated normals */
double orc_erf(double x) {                                                        /* random.h:47-61 (Abramowitz-Stegun 7.1.26) */
    double t = (x >= 0) ? 1.0 / (1.0 + 0.3275911 * x) : 1.0 / (1.0 - 0.3275911 * x);
    double result = 1.0 - (t * (0.254829592 + t * (-0.284496736 + t * (1.421413741 + t * (-1.453152027 + t * 1.061405429))))) * exp(-x * x);
    return (x >= 0) ? result : -result;
}
double orc_cdf_gaussian(double x) { return 0.5 + 0.5 * orc_erf(0.707106781 * x); }   /* random.h:67-69 */
static double ran_exp(void) { return -log(1 - orc_ran_uniform()); }               /* random.h:178-180 */
static double ran_left_tgaussian0(double left) {                                  /* random.h:72-102 */
    if (left <= 0.0) {                                                            /* naive: acceptance probability > 0.5 */
        double result;
        do { result = orc_ran_gaussian(); } while (result < left);
        return result;
    }
    double alpha_star = 0.5 * (left + sqrt(left * left + 4.0));                   /* Robert: translated exponential proposal */
    for (;;) {
        double z = ran_exp() / alpha_star + left;
        double d = z - alpha_star;
        d = exp(-(d * d) / 2);
        double u = orc_ran_uniform();
        if (u < d) return z;
    }
}
double orc_ran_left_tgaussian(double left, double mean, double stdev) { return mean + stdev * ran_left_tgaussian0((left - mean) / stdev); }   /* random.h:104-106 */
double orc_ran_right_tgaussian(double right, double mean, double stdev) { return mean + stdev * (-ran_left_tgaussian0(-((right - mean) / stdev))); } /* random.h:108-114 */

/* ------------------------------------------------------------------ sparse containers */
typedef struct spm {          /* sparse_row[] + sparse_entry[] (fmatrix.h:36-44) */
    uint32_t n;               /* number of rows of THIS matrix (cases for CSR, features for CSC) */
    uint64_t *ptr;            /* n+1 */
    uint32_t *id;
    float *val;
} spm;

static void spm_free(spm *m) { free(m->ptr); free(m->id); free(m->val); memset(m, 0, sizeof(*m)); }

/* Data::create_data_t (Data.h:457-509): counting transpose, row order preserved inside each column */
static int spm_transpose(const spm *in, uint32_t n_out, spm *out) {
    uint64_t nnz = in->ptr[in->n];
    out->n = n_out;
    out->ptr = (uint64_t *)calloc((size_t)n_out + 1, sizeof(uint64_t));
    out->id = (uint32_t *)malloc((nnz ? nnz : 1) * sizeof(uint32_t));
    out->val = (float *)malloc((nnz ? nnz : 1) * sizeof(float));
    if (!out->ptr || !out->id || !out->val) return -1;
    for (uint64_t p = 0; p < nnz; p++) {
        if (in->id[p] >= n_out) return -2;
        out->ptr[in->id[p] + 1]++;
    }
    for (uint32_t j = 0; j < n_out; j++) out->ptr[j + 1] += out->ptr[j];
    uint64_t *fill = (uint64_t *)malloc(((size_t)n_out + 1) * sizeof(uint64_t));
    memcpy(fill, out->ptr, ((size_t)n_out + 1) * sizeof(uint64_t));
    for (uint32_t i = 0; i < in->n; i++)
        for (uint64_t p = in->ptr[i]; p < in->ptr[i + 1]; p++) {
            uint64_t d = fill[in->id[p]]++;
            out->id[d] = i;
            out->val[d] = in->val[p];
        }
    free(fill);
    return 0;
}

static int spm_copy_csr(spm *dst, uint32_t n_rows, const uint64_t *rowptr, const uint32_t *col, const float *val) {
    uint64_t nnz = rowptr[n_rows];
    dst->n = n_rows;
    dst->ptr = (uint64_t *)malloc(((size_t)n_rows + 1) * sizeof(uint64_t));
    dst->id = (uint32_t *)malloc((nnz ? nnz : 1) * sizeof(uint32_t));
    dst->val = (float *)malloc((nnz ? nnz : 1) * sizeof(float));
    if (!dst->ptr || !dst->id || !dst->val) return -1;
    memcpy(dst->ptr, rowptr, ((size_t)n_rows + 1) * sizeof(uint64_t));
    memcpy(dst->id, col, nnz * sizeof(uint32_t));
    memcpy(dst->val, val, nnz * sizeof(float));
    return 0;
}

typedef struct split {
    uint32_t n_cases, n_feat;
    spm csr, csc;             /* data, data_t */
    float *y;
    float min_target, max_target;
} split;

/* ------------------------------------------------------------------ handle */
struct orc {
    int method, K, k0, k1;
    int task;                                /* 0 regression, 1 binary classification (fm_learn.h:67-68; mcmc only) */
    uint32_t D, G;
    uint32_t *attr_group, *n_per_group;      /* DataMetaInfo (Data.h:35-69) */
    split sp[2];
    double min_target, max_target;           /* from TRAIN (libfm.cpp:332-333) */
    /* model (fm_model.h:35-64) -- for mcmc these ARE the parameters */
    double w0, *w, *v;                       /* v row-major [K][D] */
    /* VB hyper + variational params (vb.h:36-46) */
    double alpha, sigma_0, *sigma_w, *sigma_v;   /* sigma_v(g,f) -> [g*K+f] */
    double mu_0_dash, sigma_0_dash, *mu_w, *sg_w, *mu_v, *sg_v;
    /* caches (mcmc.h:52-55, vb.h:17-21) */
    double *e, *q, *t, *tq, *tz;             /* train */
    double *e_test, *q_test;
    double *pred_this, *pred_sum_all;
    uint32_t iter;
    uint32_t nan_inf;
    /* mcmc (mcmc.h:78-95) */
    double alpha_0, gamma_0, beta_0, mu_0, w0_mean_0;
    double *w_mu, *w_lambda, *v_mu, *v_lambda;   /* v_*(g,f) -> [g*K+f] */
    int do_sample, do_multilevel;
    double reg0;
    double *grp_tmp;
    /* vb_online (vbo.h:23-35) */
    uint32_t num_batch, size_except_last, *shuffle, *col_count, *t_wj, *t_vj, t_w0, t0_w0, t0_wj, t0_vj;
    double lamda, new_w0, *new_wj, *new_vj;
    double nat_mu_0, nat_sg_0, *nat_mu_w, *nat_sg_w, *nat_mu_v, *nat_sg_v;
    double last_free_energy; int have_fe;
};

orc_t *orc_create(int method, uint32_t D, int K, int k0, int k1) {
    orc_t *h = (orc_t *)calloc(1, sizeof(orc_t));
    if (!h) return NULL;
    h->method = method; h->D = D; h->K = K; h->k0 = k0 != 0; h->k1 = k1 != 0;
    h->G = 1;
    h->attr_group = (uint32_t *)calloc(D ? D : 1, sizeof(uint32_t));
    h->n_per_group = (uint32_t *)calloc(1, sizeof(uint32_t));
    h->n_per_group[0] = D;                     /* Data.h:42-48 */
    h->do_sample = 1; h->do_multilevel = 1;    /* libfm.cpp:304-305 */
    h->num_batch = 50;                         /* libfm.cpp:320 */
    return h;
}

static void split_free(split *s) { spm_free(&s->csr); spm_free(&s->csc); free(s->y); memset(s, 0, sizeof(*s)); }

void orc_destroy(orc_t *h) {
    if (!h) return;
    split_free(&h->sp[0]); split_free(&h->sp[1]);
    free(h->attr_group); free(h->n_per_group);
    free(h->w); free(h->v); free(h->sigma_w); free(h->sigma_v);
    free(h->mu_w); free(h->sg_w); free(h->mu_v); free(h->sg_v);
    free(h->e); free(h->q); free(h->t); free(h->tq); free(h->tz); free(h->e_test); free(h->q_test);
    free(h->pred_this); free(h->pred_sum_all);
    free(h->w_mu); free(h->w_lambda); free(h->v_mu); free(h->v_lambda); free(h->grp_tmp);
    free(h->shuffle); free(h->col_count); free(h->t_wj); free(h->t_vj); free(h->new_wj); free(h->new_vj);
    free(h->nat_mu_w); free(h->nat_sg_w); free(h->nat_mu_v); free(h->nat_sg_v);
    free(h);
}

int orc_set_split(orc_t *h, int s, uint32_t n_rows, uint32_t n_feat, const uint64_t *rowptr, const uint32_t *col,
                  const float *val, const float *y) {
    if (s < 0 || s > 1) return -1;
    split *sp = &h->sp[s];
    split_free(sp);
    sp->n_cases = n_rows; sp->n_feat = n_feat;
    if (spm_copy_csr(&sp->csr, n_rows, rowptr, col, val)) return -2;
    /* vb_online: batches are transposed with D rows (Data.h:453,511-563); test/train data_t have n_feat rows */
    if (spm_transpose(&sp->csr, n_feat, &sp->csc)) return -3;
    sp->y = (float *)malloc((n_rows ? n_rows : 1) * sizeof(float));
    memcpy(sp->y, y, n_rows * sizeof(float));
    sp->min_target = +FLT_MAX; sp->max_target = -FLT_MAX;          /* Data.h:181-182,200-201 */
    for (uint32_t i = 0; i < n_rows; i++) {
        if (y[i] < sp->min_target) sp->min_target = y[i];
        if (y[i] > sp->max_target) sp->max_target = y[i];
    }
    if (s == 0) { h->min_target = sp->min_target; h->max_target = sp->max_target; }
    return 0;
}

int orc_set_groups(orc_t *h, const uint32_t *attr_group, uint32_t n_groups) {      /* Data.h:49-61 */
    free(h->n_per_group);
    h->G = n_groups;
    h->n_per_group = (uint32_t *)calloc(n_groups, sizeof(uint32_t));
    for (uint32_t j = 0; j < h->D; j++) {
        if (attr_group[j] >= n_groups) return -1;
        h->attr_group[j] = attr_group[j];
        h->n_per_group[attr_group[j]]++;
    }
    return 0;
}

int orc_set_mcmc_options(orc_t *h, int do_sample, int do_multilevel) {
    h->do_sample = do_sample; h->do_multilevel = do_multilevel; return 0;
}
int orc_set_num_batch(orc_t *h, uint32_t nb) { h->num_batch = nb; return 0; }
int orc_set_task(orc_t *h, int task) { if (task != 0 && (task != 1 || h->method != ORC_MCMC)) return -1; h->task = task; return 0; }
/* -regular r0,r1,r2 for mcmc/als (libfm.cpp:367-405): call after orc_init */
int orc_set_regular(orc_t *h, double r0, double rw, double rv) {
    if (h->method != ORC_MCMC || !h->w_lambda) return -1;
    h->reg0 = r0;
    for (uint32_t g = 0; g < h->G; g++) h->w_lambda[g] = rw;
    for (size_t i = 0; i < (size_t)h->G * h->K; i++) h->v_lambda[i] = rv;
    return 0;
}

static double *dalloc(size_t n) { return (double *)calloc(n ? n : 1, sizeof(double)); }

/* ------------------------------------------------------------------ init */
int orc_init(orc_t *h, long seed, double init_stdev) {
    uint32_t D = h->D; int K = h->K; uint32_t G = h->G;
    srand((unsigned)seed);                                           /* libfm.cpp:123-124 */
    /* fm_model::init (fm_model.h:92-101): w0=0, w=0, v ~ N(init_mean=0, init_stdev) in [f][j] order */
    h->w0 = 0.0;
    h->w = dalloc(D); h->v = dalloc((size_t)K * D);
    for (int f = 0; f < K; f++)
        for (uint32_t j = 0; j < D; j++) h->v[(size_t)f * D + j] = ran_gaussian_ms(0.0, init_stdev);  /* matrix.h:342-348 */
    /* libfm.cpp:298/307/313: fm.w.init_normal(init_mean, init_stdev) for mcmc, vb and vb_online */
    for (uint32_t j = 0; j < D; j++) h->w[j] = ran_gaussian_ms(0.0, init_stdev);                      /* matrix.h:334-338 */
    h->grp_tmp = dalloc(G);
    if (h->method == ORC_MCMC) {                                     /* mcmc.h:1092-1117 + libfm.cpp:372-377 */
        h->alpha_0 = 1.0; h->gamma_0 = 1.0; h->beta_0 = 1.0; h->mu_0 = 0.0;
        h->alpha = 1; h->w0_mean_0 = 0.0; h->reg0 = 0.0;
        h->w_mu = dalloc(G); h->w_lambda = dalloc(G);
        h->v_mu = dalloc((size_t)G * K); h->v_lambda = dalloc((size_t)G * K);
        return 0;
    }
    /* fm_learn_vb::init (vb.h:685-712) / fm_learn_vb_online::init (vbo.h:668-765) */
    h->alpha = 1.0; h->sigma_0 = 1.0; h->mu_0_dash = 0.0; h->sigma_0_dash = 0.02;
    h->sigma_w = dalloc(G); h->sigma_v = dalloc((size_t)G * K);
    h->mu_w = dalloc(D); h->sg_w = dalloc(D);
    h->mu_v = dalloc((size_t)K * D); h->sg_v = dalloc((size_t)K * D);
    for (uint32_t g = 0; g < G; g++) h->sigma_w[g] = 1;
    for (size_t i = 0; i < (size_t)G * K; i++) h->sigma_v[i] = 1;
    if (h->method == ORC_VB_ONLINE) {                                /* vbo.h:683-726 (before the draws) */
        h->lamda = 0.5; h->t0_w0 = 1; h->t0_wj = 1; h->t0_vj = 1; h->t_w0 = 0;
        h->new_w0 = pow((double)(h->t0_w0 + h->t_w0), -h->lamda);
        h->new_wj = dalloc(D); h->new_vj = dalloc(D);
        h->t_wj = (uint32_t *)calloc(D ? D : 1, sizeof(uint32_t));
        h->t_vj = (uint32_t *)calloc(D ? D : 1, sizeof(uint32_t));
        h->col_count = (uint32_t *)calloc(D ? D : 1, sizeof(uint32_t));
        for (uint32_t j = 0; j < D; j++) { h->new_wj[j] = pow((double)(h->t0_wj + 0), -h->lamda); h->new_vj[j] = pow((double)(h->t0_vj + 0), -h->lamda); }
        h->nat_mu_0 = 0.0; h->nat_sg_0 = 1 / h->sigma_0_dash;
        const spm *tr = &h->sp[0].csr;                               /* vbo.h:704-726: count of each feature id in the train file */
        if (tr->ptr) for (uint64_t p = 0; p < tr->ptr[tr->n]; p++) h->col_count[tr->id[p]] += 1;
    }
    for (uint32_t j = 0; j < D; j++) h->mu_w[j] = 0.1 * ran_gaussian_ms(0, 1);                 /* vb.h:709, matrix.h:360-364 */
    for (uint32_t j = 0; j < D; j++) h->sg_w[j] = .02;                                           /* vb.h:710 */
    for (int f = 0; f < K; f++)
        for (uint32_t j = 0; j < D; j++) h->mu_v[(size_t)f * D + j] = 0.1 * ran_gaussian_ms(0, 1); /* vb.h:711, matrix.h:374-380 */
    for (size_t i = 0; i < (size_t)K * D; i++) h->sg_v[i] = .02;                                 /* vb.h:712 */
    if (h->method == ORC_VB_ONLINE) {                                /* vbo.h:750-765 */
        h->nat_mu_w = dalloc(D); h->nat_sg_w = dalloc(D);
        h->nat_mu_v = dalloc((size_t)K * D); h->nat_sg_v = dalloc((size_t)K * D);
        for (uint32_t j = 0; j < D; j++) { h->nat_mu_w[j] = h->mu_w[j]; h->nat_mu_w[j] /= 0.02; h->nat_sg_w[j] = 1 / h->sg_w[j]; }
        for (size_t i = 0; i < (size_t)K * D; i++) { h->nat_mu_v[i] = h->mu_v[i]; h->nat_mu_v[i] /= 0.02; h->nat_sg_v[i] = 1 / h->sg_v[i]; }
    }
    return 0;
}

/* ------------------------------------------------------------------ prediction (vb.h:70-203 == vbo.h:71-203; mcmc.h:117-348 w/o relations) */
static void predict_eterms(const orc_t *h, const split *sp, const double *vmat, const double *wvec, double w0,
                           double *e, double *q) {
    const spm *t = &sp->csc;
    uint32_t D = h->D;
    for (uint32_t i = 0; i < sp->n_cases; i++) { e[i] = 0.0; q[i] = 0.0; }
    for (int f = 0; f < h->K; f++) {                                 /* (1) 1/2 sum_f (sum_j v_jf x_j)^2 */
        const double *v = vmat + (size_t)f * D;
        for (uint32_t j = 0; j < t->n; j++) {
            double v_if = v[j];
            for (uint64_t p = t->ptr[j]; p < t->ptr[j + 1]; p++) q[t->id[p]] += v_if * t->val[p];
        }
        for (uint32_t c = 0; c < sp->n_cases; c++) { double q_all = q[c]; e[c] += 0.5 * q_all * q_all; q[c] = 0.0; }
    }
    for (int f = 0; f < h->K; f++) {                                 /* (2) -1/2 sum_f sum_j v_jf^2 x_j^2 */
        const double *v = vmat + (size_t)f * D;
        for (uint32_t j = 0; j < t->n; j++) {
            double v_if = v[j];
            for (uint64_t p = t->ptr[j]; p < t->ptr[j + 1]; p++) { float x = t->val[p]; q[t->id[p]] -= 0.5 * v_if * v_if * x * x; }
        }
    }
    if (h->k1) {                                                     /* (3) + sum_j w_j x_j */
        for (uint32_t j = 0; j < t->n; j++) {
            double w_i = wvec[j];
            for (uint64_t p = t->ptr[j]; p < t->ptr[j + 1]; p++) q[t->id[p]] += w_i * t->val[p];
        }
    }
    for (uint32_t c = 0; c < sp->n_cases; c++) {                     /* merge */
        double q_all = q[c];
        e[c] = e[c] + q_all;
        if (h->k0) e[c] += w0;
        q[c] = 0.0;
    }
}

/* vb.h:207-312 */
static void predict_tterms(const orc_t *h, const split *sp, double *t_, double *tq, double *tz) {
    const spm *t = &sp->csc;
    uint32_t D = h->D;
    for (uint32_t i = 0; i < sp->n_cases; i++) { tq[i] = 0.0; tz[i] = 0.0; t_[i] = 0.0; }
    for (int f = 0; f < h->K; f++) {
        const double *v = h->mu_v + (size_t)f * D, *vs = h->sg_v + (size_t)f * D;
        for (uint32_t j = 0; j < t->n; j++) {
            double v_if = v[j], v_if_sigma = vs[j];
            for (uint64_t p = t->ptr[j]; p < t->ptr[j + 1]; p++) {
                float x = t->val[p]; uint32_t c = t->id[p];
                tq[c] += v_if * x * v_if * x;
                tz[c] += v_if_sigma * x * x;
            }
        }
        for (uint32_t c = 0; c < sp->n_cases; c++) {
            double q_all = tq[c], z_all = tz[c];
            t_[c] += (0.5 * z_all * z_all + z_all * q_all);
            tq[c] = 0.0; tz[c] = 0.0;
        }
    }
    for (int f = 0; f < h->K; f++) {
        const double *v = h->mu_v + (size_t)f * D, *vs = h->sg_v + (size_t)f * D;
        for (uint32_t j = 0; j < t->n; j++) {
            double v_if = v[j], v_if_sigma = vs[j];
            for (uint64_t p = t->ptr[j]; p < t->ptr[j + 1]; p++) {
                float x = t->val[p]; uint32_t c = t->id[p];
                tq[c] -= (v_if * v_if * x * x * x * x * v_if_sigma + 0.5 * x * x * x * x * v_if_sigma * v_if_sigma);
            }
        }
    }
    if (h->k1) {
        for (uint32_t j = 0; j < t->n; j++) {
            double w_i = h->sg_w[j];
            for (uint64_t p = t->ptr[j]; p < t->ptr[j + 1]; p++) { float x = t->val[p]; tq[t->id[p]] += w_i * x * x; }
        }
    }
    for (uint32_t c = 0; c < sp->n_cases; c++) {
        double q_all = tq[c];
        t_[c] = t_[c] + q_all;
        if (h->k0) t_[c] += h->sigma_0_dash;
        tq[c] = 0.0;
    }
}

/* ------------------------------------------------------------------ VB (vb.h:354-681) */
static void vb_add_main_q(orc_t *h, const spm *t, int f) {            /* vb.h:354-381 */
    const double *v = h->mu_v + (size_t)f * h->D, *vs = h->sg_v + (size_t)f * h->D;
    for (uint32_t j = 0; j < t->n; j++) {
        double v_if = v[j], v_if_sigma = vs[j];
        for (uint64_t p = t->ptr[j]; p < t->ptr[j + 1]; p++) {
            float x = t->val[p]; uint32_t c = t->id[p];
            h->q[c] += v_if * x;
            h->tq[c] += v_if_sigma * x * x;
            h->tz[c] += v_if * v_if * x * x;
        }
    }
}

static void vb_update_w0(orc_t *h, uint32_t n) {                      /* vb.h:504-525 */
    double sigma_old = h->sigma_0_dash;
    h->sigma_0_dash = 1.0 / (h->sigma_0 + n * h->alpha);
    double w0_temp = 0.0, mu_old = h->mu_0_dash;
    for (uint32_t i = 0; i < n; i++) w0_temp += h->e[i] + h->mu_0_dash;
    h->mu_0_dash = h->sigma_0_dash * h->alpha * w0_temp;
    for (uint32_t i = 0; i < n; i++) {
        h->e[i] = h->e[i] + (mu_old - h->mu_0_dash);
        h->t[i] = h->t[i] + (h->sigma_0_dash - sigma_old);
    }
}

static void vb_update_w(orc_t *h, double *mu, double *sigma, double sigma_w, const spm *t, uint32_t j) { /* vb.h:527-574 */
    double w_sigma_sqr = 0, w_mean = 0, mu_old = *mu, sigma_old = *sigma;
    for (uint64_t p = t->ptr[j]; p < t->ptr[j + 1]; p++) {
        float x = t->val[p];
        w_mean += x * (h->e[t->id[p]] + x * *mu);
        w_sigma_sqr += x * x;
    }
    *sigma = (double)1.0 / (sigma_w + h->alpha * w_sigma_sqr);
    *mu = *sigma * h->alpha * w_mean;
    if (isnan(*sigma) || isinf(*sigma)) { h->nan_inf++; *sigma = sigma_old; }
    if (isnan(*mu) || isinf(*mu)) { h->nan_inf++; *mu = mu_old; return; }
    for (uint64_t p = t->ptr[j]; p < t->ptr[j + 1]; p++) {
        double hh = t->val[p]; uint32_t c = t->id[p];
        h->e[c] += hh * (mu_old - *mu);
        h->t[c] += hh * hh * (*sigma - sigma_old);
    }
}

static void vb_update_v(orc_t *h, double *mu, double *sigma, double sigma_v_g, const spm *t, uint32_t j) { /* vb.h:577-644 */
    double v_sigma_sqr = 0, v_mean = 0, mu_old = *mu, sigma_old = *sigma;
    for (uint64_t p = t->ptr[j]; p < t->ptr[j + 1]; p++) {
        float x = t->val[p]; uint32_t c = t->id[p];
        double hh = h->q[c] - x * *mu;
        double h1 = h->tq[c] - x * x * *sigma;
        v_mean += x * hh * (h->e[c] + x * *mu * hh);
        v_sigma_sqr += x * x * hh * hh + x * x * h1;
    }
    *sigma = (double)1.0 / (sigma_v_g + h->alpha * v_sigma_sqr);
    *mu = *sigma * h->alpha * v_mean;
    if (isnan(*sigma) || isinf(*sigma)) { *sigma = sigma_old; h->nan_inf++; }
    if (isnan(*mu) || isinf(*mu)) { h->nan_inf++; *mu = mu_old; return; }
    for (uint64_t p = t->ptr[j]; p < t->ptr[j + 1]; p++) {
        float x = t->val[p]; uint32_t c = t->id[p];
        double hh = x * (h->q[c] - x * mu_old);
        double h1 = x * x * (h->tq[c] - x * x * sigma_old);
        double h2 = x * x * (h->tz[c] - x * x * mu_old * mu_old);
        h->q[c] += x * (*mu - mu_old);
        h->tq[c] += x * x * (*sigma - sigma_old);
        h->tz[c] += x * x * (*mu * *mu - mu_old * mu_old);
        h->e[c] += hh * (mu_old - *mu);
        h->t[c] += (h1 + h2) * (*sigma - sigma_old);
        h->t[c] += h1 * (*mu * *mu - mu_old * mu_old);
    }
}

static double vb_free_energy(orc_t *h, uint32_t n) {                 /* vb.h:646-681 (== vbo.h:629-664) */
    double temp = 0.0, fe = 0.0;
    uint32_t D = h->D; int K = h->K;
    for (uint32_t i = 0; i < n; i++) temp += h->e[i] * h->e[i] + h->t[i];
    double temp1 = 2 * 3.14 * (1.0 / h->alpha);
    fe += -0.5 * h->alpha * temp - .5 * n * log(temp1);
    fe += -0.5 * h->sigma_0 * (h->mu_0_dash * h->mu_0_dash + h->sigma_0_dash) + 0.5 * log(h->sigma_0_dash * h->sigma_0) + .5;
    for (uint32_t i = 0; i < D; i++) {
        uint32_t g = h->attr_group[i];
        fe += -0.5 * h->sigma_w[g] * (h->mu_w[i] * h->mu_w[i] + h->sg_w[i]) + 0.5 * log(h->sg_w[i] * h->sigma_w[g]) + .5;
    }
    for (int f = 0; f < K; f++) {
        const double *v = h->mu_v + (size_t)f * D, *v1 = h->sg_v + (size_t)f * D;
        for (uint32_t i = 0; i < D; i++) {
            uint32_t g = h->attr_group[i];
            fe += -0.5 * h->sigma_v[g * K + f] * (v[i] * v[i] + v1[i]) + 0.5 * log(v1[i] * h->sigma_v[g * K + f]) + .5;
        }
    }
    h->last_free_energy = fe; h->have_fe = 1;
    return fe;
}

static void vb_update_all(orc_t *h) {                                /* vb.h:383-501 */
    const split *tr = &h->sp[0];
    const spm *t = &tr->csc;
    uint32_t n = tr->n_cases, D = h->D; int K = h->K;
    h->have_fe = 0;
    if (h->k0) vb_update_w0(h, n);
    if (h->k1)
        for (uint32_t j = 0; j < t->n; j++) vb_update_w(h, &h->mu_w[j], &h->sg_w[j], h->sigma_w[h->attr_group[j]], t, j);
    if (D > 0)
        for (int f = 0; f < K; f++) {
            for (uint32_t c = 0; c < n; c++) { h->q[c] = 0.0; h->tq[c] = 0.0; h->tz[c] = 0.0; }
            vb_add_main_q(h, t, f);
            double *v = h->mu_v + (size_t)f * D, *v1 = h->sg_v + (size_t)f * D;
            for (uint32_t j = 0; j < t->n; j++) vb_update_v(h, &v[j], &v1[j], h->sigma_v[h->attr_group[j] * K + f], t, j);
        }
    {                                                                /* alpha (vb.h:446-470) */
        double alpha_temp = 0.0;
        for (uint32_t i = 0; i < n; i++) alpha_temp += h->e[i] * h->e[i] + h->t[i];
        double alpha_old = h->alpha;
        h->alpha = (double)n / alpha_temp;
        if (isnan(h->alpha) || isinf(h->alpha)) { h->nan_inf++; h->alpha = alpha_old; return; }
    }
    h->sigma_0 = 1.0 / (h->mu_0_dash * h->mu_0_dash + h->sigma_0_dash);    /* vb.h:473 */
    for (uint32_t g = 0; g < h->G; g++) h->grp_tmp[g] = 0.0;               /* vb.h:475-483 */
    for (uint32_t i = 0; i < D; i++) h->grp_tmp[h->attr_group[i]] += h->mu_w[i] * h->mu_w[i] + h->sg_w[i];
    for (uint32_t g = 0; g < h->G; g++) h->sigma_w[g] = (double)h->n_per_group[g] / h->grp_tmp[g];
    for (int f = 0; f < K; f++) {                                          /* vb.h:486-498 */
        for (uint32_t g = 0; g < h->G; g++) h->grp_tmp[g] = 0.0;
        const double *v = h->mu_v + (size_t)f * D, *v1 = h->sg_v + (size_t)f * D;
        for (uint32_t i = 0; i < D; i++) h->grp_tmp[h->attr_group[i]] += v[i] * v[i] + v1[i];
        for (uint32_t g = 0; g < h->G; g++) h->sigma_v[g * K + f] = (double)h->n_per_group[g] / h->grp_tmp[g];
    }
    vb_free_energy(h, n);
}

/* _evaluate (vbs.h:261-279 == mcmcs.h:307-325) */
static double eval_rmse(const orc_t *h, const double *pred, const float *target, uint32_t n, double normalizer) {
    double _rmse = 0; uint32_t num = 0;
    for (uint32_t c = 0; c < n; c++) {
        double p = pred[c] * normalizer;
        p = fmin(h->max_target, p);
        p = fmax(h->min_target, p);
        double err = p - target[c];
        _rmse += err * err;
        num++;
    }
    return sqrt(_rmse / num);
}

/* ------------------------------------------------------------------ vb_online (vbo.h:354-627) */
static void vbo_update_w0(orc_t *h, uint32_t bs, uint32_t _size) {    /* vbo.h:471-497 */
    double sigma_dash = h->sigma_0_dash, w0_temp = 0.0, mu_dash = h->mu_0_dash;
    double mu_old = h->nat_mu_0, sigma_old = h->nat_sg_0, eta1 = 0.0, eta2 = 0.0;
    for (uint32_t i = 0; i < bs; i++) {
        w0_temp = h->e[i] + h->mu_0_dash;
        h->nat_sg_0 = ((1 - h->new_w0) * sigma_old) + h->new_w0 * (h->sigma_0 + _size * h->alpha);
        h->nat_mu_0 = ((1 - h->new_w0) * mu_old) + h->new_w0 * _size * h->alpha * w0_temp;
        eta1 += h->nat_mu_0; eta2 += h->nat_sg_0;
    }
    h->nat_mu_0 = eta1 / bs; h->nat_sg_0 = eta2 / bs;
    h->mu_0_dash = h->nat_mu_0 / h->nat_sg_0;
    h->sigma_0_dash = 1.0 / h->nat_sg_0;
    for (uint32_t i = 0; i < bs; i++) {
        h->e[i] = h->e[i] + (mu_dash - h->mu_0_dash);
        h->t[i] = h->t[i] + (h->sigma_0_dash - sigma_dash);
    }
}

static void vbo_update_w(orc_t *h, double *mu, double *sigma, double sigma_w, const spm *t, uint32_t col) { /* vbo.h:499-556 */
    double w_sigma_sqr = 0, w_mean = 0, mu_dash = *mu, sigma_dash = *sigma, eta1 = 0.0, eta2 = 0.0;
    double mu_old = h->nat_mu_w[col], sigma_old = h->nat_sg_w[col];
    uint32_t size = (uint32_t)(t->ptr[col + 1] - t->ptr[col]);
    for (uint64_t p = t->ptr[col]; p < t->ptr[col + 1]; p++) {
        float x = t->val[p];
        w_mean = x * (h->e[t->id[p]] + x * *mu);
        w_sigma_sqr = x * x;
        h->nat_sg_w[col] = ((1 - h->new_wj[col]) * sigma_old) + h->new_wj[col] * (sigma_w + h->alpha * h->col_count[col] * w_sigma_sqr);
        h->nat_mu_w[col] = ((1 - h->new_wj[col]) * mu_old) + h->new_wj[col] * h->col_count[col] * h->alpha * w_mean;
        eta1 += h->nat_mu_w[col]; eta2 += h->nat_sg_w[col];
    }
    h->t_wj[col] += size;
    h->new_wj[col] = pow((double)(h->t0_wj + h->t_wj[col]), -h->lamda);
    h->nat_mu_w[col] = eta1 / size; h->nat_sg_w[col] = eta2 / size;
    *mu = h->nat_mu_w[col] / h->nat_sg_w[col];
    *sigma = 1 / h->nat_sg_w[col];
    if (isnan(*sigma) || isinf(*sigma)) { h->nan_inf++; *sigma = sigma_dash; }
    if (isnan(*mu) || isinf(*mu)) { h->nan_inf++; *mu = mu_dash; return; }
    for (uint64_t p = t->ptr[col]; p < t->ptr[col + 1]; p++) {
        double hh = t->val[p]; uint32_t c = t->id[p];
        h->e[c] += hh * (mu_dash - *mu);
        h->t[c] += hh * hh * (*sigma - sigma_dash);
    }
}

static void vbo_update_v(orc_t *h, int f, double *mu, double *sigma, double sigma_v_g, const spm *t, uint32_t col) { /* vbo.h:558-627 */
    double v_sigma_sqr = 0, v_mean = 0, mu_dash = *mu, sigma_dash = *sigma, eta1 = 0.0, eta2 = 0.0;
    size_t fc = (size_t)f * h->D + col;
    double mu_old = h->nat_mu_v[fc], sigma_old = h->nat_sg_v[fc];
    uint32_t size = (uint32_t)(t->ptr[col + 1] - t->ptr[col]);
    for (uint64_t p = t->ptr[col]; p < t->ptr[col + 1]; p++) {
        float x = t->val[p]; uint32_t c = t->id[p];
        double hh = h->q[c] - x * *mu;
        double h1 = h->tq[c] - x * x * *sigma;
        v_mean = x * hh * (h->e[c] + x * *mu * hh);
        v_sigma_sqr = x * x * hh * hh + x * x * h1;
        h->nat_sg_v[fc] = (1 - h->new_vj[col]) * sigma_old + h->new_vj[col] * (sigma_v_g + h->alpha * h->col_count[col] * v_sigma_sqr);
        h->nat_mu_v[fc] = ((1 - h->new_vj[col]) * mu_old) + h->new_vj[col] * h->col_count[col] * h->alpha * v_mean;
        eta1 += h->nat_mu_v[fc]; eta2 += h->nat_sg_v[fc];
    }
    h->nat_mu_v[fc] = eta1 / size; h->nat_sg_v[fc] = eta2 / size;
    *mu = h->nat_mu_v[fc] / h->nat_sg_v[fc];
    *sigma = 1 / h->nat_sg_v[fc];
    if (isnan(*sigma) || isinf(*sigma)) { *sigma = sigma_dash; h->nan_inf++; }
    if (isnan(*mu) || isinf(*mu)) { h->nan_inf++; *mu = mu_dash; return; }
    for (uint64_t p = t->ptr[col]; p < t->ptr[col + 1]; p++) {
        float x = t->val[p]; uint32_t c = t->id[p];
        double hh = x * (h->q[c] - x * mu_dash);
        double h1 = x * x * (h->tq[c] - x * x * sigma_dash);
        double h2 = x * x * (h->tz[c] - x * x * mu_dash * mu_dash);
        h->q[c] += x * (*mu - mu_dash);
        h->tq[c] += x * x * (*sigma - sigma_dash);
        h->tz[c] += x * x * (*mu * *mu - mu_dash * mu_dash);
        h->e[c] += hh * (mu_dash - *mu);
        h->t[c] += (h1 + h2) * (*sigma - sigma_dash);
        h->t[c] += h1 * (*mu * *mu - mu_dash * mu_dash);
    }
}

static void vbo_update_all(orc_t *h, const split *b, uint32_t _size) { /* vbo.h:354-468 */
    const spm *t = &b->csc;
    uint32_t bs = b->n_cases, D = h->D; int K = h->K;
    if (h->k0) vbo_update_w0(h, bs, _size);
    if (h->k1)
        for (uint32_t j = 0; j < t->n; j++) {
            if (t->ptr[j + 1] == t->ptr[j]) continue;
            vbo_update_w(h, &h->mu_w[j], &h->sg_w[j], h->sigma_w[h->attr_group[j]], t, j);
        }
    if (D > 0) {
        for (int f = 0; f < K; f++) {
            for (uint32_t c = 0; c < bs; c++) { h->q[c] = 0.0; h->tq[c] = 0.0; h->tz[c] = 0.0; }
            vb_add_main_q(h, t, f);
            double *v = h->mu_v + (size_t)f * D, *v1 = h->sg_v + (size_t)f * D;
            for (uint32_t j = 0; j < t->n; j++) {
                if (t->ptr[j + 1] == t->ptr[j]) continue;
                vbo_update_v(h, f, &v[j], &v1[j], h->sigma_v[h->attr_group[j] * K + f], t, j);
                if (f == 0) h->t_vj[j] += (uint32_t)(t->ptr[j + 1] - t->ptr[j]);
            }
        }
        for (uint32_t j = 0; j < t->n; j++) h->new_vj[j] = pow((double)(h->t0_vj + h->t_vj[j]), -h->lamda);
    }
    {
        double alpha_temp = 0.0;
        for (uint32_t i = 0; i < bs; i++) alpha_temp += h->e[i] * h->e[i] + h->t[i];
        double alpha_old = h->alpha;
        h->alpha = (1 - h->new_w0) * alpha_old + h->new_w0 * ((double)bs / alpha_temp);
        if (isnan(h->alpha) || isinf(h->alpha)) { h->nan_inf++; h->alpha = alpha_old; return; }
    }
    h->sigma_0 = (1 - h->new_w0) * h->sigma_0 + h->new_w0 * (1.0 / (h->mu_0_dash * h->mu_0_dash + h->sigma_0_dash));
    for (uint32_t g = 0; g < h->G; g++) h->grp_tmp[g] = 0.0;
    for (uint32_t i = 0; i < D; i++) h->grp_tmp[h->attr_group[i]] += h->mu_w[i] * h->mu_w[i] + h->sg_w[i];
    for (uint32_t g = 0; g < h->G; g++) h->sigma_w[g] = (1 - h->new_w0) * h->sigma_w[g] + h->new_w0 * ((double)h->n_per_group[g] / h->grp_tmp[g]);
    for (int f = 0; f < K; f++) {
        for (uint32_t g = 0; g < h->G; g++) h->grp_tmp[g] = 0.0;
        const double *v = h->mu_v + (size_t)f * D, *v1 = h->sg_v + (size_t)f * D;
        for (uint32_t i = 0; i < D; i++) h->grp_tmp[h->attr_group[i]] += v[i] * v[i] + v1[i];
        for (uint32_t g = 0; g < h->G; g++)
            h->sigma_v[g * K + f] = (1 - h->new_w0) * h->sigma_v[g * K + f] + h->new_w0 * ((double)h->n_per_group[g] / h->grp_tmp[g]);
    }
    h->t_w0 += 1;
    h->new_w0 = pow((double)(h->t0_w0 + h->t_w0), -h->lamda);
}

/* one epoch (vbos.h:66-288). The reference writes/re-parses text batch files; the row->batch rule and
 * the within-batch file order are what matter and are kept. */
static int vbo_epoch(orc_t *h, orc_stats *out) {
    const split *tr = &h->sp[0];
    uint32_t n = tr->n_cases, nb = h->num_batch;
    /* std::random_shuffle (libstdc++ stl_algo.h): for i in [1,n): swap(a[i], a[rand() % (i+1)])  (vbos.h:74) */
    for (uint32_t i = 1; i < n; i++) {
        uint32_t j = (uint32_t)(rand() % (i + 1));
        if (i != j) { uint32_t tmp = h->shuffle[i]; h->shuffle[i] = h->shuffle[j]; h->shuffle[j] = tmp; }
    }
    uint32_t *batch_of = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    uint32_t *cnt = (uint32_t *)calloc(nb + 1, sizeof(uint32_t));
    for (uint32_t r = 0; r < n; r++) {                               /* vbos.h:87-95 */
        uint32_t group = (uint32_t)ceil(((double)h->shuffle[r] / h->size_except_last));
        batch_of[r] = group - 1; cnt[group - 1]++;
    }
    h->have_fe = 0;
    double fe_last = 0; int have = 0;
    for (uint32_t j = 0; j < nb; j++) {                              /* vbos.h:103-157 */
        split b; memset(&b, 0, sizeof(b));
        uint32_t bs = cnt[j];
        uint64_t nnz = 0;
        for (uint32_t r = 0; r < n; r++) if (batch_of[r] == j) nnz += tr->csr.ptr[r + 1] - tr->csr.ptr[r];
        b.n_cases = bs; b.n_feat = h->D;
        b.csr.n = bs;
        b.csr.ptr = (uint64_t *)calloc((size_t)bs + 1, sizeof(uint64_t));
        b.csr.id = (uint32_t *)malloc((nnz ? nnz : 1) * sizeof(uint32_t));
        b.csr.val = (float *)malloc((nnz ? nnz : 1) * sizeof(float));
        b.y = (float *)malloc((bs ? bs : 1) * sizeof(float));
        uint32_t k = 0; uint64_t w = 0;
        for (uint32_t r = 0; r < n; r++) if (batch_of[r] == j) {
            for (uint64_t p = tr->csr.ptr[r]; p < tr->csr.ptr[r + 1]; p++) { b.csr.id[w] = tr->csr.id[p]; b.csr.val[w] = tr->csr.val[p]; w++; }
            b.y[k] = tr->y[r]; k++; b.csr.ptr[k] = w;
        }
        spm_transpose(&b.csr, h->D, &b.csc);                         /* Data.h:453,511-563 */
        free(h->e); free(h->q); free(h->t); free(h->tq); free(h->tz);
        h->e = dalloc(bs); h->q = dalloc(bs); h->t = dalloc(bs); h->tq = dalloc(bs); h->tz = dalloc(bs);
        if (bs > 0) {
            predict_eterms(h, &b, h->mu_v, h->mu_w, h->mu_0_dash, h->e, h->q);       /* vbos.h:120 */
            predict_tterms(h, &b, h->t, h->tq, h->tz);                               /* vbos.h:121 */
            for (uint32_t c = 0; c < bs; c++) h->e[c] = b.y[c] - h->e[c];            /* vbos.h:125-127 */
        }
        vbo_update_all(h, &b, n);                                                    /* vbos.h:141 */
        if (j + 1 == nb || j == 0) { fe_last = vb_free_energy(h, bs); have = 1; }   /* vbos.h:143-146 */
        split_free(&b);
    }
    free(batch_of); free(cnt);
    const split *te = &h->sp[1];
    predict_eterms(h, te, h->mu_v, h->mu_w, h->mu_0_dash, h->e_test, h->q_test);     /* vbos.h:190 */
    for (uint32_t c = 0; c < te->n_cases; c++) {                                     /* vbos.h:208-215 */
        double p = h->e_test[c];
        p = fmin(h->max_target, p); p = fmax(h->min_target, p);
        h->pred_this[c] = p;
    }
    out->test_rmse = eval_rmse(h, h->pred_this, te->y, te->n_cases, 1.0);            /* vbos.h:242 */
    out->train_stat = NAN;
    out->free_energy = fe_last; out->has_free_energy = have;
    out->alpha = h->alpha;
    return 0;
}

/* ------------------------------------------------------------------ MCMC (mcmc.h:384-1089, non-relation paths) */
static void mcmc_draw_alpha(orc_t *h, uint32_t n) {                  /* mcmc.h:901-929 */
    if (!h->do_multilevel) { h->alpha = h->alpha_0; return; }
    double alpha_n = h->alpha_0 + n, gamma_n = h->gamma_0;
    for (uint32_t i = 0; i < n; i++) gamma_n += h->e[i] * h->e[i];
    double alpha_old = h->alpha;
    h->alpha = ran_gamma_ab(alpha_n / 2.0, gamma_n / 2.0);
    if (isnan(h->alpha) || isinf(h->alpha)) { h->nan_inf++; h->alpha = alpha_old; }
}

static void mcmc_draw_w0(orc_t *h, uint32_t n) {                     /* mcmc.h:628-668 */
    double w0_mean = 0;
    for (uint32_t i = 0; i < n; i++) w0_mean += h->e[i] - h->w0;
    double w0_sigma_sqr = (double)1.0 / (h->reg0 + h->alpha * n);
    w0_mean = -w0_sigma_sqr * (h->alpha * w0_mean - h->w0_mean_0 * h->reg0);
    double w0_old = h->w0;
    if (h->do_sample) h->w0 = ran_gaussian_ms(w0_mean, sqrt(w0_sigma_sqr)); else h->w0 = w0_mean;
    if (isnan(h->w0) || isinf(h->w0)) { h->nan_inf++; h->w0 = w0_old; return; }
    for (uint32_t i = 0; i < n; i++) h->e[i] -= (w0_old - h->w0);
}

static void mcmc_draw_w(orc_t *h, double *w, double w_mu, double w_lambda, const spm *t, uint32_t j, int empty) { /* mcmc.h:671-718 */
    double w_sigma_sqr = 0, w_mean = 0;
    uint64_t p0 = empty ? 0 : t->ptr[j], p1 = empty ? 0 : t->ptr[j + 1];
    for (uint64_t p = p0; p < p1; p++) {
        float x = t->val[p];
        w_mean += x * (h->e[t->id[p]] - *w * x);
        w_sigma_sqr += x * x;
    }
    w_sigma_sqr = (double)1.0 / (w_lambda + h->alpha * w_sigma_sqr);
    w_mean = -w_sigma_sqr * (h->alpha * w_mean - w_mu * w_lambda);
    double w_old = *w;
    if (isnan(w_sigma_sqr) || isinf(w_sigma_sqr)) *w = 0.0;
    else if (h->do_sample) *w = ran_gaussian_ms(w_mean, sqrt(w_sigma_sqr));
    else *w = w_mean;
    if (isnan(*w) || isinf(*w)) { h->nan_inf++; *w = w_old; return; }
    for (uint64_t p = p0; p < p1; p++) { double hh = t->val[p]; h->e[t->id[p]] -= hh * (w_old - *w); }
}

static void mcmc_draw_v(orc_t *h, double *v, double v_mu, double v_lambda, const spm *t, uint32_t j, int empty) { /* mcmc.h:780-835 */
    double v_sigma_sqr = 0, v_mean = 0;
    uint64_t p0 = empty ? 0 : t->ptr[j], p1 = empty ? 0 : t->ptr[j + 1];
    for (uint64_t p = p0; p < p1; p++) {
        float x = t->val[p]; uint32_t c = t->id[p];
        double hh = x * (h->q[c] - x * *v);
        v_mean += hh * h->e[c];
        v_sigma_sqr += hh * hh;
    }
    v_mean -= *v * v_sigma_sqr;
    v_sigma_sqr = (double)1.0 / (v_lambda + h->alpha * v_sigma_sqr);
    v_mean = -v_sigma_sqr * (h->alpha * v_mean - v_mu * v_lambda);
    double v_old = *v;
    if (isnan(v_sigma_sqr) || isinf(v_sigma_sqr)) *v = 0.0;
    else if (h->do_sample) *v = ran_gaussian_ms(v_mean, sqrt(v_sigma_sqr));
    else *v = v_mean;
    if (isnan(*v) || isinf(*v)) { h->nan_inf++; *v = v_old; return; }
    for (uint64_t p = p0; p < p1; p++) {
        float x = t->val[p]; uint32_t c = t->id[p];
        double hh = x * (h->q[c] - x * v_old);
        h->q[c] -= x * (v_old - *v);
        h->e[c] -= hh * (v_old - *v);
    }
}

static void mcmc_draw_w_lambda(orc_t *h) {                           /* mcmc.h:970-1007 */
    if (!h->do_multilevel) return;
    for (uint32_t g = 0; g < h->G; g++) h->grp_tmp[g] = h->beta_0 * (h->w_mu[g] - h->mu_0) * (h->w_mu[g] - h->mu_0) + h->gamma_0;
    for (uint32_t i = 0; i < h->D; i++) { uint32_t g = h->attr_group[i]; h->grp_tmp[g] += (h->w[i] - h->w_mu[g]) * (h->w[i] - h->w_mu[g]); }
    for (uint32_t g = 0; g < h->G; g++) {
        double a = h->alpha_0 + h->n_per_group[g] + 1, old = h->w_lambda[g];
        if (h->do_sample) h->w_lambda[g] = ran_gamma_ab(a / 2.0, h->grp_tmp[g] / 2.0); else h->w_lambda[g] = a / h->grp_tmp[g];
        if (isnan(h->w_lambda[g]) || isinf(h->w_lambda[g])) { h->nan_inf++; h->w_lambda[g] = old; return; }
    }
}
static void mcmc_draw_w_mu(orc_t *h) {                               /* mcmc.h:931-968 */
    if (!h->do_multilevel) { for (uint32_t g = 0; g < h->G; g++) h->w_mu[g] = h->mu_0; return; }
    for (uint32_t g = 0; g < h->G; g++) h->grp_tmp[g] = 0.0;
    for (uint32_t i = 0; i < h->D; i++) h->grp_tmp[h->attr_group[i]] += h->w[i];
    for (uint32_t g = 0; g < h->G; g++) {
        h->grp_tmp[g] = (h->grp_tmp[g] + h->beta_0 * h->mu_0) / (h->n_per_group[g] + h->beta_0);
        double s2 = (double)1.0 / ((h->n_per_group[g] + h->beta_0) * h->w_lambda[g]), old = h->w_mu[g];
        if (h->do_sample) h->w_mu[g] = ran_gaussian_ms(h->grp_tmp[g], sqrt(s2)); else h->w_mu[g] = h->grp_tmp[g];
        if (isnan(h->w_mu[g]) || isinf(h->w_mu[g])) { h->nan_inf++; h->w_mu[g] = old; return; }
    }
}
static void mcmc_draw_v_lambda(orc_t *h) {                           /* mcmc.h:1051-1089 */
    if (!h->do_multilevel) return;
    int K = h->K; uint32_t D = h->D;
    for (int f = 0; f < K; f++) {
        for (uint32_t g = 0; g < h->G; g++) { double m = h->v_mu[g * K + f]; h->grp_tmp[g] = h->beta_0 * (m - h->mu_0) * (m - h->mu_0) + h->gamma_0; }
        for (uint32_t i = 0; i < D; i++) { uint32_t g = h->attr_group[i]; double d = h->v[(size_t)f * D + i] - h->v_mu[g * K + f]; h->grp_tmp[g] += d * d; }
        for (uint32_t g = 0; g < h->G; g++) {
            double a = h->alpha_0 + h->n_per_group[g] + 1, old = h->v_lambda[g * K + f];
            if (h->do_sample) h->v_lambda[g * K + f] = ran_gamma_ab(a / 2.0, h->grp_tmp[g] / 2.0); else h->v_lambda[g * K + f] = a / h->grp_tmp[g];
            if (isnan(h->v_lambda[g * K + f]) || isinf(h->v_lambda[g * K + f])) { h->nan_inf++; h->v_lambda[g * K + f] = old; return; }
        }
    }
}
static void mcmc_draw_v_mu(orc_t *h) {                               /* mcmc.h:1011-1049 */
    int K = h->K; uint32_t D = h->D;
    if (!h->do_multilevel) { for (size_t i = 0; i < (size_t)h->G * K; i++) h->v_mu[i] = h->mu_0; return; }
    for (int f = 0; f < K; f++) {
        for (uint32_t g = 0; g < h->G; g++) h->grp_tmp[g] = 0.0;
        for (uint32_t i = 0; i < D; i++) h->grp_tmp[h->attr_group[i]] += h->v[(size_t)f * D + i];
        for (uint32_t g = 0; g < h->G; g++) {
            h->grp_tmp[g] = (h->grp_tmp[g] + h->beta_0 * h->mu_0) / (h->n_per_group[g] + h->beta_0);
            double s2 = (double)1.0 / ((h->n_per_group[g] + h->beta_0) * h->v_lambda[g * K + f]), old = h->v_mu[g * K + f];
            if (h->do_sample) h->v_mu[g * K + f] = ran_gaussian_ms(h->grp_tmp[g], sqrt(s2)); else h->v_mu[g * K + f] = h->grp_tmp[g];
            if (isnan(h->v_mu[g * K + f]) || isinf(h->v_mu[g * K + f])) { h->nan_inf++; h->v_mu[g * K + f] = old; return; }
        }
    }
}

static void mcmc_draw_all(orc_t *h) {                                /* mcmc.h:411-623 */
    const split *tr = &h->sp[0];
    const spm *t = &tr->csc;
    uint32_t n = tr->n_cases, D = h->D; int K = h->K;
    mcmc_draw_alpha(h, n);
    if (h->k0) mcmc_draw_w0(h, n);
    if (h->k1) {
        mcmc_draw_w_lambda(h);
        mcmc_draw_w_mu(h);
        for (uint32_t j = 0; j < t->n; j++) { uint32_t g = h->attr_group[j]; mcmc_draw_w(h, &h->w[j], h->w_mu[g], h->w_lambda[g], t, j, 0); }
        for (uint32_t j = t->n; j < D; j++) { uint32_t g = h->attr_group[j]; mcmc_draw_w(h, &h->w[j], h->w_mu[g], h->w_lambda[g], t, j, 1); }
    }
    if (K > 0) { mcmc_draw_v_lambda(h); mcmc_draw_v_mu(h); }
    for (int f = 0; f < K; f++) {
        for (uint32_t c = 0; c < n; c++) h->q[c] = 0.0;
        double *v = h->v + (size_t)f * D;
        for (uint32_t j = 0; j < t->n; j++) {                        /* add_main_q mcmc.h:384-409 */
            double v_if = v[j];
            for (uint64_t p = t->ptr[j]; p < t->ptr[j + 1]; p++) h->q[t->id[p]] += v_if * t->val[p];
        }
        for (uint32_t j = 0; j < t->n; j++) { uint32_t g = h->attr_group[j]; mcmc_draw_v(h, &v[j], h->v_mu[g * K + f], h->v_lambda[g * K + f], t, j, 0); }
        for (uint32_t j = t->n; j < D; j++) { uint32_t g = h->attr_group[j]; mcmc_draw_v(h, &v[j], h->v_mu[g * K + f], h->v_lambda[g * K + f], t, j, 1); }
    }
}

/* ------------------------------------------------------------------ outer loops */
int orc_begin(orc_t *h) {
    const split *tr = &h->sp[0], *te = &h->sp[1];
    uint32_t n = tr->n_cases, nt = te->n_cases;
    h->e_test = dalloc(nt); h->q_test = dalloc(nt);
    h->pred_this = dalloc(nt); h->pred_sum_all = dalloc(nt);
    h->iter = 0;
    if (h->method == ORC_VB) {                                       /* vbs.h:37-44 */
        h->e = dalloc(n); h->q = dalloc(n); h->t = dalloc(n); h->tq = dalloc(n); h->tz = dalloc(n);
        predict_eterms(h, tr, h->mu_v, h->mu_w, h->mu_0_dash, h->e, h->q);
        predict_eterms(h, te, h->mu_v, h->mu_w, h->mu_0_dash, h->e_test, h->q_test);
        predict_tterms(h, tr, h->t, h->tq, h->tz);
        for (uint32_t c = 0; c < n; c++) h->e[c] = tr->y[c] - h->e[c];
    } else if (h->method == ORC_MCMC) {                              /* mcmcs.h:75-80 */
        h->e = dalloc(n); h->q = dalloc(n);
        predict_eterms(h, tr, h->v, h->w, h->w0, h->e, h->q);
        predict_eterms(h, te, h->v, h->w, h->w0, h->e_test, h->q_test);
        for (uint32_t c = 0; c < n; c++) h->e[c] = h->e[c] - tr->y[c];
    } else {                                                         /* vbos.h:54-62 */
        h->size_except_last = (uint32_t)ceil((double)n / h->num_batch);
        h->shuffle = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
        for (uint32_t i = 0; i < n; i++) h->shuffle[i] = i + 1;
    }
    return 0;
}

int orc_iterate(orc_t *h, orc_stats *out) {
    const split *tr = &h->sp[0], *te = &h->sp[1];
    uint32_t n = tr->n_cases, nt = te->n_cases;
    memset(out, 0, sizeof(*out));
    h->nan_inf = 0;
    if (h->method == ORC_VB_ONLINE) { int r = vbo_epoch(h, out); out->nan_inf_count = h->nan_inf; h->iter++; return r; }
    if (h->method == ORC_VB) {                                       /* vbs.h:75-258 */
        vb_update_all(h);
        predict_eterms(h, te, h->mu_v, h->mu_w, h->mu_0_dash, h->e_test, h->q_test);
        for (uint32_t c = 0; c < nt; c++) {
            double p = h->e_test[c];
            p = fmin(h->max_target, p); p = fmax(h->min_target, p);
            h->pred_this[c] = p;
        }
        double rmse_train = 0.0;
        for (uint32_t c = 0; c < n; c++) {                           /* vbs.h:153-162 */
            double p = h->e[c];
            p = fmin(h->max_target, p); p = fmax(h->min_target, p);
            rmse_train += p * p;
        }
        out->train_stat = sqrt(rmse_train / n);
        out->test_rmse = eval_rmse(h, h->pred_this, te->y, nt, 1.0);
        out->rmse_this = out->test_rmse;
        out->free_energy = h->last_free_energy; out->has_free_energy = h->have_fe;
    } else {                                                         /* mcmcs.h:96-303 */
        mcmc_draw_all(h);
        predict_eterms(h, tr, h->v, h->w, h->w0, h->e, h->q);
        predict_eterms(h, te, h->v, h->w, h->w0, h->e_test, h->q_test);
        if (h->task == 1) {                                          /* mcmcs.h:176-221, 262-275, 326-398 */
            for (uint32_t c = 0; c < nt; c++) {
                double p = orc_cdf_gaussian(h->e_test[c]);
                h->pred_this[c] = p;
                h->pred_sum_all[c] += p;
            }
            uint32_t acc_train = 0;
            for (uint32_t c = 0; c < n; c++) {
                double p = orc_cdf_gaussian(h->e[c]);
                if (((p >= 0.5) && (tr->y[c] > 0.0)) || ((p < 0.5) && (tr->y[c] < 0.0))) acc_train++;
                double sampled_target, mu = h->e[c];
                if (tr->y[c] >= 0.0) {
                    if (h->do_sample) sampled_target = orc_ran_left_tgaussian(0.0, mu, 1.0);
                    else {                                                       /* expected value of the truncated normal (3.141: the reference's pi) */
                        double phi_minus_mu = exp(-mu * mu / 2.0) / sqrt(3.141 * 2);
                        double Phi_minus_mu = orc_cdf_gaussian(-mu);
                        sampled_target = mu + phi_minus_mu / (1 - Phi_minus_mu);
                    }
                } else {
                    if (h->do_sample) sampled_target = orc_ran_right_tgaussian(0.0, mu, 1.0);
                    else {
                        double phi_minus_mu = exp(-mu * mu / 2.0) / sqrt(3.141 * 2);
                        double Phi_minus_mu = orc_cdf_gaussian(-mu);
                        sampled_target = mu - phi_minus_mu / Phi_minus_mu;
                    }
                }
                h->e[c] = h->e[c] - sampled_target;
            }
            out->train_stat = (double)acc_train / n;                 /* "Train=" */
            uint32_t a_this = 0, a_all = 0;
            double ll = 0.0;
            for (uint32_t c = 0; c < nt; c++) {                      /* _evaluate_class (this draw) and the accuracy part of _evaluate_class_map (running mean) */
                double p = h->pred_this[c], pa = h->pred_sum_all[c] * (1.0 / (h->iter + 1));
                if (((p >= 0.5) && (te->y[c] > 0.0)) || ((p < 0.5) && (te->y[c] < 0.0))) a_this++;
                if (((pa >= 0.5) && (te->y[c] > 0.0)) || ((pa < 0.5) && (te->y[c] < 0.0))) a_all++;
                double m = (te->y[c] + 1.0) * 0.5, pll = p;
                if (pll > 0.99) pll = 0.99;
                if (pll < 0.01) pll = 0.01;
                ll -= m * log10(pll) + (1 - m) * log10(1 - pll);
            }
            out->rmse_this = (double)a_this / nt;                    /* acc_mcmc_this */
            out->test_rmse = (double)a_all / nt;                     /* "Test=" (acc_mcmc_all) */
            out->free_energy = ll / nt;                              /* ll_mcmc_this, carried in the free-energy slot */
            out->alpha = h->alpha;
            out->nan_inf_count = h->nan_inf;
            h->iter++;
            return 0;
        }
        for (uint32_t c = 0; c < nt; c++) {                          /* mcmcs.h:154-163 */
            double p = h->e_test[c];
            h->pred_this[c] = p;
            p = fmin(h->max_target, p); p = fmax(h->min_target, p);
            h->pred_sum_all[c] += p;
        }
        double rmse_train = 0.0;
        for (uint32_t c = 0; c < n; c++) {                           /* mcmcs.h:166-174 */
            double p = h->e[c];
            p = fmin(h->max_target, p); p = fmax(h->min_target, p);
            double err = p - tr->y[c];
            rmse_train += err * err;
            h->e[c] = h->e[c] - tr->y[c];
        }
        out->train_stat = sqrt(rmse_train / n);
        out->rmse_this = eval_rmse(h, h->pred_this, te->y, nt, 1.0);
        out->test_rmse = eval_rmse(h, h->pred_sum_all, te->y, nt, 1.0 / (h->iter + 1));   /* mcmcs.h:241,245 */
    }
    out->alpha = h->alpha;
    out->nan_inf_count = h->nan_inf;
    h->iter++;
    return 0;
}

/* ------------------------------------------------------------------ state access */
int orc_get_state(orc_t *h, double *w0m, double *w0v, double *wm, double *wv, double *vm, double *vv) {
    size_t D = h->D, KD = (size_t)h->K * h->D;
    if (h->method == ORC_MCMC) {
        if (w0m) *w0m = h->w0;
        if (w0v) *w0v = 0.0;
        if (wm) memcpy(wm, h->w, D * sizeof(double));
        if (vm) memcpy(vm, h->v, KD * sizeof(double));
        if (wv) memset(wv, 0, D * sizeof(double));
        if (vv) memset(vv, 0, KD * sizeof(double));
    } else {
        if (w0m) *w0m = h->mu_0_dash;
        if (w0v) *w0v = h->sigma_0_dash;
        if (wm) memcpy(wm, h->mu_w, D * sizeof(double));
        if (wv) memcpy(wv, h->sg_w, D * sizeof(double));
        if (vm) memcpy(vm, h->mu_v, KD * sizeof(double));
        if (vv) memcpy(vv, h->sg_v, KD * sizeof(double));
    }
    return 0;
}
int orc_set_state(orc_t *h, double w0m, double w0v, const double *wm, const double *wv, const double *vm, const double *vv) {
    size_t D = h->D, KD = (size_t)h->K * h->D;
    if (h->method == ORC_MCMC) {
        h->w0 = w0m;
        if (wm) memcpy(h->w, wm, D * sizeof(double));
        if (vm) memcpy(h->v, vm, KD * sizeof(double));
    } else {
        h->mu_0_dash = w0m; h->sigma_0_dash = w0v;
        if (wm) memcpy(h->mu_w, wm, D * sizeof(double));
        if (wv) memcpy(h->sg_w, wv, D * sizeof(double));
        if (vm) memcpy(h->mu_v, vm, KD * sizeof(double));
        if (vv) memcpy(h->sg_v, vv, KD * sizeof(double));
    }
    return 0;
}
int orc_get_hyper(orc_t *h, double *alpha, double *sigma_0, double *sigma_w, double *sigma_v) {
    size_t G = h->G, GK = (size_t)h->G * h->K;
    if (alpha) *alpha = h->alpha;
    if (h->method == ORC_MCMC) {
        if (sigma_0) *sigma_0 = h->reg0;
        if (sigma_w) memcpy(sigma_w, h->w_lambda, G * sizeof(double));
        if (sigma_v) memcpy(sigma_v, h->v_lambda, GK * sizeof(double));
    } else {
        if (sigma_0) *sigma_0 = h->sigma_0;
        if (sigma_w) memcpy(sigma_w, h->sigma_w, G * sizeof(double));
        if (sigma_v) memcpy(sigma_v, h->sigma_v, GK * sizeof(double));
    }
    return 0;
}
int orc_get_train_cache(orc_t *h, double *e, double *t) {
    uint32_t n = h->sp[0].n_cases;
    if (e && h->e) memcpy(e, h->e, n * sizeof(double));
    if (t && h->t) memcpy(t, h->t, n * sizeof(double));
    return 0;
}
int orc_get_test_pred(orc_t *h, double *p) {
    if (h->method == ORC_MCMC) {                                     /* mcmc.h:355-379: mean of the draws (sampling) or the last prediction (als), clamped */
        for (uint32_t c = 0; c < h->sp[1].n_cases; c++) {
            double v = h->do_sample ? h->pred_sum_all[c] / (h->iter ? h->iter : 1) : h->pred_this[c];
            if (h->task == 1) { v = fmin(1.0, v); v = fmax(0.0, v); }     /* mcmc.h:372-374 */
            else { v = fmin(h->max_target, v); v = fmax(h->min_target, v); }
            p[c] = v;
        }
    } else memcpy(p, h->pred_this, h->sp[1].n_cases * sizeof(double));
    return 0;
}

/* ------------------------------------------------------------------ formats */
void orc_csr_free(orc_csr *m) { free(m->rowptr); free(m->col); free(m->val); free(m->y); memset(m, 0, sizeof(*m)); }

/* Data::load text branch (Data.h:173-283): two sscanf passes, blank and '#' lines skipped */
int orc_parse_text(const char *path, orc_csr *out) {
    memset(out, 0, sizeof(*out));
    out->min_target = +FLT_MAX; out->max_target = -FLT_MAX;
    int num_feature = 0, has_feature = 0;
    uint64_t num_values = 0; uint32_t num_rows = 0;
    char *line = NULL; size_t cap = 0; ssize_t len;
    for (int pass = 0; pass < 2; pass++) {
        FILE *f = fopen(path, "r");
        if (!f) return -1;
        uint32_t row_id = 0; uint64_t cache_id = 0;
        if (pass == 1) {
            if (has_feature) num_feature++;                          /* Data.h:220-222 */
            out->n_rows = num_rows; out->nnz = num_values; out->n_feat = (uint32_t)num_feature;
            out->rowptr = (uint64_t *)calloc((size_t)num_rows + 1, sizeof(uint64_t));
            out->col = (uint32_t *)malloc((num_values ? num_values : 1) * sizeof(uint32_t));
            out->val = (float *)malloc((num_values ? num_values : 1) * sizeof(float));
            out->y = (float *)malloc((num_rows ? num_rows : 1) * sizeof(float));
        }
        while ((len = getline(&line, &cap, f)) >= 0) {
            while (len > 0 && (line[len - 1] == '\n')) line[--len] = 0;
            const char *p = line;
            while ((*p == ' ') || (*p == 9)) p++;
            if ((*p == 0) || (*p == '#')) continue;
            float _value; int nchar, _feature;
            if (sscanf(p, "%f%n", &_value, &nchar) >= 1) {
                p += nchar;
                if (pass == 0) {
                    if (_value < out->min_target) out->min_target = _value;
                    if (_value > out->max_target) out->max_target = _value;
                    num_rows++;
                } else out->y[row_id] = _value;
                while (sscanf(p, "%d:%f%n", &_feature, &_value, &nchar) >= 2) {
                    p += nchar;
                    if (pass == 0) { if (_feature > num_feature) num_feature = _feature; has_feature = 1; num_values++; }
                    else { out->col[cache_id] = (uint32_t)_feature; out->val[cache_id] = _value; cache_id++; }
                }
                if (pass == 1) { row_id++; out->rowptr[row_id] = cache_id; }
                while ((*p != 0) && ((*p == ' ') || (*p == 9))) p++;
                if ((*p != 0) && (*p != '#')) { fclose(f); free(line); return -2; }   /* "cannot parse line" */
            } else { fclose(f); free(line); return -2; }
        }
        fclose(f);
    }
    free(line);
    return 0;
}

int orc_transpose(const orc_csr *in, uint32_t n_out_rows, orc_csr *out) {
    spm a = { in->n_rows, in->rowptr, in->col, in->val }, b;
    memset(out, 0, sizeof(*out));
    int r = spm_transpose(&a, n_out_rows, &b);
    if (r) return r;
    out->n_rows = n_out_rows; out->n_feat = in->n_rows; out->nnz = in->nnz;
    out->rowptr = b.ptr; out->col = b.id; out->val = b.val;
    return 0;
}

#pragma pack(push, 1)
typedef struct { uint32_t id, float_size; uint64_t num_values; uint32_t num_rows, num_cols; } file_header; /* fmatrix.h:46-52 */
#pragma pack(pop)

int orc_write_x(const char *path, const orc_csr *m, uint32_t num_cols) {
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    file_header fh = { 2, 4, m->nnz, m->n_rows, num_cols };
    fwrite(&fh, sizeof(fh), 1, f);
    for (uint32_t i = 0; i < m->n_rows; i++) {
        uint32_t size = (uint32_t)(m->rowptr[i + 1] - m->rowptr[i]);
        fwrite(&size, 4, 1, f);
        for (uint64_t p = m->rowptr[i]; p < m->rowptr[i + 1]; p++) { fwrite(&m->col[p], 4, 1, f); fwrite(&m->val[p], 4, 1, f); }
    }
    fclose(f);
    return 0;
}
int orc_write_y(const char *path, const float *y, uint32_t n) {
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    uint32_t hdr[3] = { 1, 4, n };
    fwrite(hdr, 4, 3, f); fwrite(y, 4, n, f); fclose(f);
    return 0;
}
int orc_read_x(const char *path, orc_csr *out) {
    memset(out, 0, sizeof(*out));
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    file_header fh;
    if (fread(&fh, sizeof(fh), 1, f) != 1 || fh.id != 2 || fh.float_size != 4) { fclose(f); return -2; }
    out->n_rows = fh.num_rows; out->n_feat = fh.num_cols; out->nnz = fh.num_values;
    out->rowptr = (uint64_t *)calloc((size_t)fh.num_rows + 1, sizeof(uint64_t));
    out->col = (uint32_t *)malloc((fh.num_values ? fh.num_values : 1) * 4);
    out->val = (float *)malloc((fh.num_values ? fh.num_values : 1) * 4);
    uint64_t w = 0;
    for (uint32_t i = 0; i < fh.num_rows; i++) {
        uint32_t size;
        if (fread(&size, 4, 1, f) != 1) { fclose(f); return -3; }
        for (uint32_t k = 0; k < size; k++) {
            if (fread(&out->col[w], 4, 1, f) != 1 || fread(&out->val[w], 4, 1, f) != 1) { fclose(f); return -3; }
            w++;
        }
        out->rowptr[i + 1] = w;
    }
    fclose(f);
    return 0;
}
int orc_read_y(const char *path, float **y, uint32_t *n) {
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    uint32_t hdr[3];
    if (fread(hdr, 4, 3, f) != 3 || hdr[0] != 1 || hdr[1] != 4) { fclose(f); return -2; }
    *n = hdr[2];
    *y = (float *)malloc((hdr[2] ? hdr[2] : 1) * 4);
    if (fread(*y, 4, hdr[2], f) != hdr[2]) { fclose(f); return -3; }
    fclose(f);
    return 0;
}
