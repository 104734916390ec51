/*
 * ORACLE -- test infrastructure, never shipped, never linked by the product.
 *
 * Plain-C restatement of the numeric hot path of QuadraticProgramNetworks.jl
 * v0.4.0 (reference at /root/reference).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * PARITY STATUS: "parity unpinned" against PATH itself.  The AVI arithmetic of
 * the reference lives in closed-source PATH 5.x behind PATHSolver.jl (compat
 * "1.7", /root/reference/Project.toml:28, no Manifest) and cannot run here; what
 * is restated is PATH's published pivotal method for affine problems (Cao &
 * Ferris 1996; Dirkse & Ferris 1995) in the form the package's own scratch
 * implementation gives it (/root/reference/src/deprecated/avi_scratch.jl:2-134),
 * with the SURVEY.md A9 defects repaired and a crash / extreme-point phase
 * added.  The restatement is pinned on KAT-0..8, on
 * check_avi_solution residuals, on scipy LP/QP optima and on uniqueness for
 * strongly monotone instances; see tests/test_oracle_*.py.
 *
 * All matrices are column-major fp64 (Julia native).  Every dot product is a
 * sequential fma chain in index order so the CUDA kernels can reproduce the
 * results bit for bit.
 *
 * Build: gcc -O2 -ffp-contract=off -pthread -shared -fPIC (see oracle/Makefile).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define QPO_SUCCESS 1
#define QPO_RAY_TERM 2
#define QPO_MAX_ITERS 3
#define QPO_FAILURE 4

#define PIV_TOL 1e-9
#define D_TOL 1e-10
#define TIE_TOL 1e-10

enum { AT_L = 0, AT_U = 1, FLOATING = 2, BASIC = 3 };

typedef struct {
    int n, ncol;            /* ncol = n + 1 */
    double *T;              /* n x ncol, column-major */
    double *beta, *nbval, *prow, *r;
    const double *l, *u;
    int *rowvar, *colvar;   /* variable ids: z_i = i, w_i = n+i, t = 2n */
    int *rowof, *colof;     /* inverse maps, -1 when absent */
    int8_t *zst;
    int pivots;
    /* frozen rows (see freeze): rows holding a FREE variable after phase 0 */
    int nfrozen;
    int *frozen, *colvar0;  /* row ids (n), column variables at the freeze (n + 1) */
    double *T0, *beta0, *nbval0;   /* nfrozen x (n+1) row-major, nfrozen, n + 1 */
} tab_t;

/* ---- avi_scratch.jl:17-50: normal-map start ------------------------------ */
static void tab_init(tab_t *t, int n, const double *M, const double *q, const double *l,
                     const double *u, const double *z0, double *work, int *iwork, int8_t *bwork) {
    t->n = n; t->ncol = n + 1;
    t->T = work;                       work += (size_t)n * (n + 1);
    t->beta = work;                    work += n;
    t->nbval = work;                   work += n + 1;
    t->prow = work;                    work += n + 1;
    t->r = work;                       work += n;
    double *zb = work;                 /* n, scratch */
    t->l = l; t->u = u;
    t->rowvar = iwork;                 iwork += n;
    t->colvar = iwork;                 iwork += n + 1;
    t->rowof = iwork;                  iwork += 2 * n + 1;
    t->colof = iwork;
    t->zst = bwork;
    t->pivots = 0;
    t->nfrozen = 0;
    t->frozen = t->colof + 2 * n + 1;
    t->colvar0 = t->frozen + n;
    t->nbval0 = zb + n;
    t->beta0 = t->nbval0 + n + 1;
    t->T0 = t->beta0 + n;
    for (int i = 0; i < n; ++i) zb[i] = fmin(fmax(z0[i], l[i]), u[i]);
    for (int i = 0; i < n; ++i) {
        double acc = 0.0;
        for (int j = 0; j < n; ++j) acc = fma(M[(size_t)j * n + i], zb[j], acc);
        double r = ((acc + q[i]) + z0[i]) - zb[i];
        t->r[i] = r;
        t->T[(size_t)n * n + i] = -r;
        t->beta[i] = zb[i] - z0[i];
        t->rowvar[i] = n + i;
        t->zst[i] = (z0[i] <= l[i]) ? AT_L : (z0[i] >= u[i]) ? AT_U : FLOATING;
    }
    for (int j = 0; j < n; ++j) {
        for (int i = 0; i < n; ++i) t->T[(size_t)j * n + i] = -M[(size_t)j * n + i];
        t->colvar[j] = j;
        t->nbval[j] = zb[j];
    }
    t->colvar[n] = 2 * n; t->nbval[n] = 0.0;
    for (int v = 0; v <= 2 * n; ++v) { t->rowof[v] = -1; t->colof[v] = -1; }
    for (int i = 0; i < n; ++i) t->rowof[n + i] = i;
    for (int j = 0; j <= n; ++j) t->colof[t->colvar[j]] = j;
}

static void var_bounds(const tab_t *t, int var, double *lo, double *up) {
    int n = t->n;
    if (var == 2 * n) { *lo = 0.0; *up = 1.0; return; }
    if (var < n) { *lo = t->l[var]; *up = t->u[var]; return; }
    int k = var - n;
    if (t->l[k] == t->u[k]) { *lo = -INFINITY; *up = INFINITY; return; }
    if (t->zst[k] == AT_L) { *lo = 0.0; *up = INFINITY; return; }
    if (t->zst[k] == AT_U) { *lo = -INFINITY; *up = 0.0; return; }
    *lo = 0.0; *up = 0.0;   /* z_k basic or floating: w_k is an artificial fixed at 0 */
}

static int artificial_row(const tab_t *t, int i) {
    int v = t->rowvar[i], n = t->n;
    if (v < n || v == 2 * n) return 0;
    int k = v - n;
    if (t->l[k] == t->u[k]) return 0;
    return t->zst[k] == FLOATING || t->zst[k] == BASIC;
}

/* ---- avi_scratch.jl:2-7: rank-1 pivot on the compact tableau ---------------- */
static void pivot(tab_t *t, int rho, int c) {
    int n = t->n, nc = t->ncol;
    double *T = t->T, *prow = t->prow;
    double p = T[(size_t)c * n + rho];
    for (int j = 0; j < nc; ++j) prow[j] = (j == c) ? (1.0 / p) : T[(size_t)j * n + rho] / p;
    for (int j = 0; j < nc; ++j) {
        double pj = prow[j];
        double *col = T + (size_t)j * n;
        const double *dcol = T + (size_t)c * n;
        if (j == c) continue;
        if (pj != 0.0)
            for (int i = 0; i < n; ++i) {
                double d = dcol[i];
                if (i != rho && d != 0.0) col[i] = fma(-d, pj, col[i]);
            }
        col[rho] = pj;
    }
    {   /* column c last: it still held the direction d */
        double *col = T + (size_t)c * n;
        double pc = prow[c];
        for (int i = 0; i < n; ++i) {
            double d = col[i];
            if (i == rho) continue;
            col[i] = (d != 0.0) ? fma(-d, pc, 0.0) : 0.0;
        }
        col[rho] = pc;
    }
    int ev = t->colvar[c], lv = t->rowvar[rho];
    t->rowvar[rho] = ev; t->colvar[c] = lv;
    t->rowof[ev] = rho; t->colof[ev] = -1;
    t->rowof[lv] = -1;  t->colof[lv] = c;
    double tmp = t->beta[rho]; t->beta[rho] = t->nbval[c]; t->nbval[c] = tmp;
    t->pivots++;
}

static int best_artificial_row(const tab_t *t, int c) {
    double best = 0.0; int brow = -1;
    for (int i = 0; i < t->n; ++i)
        if (artificial_row(t, i)) {
            double a = fabs(t->T[(size_t)c * t->n + i]);
            if (a > best) { best = a; brow = i; }
        }
    return best > PIV_TOL ? brow : -1;
}

/* ---- avi_scratch.jl:65-77: ratio test over the finite bounds of the basics -- */
static double ratio_test(const tab_t *t, int c, double sigma, int *rho_out, int *which_out) {
    int n = t->n;
    double theta = INFINITY;
    const double *col = t->T + (size_t)c * n;
    double *ratios = t->prow;                 /* prow is free between pivots (n+1 >= n) */
    for (int i = 0; i < n; ++i) {
        double d = sigma * col[i], lo, up, r = INFINITY;
        var_bounds(t, t->rowvar[i], &lo, &up);
        if (d > D_TOL && lo > -INFINITY) r = fmax((t->beta[i] - lo) / d, 0.0);
        else if (d < -D_TOL && up < INFINITY) r = fmax((up - t->beta[i]) / (-d), 0.0);
        ratios[i] = r;
        if (r < theta) theta = r;
    }
    *rho_out = -1; *which_out = 0;
    if (theta == INFINITY) return INFINITY;
    double cut = theta + TIE_TOL * (1.0 + theta);
    double best = -1.0; int rho = -1;
    for (int i = 0; i < n; ++i)
        if (ratios[i] <= cut) {
            if (t->rowvar[i] == 2 * n) { rho = i; break; }
            double a = fabs(col[i]);
            if (a > best) { best = a; rho = i; }
        }
    *rho_out = rho;
    *which_out = (sigma * col[rho] > 0.0) ? -1 : +1;
    return ratios[rho];
}

static void move(tab_t *t, int c, double sigma, double theta) {
    if (theta == 0.0) return;
    const double *col = t->T + (size_t)c * t->n;
    double st = sigma * theta;
    for (int i = 0; i < t->n; ++i)
        if (col[i] != 0.0) t->beta[i] = fma(-st, col[i], t->beta[i]);
    t->nbval[c] = fma(sigma, theta, t->nbval[c]);
}

static void leave_at(tab_t *t, int rho, int which) {
    double lo, up;
    var_bounds(t, t->rowvar[rho], &lo, &up);
    t->beta[rho] = which < 0 ? lo : up;
}

static int try_exchange(tab_t *t, int var) {
    int c = t->colof[var];
    if (c < 0) return 0;                       /* the variable is basic: nothing to bring in */
    int rho = best_artificial_row(t, c);
    if (rho < 0) return 0;
    pivot(t, rho, c);
    return 1;
}

static int is_free(const tab_t *t, int k) { return t->l[k] == -INFINITY && t->u[k] == INFINITY; }

/* largest |T[i,c]| over rows still holding the slack of a FREE variable */
static int best_free_row(const tab_t *t, int c) {
    double best = 0.0; int brow = -1, n = t->n;
    for (int i = 0; i < n; ++i) {
        int v = t->rowvar[i];
        if (v >= n && v < 2 * n && is_free(t, v - n)) {
            double a = fabs(t->T[(size_t)c * n + i]);
            if (a > best) { best = a; brow = i; }
        }
    }
    return best > PIV_TOL ? brow : -1;
}

/* T[:, t] = B^-1 r rebuilt from the slack columns (nonbasic w_k: column = -B^-1 e_k; w_k basic
 * in row rho: B^-1 e_k = -e_rho), sequential fma over k */
static void recompute_tcol(tab_t *t) {
    int n = t->n, tc = t->colof[2 * n];
    double *out = t->prow;
    for (int i = 0; i < n; ++i) {
        double acc = 0.0;
        for (int k = 0; k < n; ++k) {
            int ck = t->colof[n + k];
            double pik = ck >= 0 ? -t->T[(size_t)ck * n + i] : (t->rowof[n + k] == i ? -1.0 : 0.0);
            if (pik != 0.0) acc = fma(pik, t->r[k], acc);
        }
        out[i] = acc;
    }
    for (int i = 0; i < n; ++i) t->T[(size_t)tc * n + i] = out[i];
}

/* Rows that hold a FREE variable after phase 0 are frozen: a free basic never blocks a ratio test, never
 * leaves the basis and is no candidate row of the crash, so no later decision reads its row -- only the final z
 * does.  The row as it stands now stays a valid equation between the variables,
 *     x_B[i] = beta0[i] - sum_j T0[i][j] * (x(colvar0[j]) - nbval0[j]),
 * and the solve evaluates exactly that at the end (frozen_values) instead of carrying the row through every
 * later pivot.  Same statement as oracle/avi_pivot.py: freeze / solution. */
static void freeze(tab_t *t) {
    int n = t->n, nf = 0;
    for (int i = 0; i < n; ++i)
        if (t->rowvar[i] < n && t->l[t->rowvar[i]] == -INFINITY && t->u[t->rowvar[i]] == INFINITY) t->frozen[nf++] = i;
    t->nfrozen = nf;
    for (int j = 0; j <= n; ++j) { t->colvar0[j] = t->colvar[j]; t->nbval0[j] = t->nbval[j]; }
    for (int f = 0; f < nf; ++f) {
        int i = t->frozen[f];
        t->beta0[f] = t->beta[i];
        for (int j = 0; j <= n; ++j) t->T0[(size_t)f * (n + 1) + j] = t->T[(size_t)j * n + i];
    }
}

static double var_value(const tab_t *t, int v) {
    return t->rowof[v] >= 0 ? t->beta[t->rowof[v]] : t->nbval[t->colof[v]];
}

static void frozen_values(const tab_t *t, double *z) {
    int n = t->n;
    for (int f = 0; f < t->nfrozen; ++f) {
        double acc = t->beta0[f];
        for (int j = 0; j <= n; ++j)
            acc = fma(-t->T0[(size_t)f * (n + 1) + j], var_value(t, t->colvar0[j]) - t->nbval0[j], acc);
        z[t->rowvar[t->frozen[f]]] = acc;
    }
}

/* ---- crash: bring interior / free variables into the basis -------------------- */
static void crash(tab_t *t) {
    int n = t->n;
    /* phase 0: free variables against rows of free variables only -- independent of the start
     * point and of q, hence computed once per shared matrix by the GPU engine */
    int piv0 = t->pivots;
    for (int i = 0; i < n; ++i) {
        if (!is_free(t, i)) continue;
        int c = t->colof[i];
        int rho = best_free_row(t, c);
        if (rho >= 0) { pivot(t, rho, c); t->zst[i] = BASIC; }
    }
    if (t->pivots > piv0) recompute_tcol(t);
    freeze(t);
    /* phase 1: everything still floating, against any artificial row */
    for (int i = 0; i < n; ++i) {
        if (t->zst[i] != FLOATING) continue;
        int c = t->colof[i];
        int rho = best_artificial_row(t, c);
        if (rho >= 0) { pivot(t, rho, c); t->zst[i] = BASIC; continue; }
        /* dependent column: walk towards an extreme point (Cao-Ferris stage 2) */
        int rb[2], wb[2]; double th[2], own[2], step[2];
        for (int s = 0; s < 2; ++s) {
            double sigma = s == 0 ? 1.0 : -1.0;
            th[s] = ratio_test(t, c, sigma, &rb[s], &wb[s]);
            own[s] = s == 0 ? (t->u[i] - t->nbval[c]) : (t->nbval[c] - t->l[i]);
            step[s] = fmin(th[s], own[s]);
        }
        int s = step[0] <= step[1] ? 0 : 1;
        double sigma = s == 0 ? 1.0 : -1.0;
        if (step[s] == INFINITY) continue;          /* lineality direction: stays parked */
        if (own[s] <= th[s]) {
            move(t, c, sigma, own[s]);
            t->nbval[c] = s == 0 ? t->u[i] : t->l[i];
            t->zst[i] = s == 0 ? AT_U : AT_L;
            continue;
        }
        move(t, c, sigma, th[s]);
        leave_at(t, rb[s], wb[s]);
        int lv = t->rowvar[rb[s]];
        pivot(t, rb[s], c);
        t->zst[i] = BASIC;
        if (lv < n) {
            t->zst[lv] = wb[s] < 0 ? AT_L : AT_U;
            if (t->rowof[n + lv] < 0) try_exchange(t, n + lv);
        } else {
            int k = lv - n;
            if (try_exchange(t, k)) t->zst[k] = BASIC;
            else try_exchange(t, n + k);
        }
    }
}

static void repair(tab_t *t) {
    int n = t->n, progress = 1;
    while (progress) {
        progress = 0;
        int any = 0;
        for (int i = 0; i < n; ++i) any |= artificial_row(t, i);
        if (!any) return;
        for (int k = 0; k < n; ++k) {
            if ((t->zst[k] == AT_L || t->zst[k] == AT_U) && t->l[k] != t->u[k] &&
                t->rowof[k] < 0 && t->rowof[n + k] < 0) {
                if (try_exchange(t, n + k)) progress = 1;
                else if (try_exchange(t, k)) { t->zst[k] = BASIC; progress = 1; }
            }
        }
        for (int k = 0; k < n; ++k)
            if (t->zst[k] == FLOATING && t->rowof[k] < 0)
                if (try_exchange(t, k)) { t->zst[k] = BASIC; progress = 1; }
    }
}

/* ---- phase 2: avi_scratch.jl:59-132 complementary pivoting ------------------ */
static int lemke(tab_t *t, int max_pivots) {
    int n = t->n;
    int ent = 2 * n; double sigma = 1.0;
    for (;;) {
        if (t->pivots > max_pivots) return QPO_MAX_ITERS;
        int c = t->colof[ent], rb, wb;
        double th = ratio_test(t, c, sigma, &rb, &wb);
        double own = ent == 2 * n ? 1.0 - t->nbval[c] : ent < n ? (t->u[ent] - t->l[ent]) : INFINITY;
        if (own == INFINITY && th == INFINITY) return QPO_RAY_TERM;
        if (own <= th) {
            move(t, c, sigma, own);
            if (ent == 2 * n) { t->nbval[c] = 1.0; return QPO_SUCCESS; }
            t->nbval[c] = sigma > 0 ? t->u[ent] : t->l[ent];
            t->zst[ent] = sigma > 0 ? AT_U : AT_L;
            ent = n + ent; sigma = -sigma;
            continue;
        }
        move(t, c, sigma, th);
        leave_at(t, rb, wb);
        int lv = t->rowvar[rb];
        int was_art = artificial_row(t, rb);
        pivot(t, rb, c);
        if (ent < n) t->zst[ent] = BASIC;
        if (lv == 2 * n) return wb > 0 ? QPO_SUCCESS : QPO_RAY_TERM;
        if (lv < n) {
            t->zst[lv] = wb < 0 ? AT_L : AT_U;
            if (t->rowof[n + lv] >= 0) return QPO_FAILURE;
            ent = n + lv; sigma = wb < 0 ? 1.0 : -1.0;
        } else {
            int k = lv - n;
            if (was_art || t->rowof[k] >= 0 || !(t->zst[k] == AT_L || t->zst[k] == AT_U)) return QPO_FAILURE;
            ent = k; sigma = t->zst[k] == AT_L ? 1.0 : -1.0;
        }
    }
}

/* ---- avi.jl:148-156 ---------------------------------------------------------- */
int qpo_check_avi(int n, const double *M, const double *q, const double *l, const double *u,
                  const double *z, double tol, double *r_out) {
    int bad = 0;
    for (int i = 0; i < n; ++i) {
        double acc = 0.0;
        for (int j = 0; j < n; ++j) acc = fma(M[(size_t)j * n + i], z[j], acc);
        double r = acc + q[i];
        if (r_out) r_out[i] = r;
        if (r > tol && fabs(z[i] - l[i]) > tol) bad++;
        if (r < -tol && fabs(z[i] - u[i]) > tol) bad++;
        if (z[i] - l[i] < -tol) bad++;
        if (z[i] - u[i] > tol) bad++;
    }
    return bad;
}

size_t qpo_avi_work_doubles(int n) { return 2 * (size_t)n * (n + 1) + 7 * (size_t)n + 3; }
size_t qpo_avi_work_ints(int n) { return 8 * (size_t)n + 4; }

/* avi.jl:63-77 with q = N*w + o already formed.  basis: 1 at lower, 2 interior/basic,
 * 3 at upper, 4 fixed. */
int qpo_avi_solve(int n, const double *M, const double *q, const double *l, const double *u,
                  const double *z0, int max_pivots, double *z, int32_t *status, int32_t *pivots,
                  int8_t *basis) {
    if (max_pivots <= 0) max_pivots = 50 * n + 100;
    double *work = (double *)malloc(qpo_avi_work_doubles(n) * sizeof(double));
    int *iwork = (int *)malloc(qpo_avi_work_ints(n) * sizeof(int));
    int8_t *bwork = (int8_t *)malloc((size_t)n + 1);
    if (!work || !iwork || !bwork) { free(work); free(iwork); free(bwork); return -1; }
    tab_t t;
    tab_init(&t, n, M, q, l, u, z0, work, iwork, bwork);
    crash(&t);
    repair(&t);
    int st = lemke(&t, max_pivots);
    for (int i = 0; i < n; ++i) {
        z[i] = t.rowof[i] >= 0 ? t.beta[t.rowof[i]] : t.nbval[t.colof[i]];
        if (basis)
            basis[i] = (l[i] == u[i]) ? 4 : (t.rowof[i] >= 0 || t.zst[i] == FLOATING) ? 2
                       : (t.zst[i] == AT_L ? 1 : 3);
    }
    frozen_values(&t, z);
    if (st == QPO_SUCCESS && qpo_check_avi(n, M, q, l, u, z, 1e-6, NULL) > 0) st = QPO_FAILURE;
    *status = st;
    *pivots = t.pivots;
    free(work); free(iwork); free(bwork);
    return 0;
}

/* Batched form of the above: q, l, u, z0, z are n x B column-major; M is n x n when
 * M_is_shared, else n x n x B.  l/u shared when lu_is_shared. */
typedef struct {
    int n, batch, M_is_shared, lu_is_shared, max_pivots, tid, nthreads, err;
    const double *M, *q, *l, *u, *z0;
    double *z; int32_t *status, *pivots; int8_t *basis;
} batch_job_t;

static void *batch_worker(void *arg) {
    batch_job_t *j = (batch_job_t *)arg;
    int n = j->n;
    /* interleaved static schedule: instance b goes to thread b mod nthreads */
    for (int b = j->tid; b < j->batch; b += j->nthreads) {
        const double *Mb = j->M_is_shared ? j->M : j->M + (size_t)b * n * n;
        const double *lb = j->lu_is_shared ? j->l : j->l + (size_t)b * n;
        const double *ub = j->lu_is_shared ? j->u : j->u + (size_t)b * n;
        int rc = qpo_avi_solve(n, Mb, j->q + (size_t)b * n, lb, ub, j->z0 + (size_t)b * n,
                               j->max_pivots, j->z + (size_t)b * n, j->status + b, j->pivots + b,
                               j->basis ? j->basis + (size_t)b * n : NULL);
        if (rc) j->err = rc;
    }
    return NULL;
}

int qpo_avi_solve_batched(int n, int batch, const double *M, int M_is_shared, const double *q,
                          const double *l, const double *u, int lu_is_shared, const double *z0,
                          int max_pivots, double *z, int32_t *status, int32_t *pivots, int8_t *basis,
                          int threads) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    batch_job_t jobs[256];
    pthread_t th[256];
    for (int t = 0; t < threads; ++t) {
        batch_job_t j = {n, batch, M_is_shared, lu_is_shared, max_pivots, t, threads, 0,
                         M, q, l, u, z0, z, status, pivots, basis};
        jobs[t] = j;
    }
    for (int t = 1; t < threads; ++t) pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
    batch_worker(&jobs[0]);
    int err = jobs[0].err;
    for (int t = 1; t < threads; ++t) { pthread_join(th[t], NULL); if (jobs[t].err) err = jobs[t].err; }
    return err;
}

/* ---- helpers -------------------------------------------------------------- */
/* y = A x, A is m x k column-major, sequential fma per row */
static void matvec(int m, int k, const double *A, const double *x, double *y) {
    for (int i = 0; i < m; ++i) {
        double acc = 0.0;
        for (int j = 0; j < k; ++j) acc = fma(A[(size_t)j * m + i], x[j], acc);
        y[i] = acc;
    }
}

/* ---- avi.jl:79-128: presolve projection, GAVI -> AVI lift, solve ------------- */
/* GAVI blocks: M d1 x (d1+d2), N d1 x np, o d1, l1/u1 d1, A d2 x (d1+d2), B d2 x np, l2/u2 d2.
 * z0 has d1+d2 entries and is overwritten by the projected start when presolve != 0.
 * z_out: d1+d2; zfull_out (optional): d1+2*d2; basis (optional): d1+2*d2 codes. */
int qpo_gavi_solve(int d1, int d2, int np, const double *M, const double *N, const double *o,
                   const double *l1, const double *u1, const double *A, const double *B,
                   const double *l2, const double *u2, const double *w, double *z0, int presolve,
                   int max_pivots, double *z_out, double *zfull_out, int32_t *status, int32_t *pivots,
                   int8_t *basis) {
    int dz = d1 + d2, n = d1 + 2 * d2;
    int piv0 = 0;
    double *c = (double *)calloc((size_t)d2 + 1, sizeof(double));
    double *s0 = (double *)calloc((size_t)d2 + 1, sizeof(double));
    matvec(d2, np, B, w, c);
    matvec(d2, dz, A, z0, s0);
    for (int i = 0; i < d2; ++i) s0[i] += c[i];
    if (presolve && d2 > 0) {
        int feasible = 1;
        for (int i = 0; i < d2; ++i) if (!(l2[i] <= s0[i] && s0[i] <= u2[i])) feasible = 0;
        if (!feasible) {
            /* avi.jl:79-99:  min 0.5|z - z0|^2  s.t.  l2 - Bw <= A z <= u2 - Bw over the columns of
             * A that are not structurally zero; KKT system [I -A'; A 0] lifted with slacks. */
            int *cols = (int *)malloc(sizeof(int) * (size_t)dz);
            int k = 0;
            for (int j = 0; j < dz; ++j) {
                int nzc = 0;
                for (int i = 0; i < d2; ++i) if (A[(size_t)j * d2 + i] != 0.0) nzc = 1;
                if (nzc) cols[k++] = j;
            }
            int pn = k + 2 * d2;
            double *PM = (double *)calloc((size_t)pn * pn, sizeof(double));
            double *pq = (double *)calloc((size_t)pn, sizeof(double));
            double *pl = (double *)malloc(sizeof(double) * (size_t)pn);
            double *pu = (double *)malloc(sizeof(double) * (size_t)pn);
            double *ps = (double *)calloc((size_t)pn, sizeof(double));
            double *pz = (double *)calloc((size_t)pn, sizeof(double));
            for (int a = 0; a < k; ++a) {
                PM[(size_t)a * pn + a] = 1.0;
                for (int i = 0; i < d2; ++i) {
                    double v = A[(size_t)cols[a] * d2 + i];
                    PM[(size_t)(k + i) * pn + a] = -v;      /* -A' */
                    PM[(size_t)a * pn + (k + i)] = v;       /*  A  */
                }
            }
            for (int i = 0; i < d2; ++i) {
                PM[(size_t)(k + d2 + i) * pn + (k + i)] = -1.0;
                PM[(size_t)(k + i) * pn + (k + d2 + i)] = 1.0;
            }
            /* constant part of A z from the columns that do not move */
            for (int i = 0; i < d2; ++i) {
                double full = 0.0, part = 0.0;
                for (int j = 0; j < dz; ++j) full = fma(A[(size_t)j * d2 + i], z0[j], full);
                for (int a = 0; a < k; ++a) part = fma(A[(size_t)cols[a] * d2 + i], z0[cols[a]], part);
                pq[k + i] = (full - part) + c[i];
            }
            for (int a = 0; a < k; ++a) { pq[a] = -z0[cols[a]]; ps[a] = z0[cols[a]]; }
            for (int i = 0; i < k + d2; ++i) { pl[i] = -INFINITY; pu[i] = INFINITY; }
            for (int i = 0; i < d2; ++i) { pl[k + d2 + i] = l2[i]; pu[k + d2 + i] = u2[i]; ps[k + d2 + i] = s0[i]; }
            int32_t pst, ppiv;
            qpo_avi_solve(pn, PM, pq, pl, pu, ps, 0, pz, &pst, &ppiv, NULL);
            piv0 = ppiv;
            if (pst == QPO_SUCCESS) for (int a = 0; a < k; ++a) z0[cols[a]] = pz[a];
            free(cols); free(PM); free(pq); free(pl); free(pu); free(ps); free(pz);
            matvec(d2, dz, A, z0, s0);
            for (int i = 0; i < d2; ++i) s0[i] += c[i];
        }
    }
    /* avi.jl:113-128 */
    double *LM = (double *)calloc((size_t)n * n, sizeof(double));
    double *lq = (double *)calloc((size_t)n, sizeof(double));
    double *ll = (double *)malloc(sizeof(double) * (size_t)n);
    double *lu = (double *)malloc(sizeof(double) * (size_t)n);
    double *ls = (double *)malloc(sizeof(double) * (size_t)n);
    double *lz = (double *)malloc(sizeof(double) * (size_t)n);
    for (int j = 0; j < dz; ++j) {
        for (int i = 0; i < d1; ++i) LM[(size_t)j * n + i] = M[(size_t)j * d1 + i];
        for (int i = 0; i < d2; ++i) LM[(size_t)j * n + d1 + i] = A[(size_t)j * d2 + i];
    }
    for (int i = 0; i < d2; ++i) {
        LM[(size_t)(dz + i) * n + d1 + i] = -1.0;
        LM[(size_t)(d1 + i) * n + dz + i] = 1.0;
    }
    matvec(d1, np, N, w, lq);
    for (int i = 0; i < d1; ++i) lq[i] += o[i];
    for (int i = 0; i < d2; ++i) lq[d1 + i] = c[i];
    for (int i = 0; i < d1; ++i) { ll[i] = l1[i]; lu[i] = u1[i]; }
    for (int i = 0; i < d2; ++i) { ll[d1 + i] = -INFINITY; lu[d1 + i] = INFINITY; ll[dz + i] = l2[i]; lu[dz + i] = u2[i]; }
    for (int i = 0; i < dz; ++i) ls[i] = z0[i];
    for (int i = 0; i < d2; ++i) ls[dz + i] = s0[i];
    int32_t st, piv;
    qpo_avi_solve(n, LM, lq, ll, lu, ls, max_pivots, lz, &st, &piv, basis);
    for (int i = 0; i < dz; ++i) z_out[i] = lz[i];
    if (zfull_out) for (int i = 0; i < n; ++i) zfull_out[i] = lz[i];
    *status = st; *pivots = piv + piv0;
    free(LM); free(lq); free(ll); free(lu); free(ls); free(lz); free(c); free(s0);
    return 0;
}

/* ---- avi_solutions.jl:511-562 (no requests): 4-bit masks ---------------------- */
static int approx_eq(double a, double b, double atol) { return a == b || fabs(a - b) <= atol; }

void qpo_comp_indices_block(int n, const double *l, const double *u, const double *r, const double *z,
                            double tol, int8_t *mask) {
    for (int i = 0; i < n; ++i) {
        int eq = approx_eq(l[i], u[i], tol), m = 0;
        if (approx_eq(z[i], l[i], tol) && r[i] >= -tol && !eq) m |= 1;
        if (l[i] - tol <= z[i] && z[i] <= u[i] + tol && approx_eq(r[i], 0.0, tol) && !eq) m |= 2;
        if (approx_eq(z[i], u[i], tol) && r[i] <= tol && !eq) m |= 4;
        if (m == 0) m = eq ? 8 : 0;          /* 0 = the reference's @assert would fire */
        mask[i] = (int8_t)m;
    }
}

/* avi_solutions.jl:587-612 */
void qpo_comp_indices(int d1, int d2, int np, const double *M, const double *N, const double *o,
                      const double *l1, const double *u1, const double *A, const double *B,
                      const double *l2, const double *u2, const double *z, const double *w, double tol,
                      int8_t *mask) {
    int dz = d1 + d2;
    double *r1 = (double *)malloc(sizeof(double) * (size_t)(d1 + 1));
    double *s2 = (double *)malloc(sizeof(double) * (size_t)(d2 + 1));
    double *tmp = (double *)malloc(sizeof(double) * (size_t)(dz + 1));
    matvec(d1, dz, M, z, r1);
    matvec(d1, np, N, w, tmp);
    for (int i = 0; i < d1; ++i) r1[i] = (r1[i] + tmp[i]) + o[i];
    matvec(d2, dz, A, z, s2);
    matvec(d2, np, B, w, tmp);
    for (int i = 0; i < d2; ++i) s2[i] += tmp[i];
    qpo_comp_indices_block(d1, l1, u1, r1, z, tol, mask);
    qpo_comp_indices_block(d2, l2, u2, z + d1, s2, tol, mask + d1);
    free(r1); free(s2); free(tmp);
}

/* ---- sets.jl:820-853: membership ---------------------------------------------- */
/* A is m x d column-major; rl/ru: 1 = strict '<', 0 = '<='.  Returns 1 iff x in poly. */
int qpo_halfspace_in(int m, int d, const double *A, const double *l, const double *u,
                     const uint8_t *rl, const uint8_t *ru, const double *x, double tol) {
    for (int i = 0; i < m; ++i) {
        double ax = 0.0;
        for (int j = 0; j < d; ++j) ax = fma(A[(size_t)j * m + i], x[j], ax);
        int lo = (rl && rl[i]) ? (l[i] - tol < ax) : (l[i] - tol <= ax);
        int up = (ru && ru[i]) ? (ax - tol < u[i]) : (ax - tol <= u[i]);
        if (!(lo && up)) return 0;
    }
    return 1;
}

/* ---- qp_processing.jl:57-149: verify_solution ----------------------------------- */
/* Least squares  Abar(nd x k) lam ~ rhs  by Householder QR with column pivoting.  All
 * reductions run down one column at a time, so a thread-per-column GPU version reproduces
 * the bits.  Dependent columns get a zero multiplier (basic solution). */
static void lstsq_basic(int nd, int k, double *Ab, double *b, double *lam, int *perm, double *v) {
    int rank = 0, steps = nd < k ? nd : k;
    for (int j = 0; j < k; ++j) perm[j] = j;
    for (int c = 0; c < steps; ++c) {
        double best = -1.0; int jb = -1;
        for (int j = c; j < k; ++j) {
            double s = 0.0;
            for (int i = c; i < nd; ++i) s = fma(Ab[(size_t)j * nd + i], Ab[(size_t)j * nd + i], s);
            if (s > best) { best = s; jb = j; }
        }
        double nrm = sqrt(best);
        if (nrm <= 1e-10) break;
        if (jb != c) {
            for (int i = 0; i < nd; ++i) { double t = Ab[(size_t)c * nd + i]; Ab[(size_t)c * nd + i] = Ab[(size_t)jb * nd + i]; Ab[(size_t)jb * nd + i] = t; }
            int tp = perm[c]; perm[c] = perm[jb]; perm[jb] = tp;
        }
        double alpha = Ab[(size_t)c * nd + c] > 0.0 ? -nrm : nrm;
        double vn = 0.0;
        for (int i = c; i < nd; ++i) { v[i] = Ab[(size_t)c * nd + i]; }
        v[c] -= alpha;
        for (int i = c; i < nd; ++i) vn = fma(v[i], v[i], vn);
        if (vn > 0.0) {
            for (int j = c; j < k; ++j) {
                double s = 0.0;
                for (int i = c; i < nd; ++i) s = fma(v[i], Ab[(size_t)j * nd + i], s);
                s = (2.0 * s) / vn;
                for (int i = c; i < nd; ++i) Ab[(size_t)j * nd + i] = fma(-s, v[i], Ab[(size_t)j * nd + i]);
            }
            double s = 0.0;
            for (int i = c; i < nd; ++i) s = fma(v[i], b[i], s);
            s = (2.0 * s) / vn;
            for (int i = c; i < nd; ++i) b[i] = fma(-s, v[i], b[i]);
        }
        rank++;
    }
    for (int j = 0; j < k; ++j) lam[j] = 0.0;
    for (int i = rank - 1; i >= 0; --i) {
        double acc = b[i];
        for (int j = i + 1; j < rank; ++j) acc = fma(-Ab[(size_t)j * nd + i], v[j], acc);   /* v reused: permuted lam */
        v[i] = acc / Ab[(size_t)i * nd + i];
    }
    for (int i = 0; i < rank; ++i) lam[perm[i]] = v[i];
}

/* Qd: nd x nv (rows dec of Q), qd: nd, A: m x nv stacked constraint rows, dec: nd indices
 * (0-based) into x.  lam_out: m.  Returns: 1 solution, 0 not a solution; *how: 0 infeasible,
 * 1 unconstrained, 2 least squares accepted, 3 sign-constrained fallback accepted,
 * 4 fallback rejected, 5 fallback solve failed.  fallback_pivots (optional): pivots the
 * sign-constrained fallback spent. */
int qpo_verify_solution(int nd, int nv, int m, const double *Qd, const double *qd, const double *A,
                        const double *l, const double *u, const int32_t *dec, const double *x,
                        double tol, double *lam_out, int32_t *how, int8_t *active, int32_t *fallback_pivots) {
    double *qt = (double *)malloc(sizeof(double) * (size_t)(nd + 1));
    double *ax = (double *)malloc(sizeof(double) * (size_t)(m + 1));
    matvec(nd, nv, Qd, x, qt);
    for (int i = 0; i < nd; ++i) qt[i] += qd[i];
    matvec(m, nv, A, x, ax);
    int ret = 0;
    if (fallback_pivots) *fallback_pivots = 0;
    for (int i = 0; i < m; ++i) { lam_out[i] = 0.0; if (active) active[i] = 0; }
    for (int i = 0; i < m; ++i)
        if (!((l[i] - 1e-3 <= ax[i]) && (ax[i] - 1e-3 <= u[i]))) { *how = 0; goto done; }
    double nq = 0.0;
    for (int i = 0; i < nd; ++i) nq = fma(qt[i], qt[i], nq);
    if (m == 0) { *how = 1; ret = sqrt(nq) <= tol; goto done; }
    {
        int *idx = (int *)malloc(sizeof(int) * (size_t)(m + 1));
        int8_t *kind = (int8_t *)malloc((size_t)m + 1);    /* 1 pos, 2 neg, 3 both, 0 inactive */
        int np_ = 0, nn = 0, nb = 0;
        for (int i = 0; i < m; ++i) {
            int pos = ax[i] < l[i] + 1e-2, neg = ax[i] > u[i] - 1e-2;
            kind[i] = (pos && neg) ? 3 : pos ? 1 : neg ? 2 : 0;
            if (active) active[i] = kind[i];
        }
        int k = 0;
        for (int i = 0; i < m; ++i) if (kind[i] == 1) { idx[k++] = i; np_++; }
        for (int i = 0; i < m; ++i) if (kind[i] == 2) { idx[k++] = i; nn++; }
        for (int i = 0; i < m; ++i) if (kind[i] == 3) { idx[k++] = i; nb++; }
        double *Ab = (double *)malloc(sizeof(double) * ((size_t)nd * k + 1));
        double *Ab0 = (double *)malloc(sizeof(double) * ((size_t)nd * k + 1));
        double *b = (double *)malloc(sizeof(double) * (size_t)(nd + 1));
        double *lam = (double *)malloc(sizeof(double) * (size_t)(k + 1));
        double *v = (double *)malloc(sizeof(double) * (size_t)(nd + k + 1));
        int *perm = (int *)malloc(sizeof(int) * (size_t)(k + 1));
        for (int t = 0; t < k; ++t) {
            double sgn = (t >= np_ && t < np_ + nn) ? -1.0 : 1.0;
            for (int i = 0; i < nd; ++i) Ab0[(size_t)t * nd + i] = Ab[(size_t)t * nd + i] = sgn * A[(size_t)dec[i] * m + idx[t]];
        }
        for (int i = 0; i < nd; ++i) b[i] = qt[i];
        lstsq_basic(nd, k, Ab, b, lam, perm, v);
        int ok = 1;
        for (int t = 0; t < np_ + nn; ++t) if (!(lam[t] > -tol)) ok = 0;
        double res = 0.0;
        for (int i = 0; i < nd; ++i) {
            double acc = 0.0;
            for (int t = 0; t < k; ++t) acc = fma(Ab0[(size_t)t * nd + i], lam[t], acc);
            double e = acc - qt[i];
            res = fma(e, e, res);
        }
        if (!(sqrt(res) <= tol)) ok = 0;
        if (ok) {
            for (int t = 0; t < k; ++t) lam_out[idx[t]] = (t >= np_ && t < np_ + nn) ? -lam[t] : lam[t];
            *how = 2; ret = 1;
        } else {
            /* qp_processing.jl:129-146: min |Ad' lam - qt|^2 with sign bounds, as the box AVI
             * (Ad Ad') lam - Ad qt  complementary to  lb <= lam <= ub */
            double *G = (double *)malloc(sizeof(double) * ((size_t)m * m + 1));
            double *h = (double *)malloc(sizeof(double) * (size_t)(m + 1));
            double *lb = (double *)malloc(sizeof(double) * (size_t)(m + 1));
            double *ub = (double *)malloc(sizeof(double) * (size_t)(m + 1));
            double *z0 = (double *)calloc((size_t)m + 1, sizeof(double));
            double *lam2 = (double *)malloc(sizeof(double) * (size_t)(m + 1));
            for (int i = 0; i < m; ++i) {
                for (int j = 0; j < m; ++j) {
                    double acc = 0.0;
                    for (int t = 0; t < nd; ++t) acc = fma(A[(size_t)dec[t] * m + i], A[(size_t)dec[t] * m + j], acc);
                    G[(size_t)j * m + i] = acc;
                }
                double acc = 0.0;
                for (int t = 0; t < nd; ++t) acc = fma(A[(size_t)dec[t] * m + i], qt[t], acc);
                h[i] = -acc;
                lb[i] = (kind[i] == 2 || kind[i] == 3) ? -INFINITY : 0.0;
                ub[i] = (kind[i] == 1 || kind[i] == 3) ? INFINITY : 0.0;
            }
            int32_t st, pv;
            qpo_avi_solve(m, G, h, lb, ub, z0, 0, lam2, &st, &pv, NULL);
            if (fallback_pivots) *fallback_pivots = pv;
            if (st != QPO_SUCCESS) { *how = 5; ret = 0; }
            else {
                double res2 = 0.0;
                for (int t = 0; t < nd; ++t) {
                    double acc = 0.0;
                    for (int i = 0; i < m; ++i) acc = fma(A[(size_t)dec[t] * m + i], lam2[i], acc);
                    double e = acc - qt[t];
                    res2 = fma(e, e, res2);
                }
                for (int i = 0; i < m; ++i) lam_out[i] = lam2[i];
                if (sqrt(res2) <= 1e-4) { *how = 3; ret = 1; } else { *how = 4; ret = 0; }
            }
            free(G); free(h); free(lb); free(ub); free(z0); free(lam2);
        }
        free(idx); free(kind); free(Ab); free(Ab0); free(b); free(lam); free(v); free(perm);
    }
done:
    free(qt); free(ax);
    return ret;
}

/* ---- algorithm.jl:13-118 for a level without children --------------------------------- */
typedef struct { int32_t nd, nv, m; const double *Qd, *qd, *A, *l, *u; const int32_t *dec; } qpo_node;
typedef struct { int32_t d1, d2, np; const double *M, *N, *o, *l1, *u1, *A, *B, *l2, *u2; } qpo_gavi;

/* proj: nv x nproj (column k = projection vector k).  lam_out (optional): sum(m_p).
 * pivots counts the AVI pivots of solve_qep, of its presolve and of verify_solution's
 * fallback.  Returns 0. */
int qpo_level_solve(int nv, int nplayers, const qpo_node *players, const qpo_gavi *g, const int32_t *dec,
                    int nd_level, const int32_t *par, int max_iters, int nproj, const double *proj,
                    const double *x_init, double *x_out, uint8_t *solved_out, int32_t *iters_out,
                    int32_t *pivots_out, double *lam_out) {
    int dz = g->d1 + g->d2, lam_total = 0, max_m = 1;
    for (int p = 0; p < nplayers; ++p) { lam_total += players[p].m; if (players[p].m > max_m) max_m = players[p].m; }
    double *x = (double *)malloc(sizeof(double) * (size_t)(nv + 1));
    double *xn = (double *)malloc(sizeof(double) * (size_t)(nv + 1));
    double *w = (double *)malloc(sizeof(double) * (size_t)(g->np + 1));
    double *z0 = (double *)malloc(sizeof(double) * (size_t)(dz + 1));
    double *z = (double *)malloc(sizeof(double) * (size_t)(dz + 1));
    double *lam = (double *)malloc(sizeof(double) * (size_t)(max_m + 1));
    double *hist = (double *)malloc(sizeof(double) * ((size_t)max_iters * (nproj > 0 ? nproj : 1) + 1));
    double *pv = (double *)malloc(sizeof(double) * (size_t)(nproj + 1));
    memcpy(x, x_init, sizeof(double) * (size_t)nv);
    int solved = 0, piv = 0, it = 0, nhist = 0;
    for (it = 1; it <= max_iters; ++it) {
        if (nproj > 0) {
            int cyc = 0;
            for (int k = 0; k < nproj; ++k) {
                double acc = 0.0;
                for (int j = 0; j < nv; ++j) acc = fma(x[j], proj[(size_t)k * nv + j], acc);
                pv[k] = acc;
            }
            for (int h = 0; h < nhist && !cyc; ++h) {
                double dd = 0.0, na = 0.0, nb = 0.0;
                for (int k = 0; k < nproj; ++k) {
                    double e = pv[k] - hist[(size_t)h * nproj + k];
                    dd = fma(e, e, dd); na = fma(pv[k], pv[k], na);
                    nb = fma(hist[(size_t)h * nproj + k], hist[(size_t)h * nproj + k], nb);
                }
                if (sqrt(dd) <= 1.4901161193847656e-8 * fmax(sqrt(na), sqrt(nb))) cyc = 1;   /* isapprox, rtol = sqrt(eps) */
            }
            if (cyc) break;
            memcpy(hist + (size_t)nhist * nproj, pv, sizeof(double) * (size_t)nproj);
            nhist++;
        }
        int all_sol = 1, off = 0;
        for (int p = 0; p < nplayers; ++p) {
            const qpo_node *nd = &players[p];
            int32_t how, fpiv;
            int sol = qpo_verify_solution(nd->nd, nd->nv, nd->m, nd->Qd, nd->qd, nd->A, nd->l, nd->u, nd->dec, x, 1e-4,
                                          lam, &how, NULL, &fpiv);
            piv += fpiv;
            if (lam_out) for (int r = 0; r < nd->m; ++r) lam_out[off + r] = sol ? lam[r] : 0.0;
            off += nd->m;
            if (!sol) all_sol = 0;
        }
        if (all_sol) { solved = 1; break; }
        for (int j = 0; j < g->np; ++j) w[j] = x[par[j]];
        for (int j = 0; j < dz; ++j) z0[j] = j < nd_level ? x[dec[j]] : 0.0;
        int32_t st, gp;
        qpo_gavi_solve(g->d1, g->d2, g->np, g->M, g->N, g->o, g->l1, g->u1, g->A, g->B, g->l2, g->u2, w, z0, 1, 0, z,
                       NULL, &st, &gp, NULL);
        piv += gp;
        if (st != QPO_SUCCESS) break;
        memcpy(xn, x, sizeof(double) * (size_t)nv);
        for (int j = 0; j < nd_level; ++j) xn[dec[j]] = z[j];
        double dn = 0.0;
        for (int j = 0; j < nv; ++j) { double e = xn[j] - x[j]; dn = fma(e, e, dn); }
        if (sqrt(dn) < 1e-4) break;
        memcpy(x, xn, sizeof(double) * (size_t)nv);
    }
    if (it > max_iters) it = max_iters;
    memcpy(x_out, x, sizeof(double) * (size_t)nv);
    *solved_out = (uint8_t)solved; *iters_out = it; *pivots_out = piv;
    free(x); free(xn); free(w); free(z0); free(z); free(lam); free(hist); free(pv);
    return 0;
}

typedef struct {
    int nv, nplayers, nd_level, max_iters, nproj, batch, tid, nthreads, lam_total;
    const qpo_node *players; const qpo_gavi *g; const int32_t *dec, *par; const double *proj, *x_init;
    double *x_out; uint8_t *solved; int32_t *iters, *pivots; double *lam;
} level_job_t;

static void *level_worker(void *arg) {
    level_job_t *j = (level_job_t *)arg;
    for (int b = j->tid; b < j->batch; b += j->nthreads)
        qpo_level_solve(j->nv, j->nplayers, j->players, j->g, j->dec, j->nd_level, j->par, j->max_iters, j->nproj, j->proj,
                        j->x_init + (size_t)b * j->nv, j->x_out + (size_t)b * j->nv, j->solved + b, j->iters + b,
                        j->pivots + b, j->lam ? j->lam + (size_t)b * j->lam_total : NULL);
    return NULL;
}

int qpo_level_solve_batched(int nv, int nplayers, const qpo_node *players, const qpo_gavi *g, const int32_t *dec,
                            int nd_level, const int32_t *par, int max_iters, int nproj, const double *proj, int batch,
                            const double *x_init, double *x_out, uint8_t *solved_out, int32_t *iters_out,
                            int32_t *pivots_out, double *lam_out, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    int lam_total = 0;
    for (int p = 0; p < nplayers; ++p) lam_total += players[p].m;
    level_job_t jobs[256];
    pthread_t th[256];
    for (int t = 0; t < threads; ++t) {
        level_job_t j = {nv, nplayers, nd_level, max_iters, nproj, batch, t, threads, lam_total, players, g, dec, par, proj,
                         x_init, x_out, solved_out, iters_out, pivots_out, lam_out};
        jobs[t] = j;
    }
    for (int t = 1; t < threads; ++t) pthread_create(&th[t], NULL, level_worker, &jobs[t]);
    level_worker(&jobs[0]);
    for (int t = 1; t < threads; ++t) pthread_join(th[t], NULL);
    return 0;
}
