// Batched bounded-variable complementary pivoting for affine variational
// inequalities on sm_100a: one CTA per instance, the fp64 compact tableau in shared
// memory, one thread per tableau row.
//
// Replaces PATHSolver.solve_mcp as called by solve_avi
// (/root/reference/src/avi.jl:63-77); the pivotal method is the one sketched in
// /root/reference/src/deprecated/avi_scratch.jl:2-134 (normal-map start :17-50, ratio
// test :65-77, rank-1 pivot :2-7, complementary entering rule :105-131) completed with a
// crash / extreme-point phase.  The arithmetic is, operation for operation, the one
// oracle/qpn_oracle.c performs (explicit fma, sequential dot products), so discrete
// outputs (basis, status, pivot counts) and z agree bit for bit.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <math_constants.h>

namespace qpn {

constexpr double PIV_TOL = 1e-9;   // smallest |pivot| accepted in a crash exchange
constexpr double D_TOL = 1e-10;    // |direction entry| treated as zero in ratio tests
constexpr double TIE_TOL = 1e-10;  // ratios within this (relative) of the minimum tie

enum : int8_t { AT_L = 0, AT_U = 1, FLOATING = 2, BASIC = 3 };
enum : int { ST_SUCCESS = 1, ST_RAY_TERM = 2, ST_MAX_ITERS = 3, ST_FAILURE = 4 };

#define QPN_INF CUDART_INF

// Debug build only (-DQPN_TRACE): barrier wrappers that verify that every warp of the CTA
// arrived at the same source line with all 32 lanes active; mismatches are recorded in a
// host-mapped buffer (4 ints per CTA), which stays readable after a device fault.
#ifdef QPN_TRACE
__device__ int* qpn_trace_ptr = nullptr;
__device__ inline void qpn_dbg_record(int slot, int value) {
    if (qpn_trace_ptr) ((volatile int*)qpn_trace_ptr)[blockIdx.x * 16 + slot] = value;
}
__device__ inline void qpn_dbg_presync(int line) {
    __shared__ int dbg_line[32];
    const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const unsigned am = __activemask();
    if (am != 0xffffffffu) qpn_dbg_record(2, line * 100 + __popc(am));       // partial warp at a barrier
    if ((threadIdx.x & 31) == 0) dbg_line[w] = line;
    __syncthreads();
    int other = 0;
    for (int k = 0; k < nw; ++k) if (dbg_line[k] != line) other = dbg_line[k];
    if (other && (threadIdx.x & 31) == 0 && qpn_trace_ptr && ((volatile int*)qpn_trace_ptr)[blockIdx.x * 16 + 3] == -1) {
        qpn_dbg_record(3, w); qpn_dbg_record(0, line); qpn_dbg_record(1, other);      // first mismatch only
    }
    __syncthreads();
}
#define QPN_SYNC() do { qpn::qpn_dbg_presync(__LINE__); } while (0)
#define QPN_SYNC_OR(x) (qpn::qpn_dbg_presync(__LINE__), __syncthreads_or(x))
#define QPN_SITE(id, extra) do {} while (0)
#else
#define QPN_SYNC() __syncthreads()
#define QPN_SYNC_OR(x) __syncthreads_or(x)
#define QPN_SITE(id, extra) do {} while (0)
#endif

// Shared-memory workspace of one instance.  Sizes in elements for a problem of size n.
struct Tab {
    int n;          // rows
    int ld;         // leading dimension of T (>= n)
    double* T;      // ld x (n+1), column-major: column j = d(basic)/d(nonbasic j), negated
    double* beta;   // n   values of the basic variables
    double* nbval;  // n+1 values of the nonbasic variables
    double* prow;   // n+1 scaled pivot row
    double* l;      // n
    double* u;      // n
    int* rowvar;    // n     variable basic in row i      (z_i = i, w_i = n+i, t = 2n)
    int* colvar;    // n+1   variable nonbasic in column j
    int* rowof;     // 2n+1  row of a variable or -1
    int* colof;     // 2n+1  column of a variable or -1
    int8_t* zst;    // n     AT_L / AT_U / FLOATING / BASIC
    double* red_d;  // 32 + 2
    int* red_i;     // 32 + 2
    int pivots;     // uniform across the CTA
};

__host__ __device__ inline size_t tab_smem_bytes(int n, int ld) {
    size_t d = (size_t)ld * (n + 1) + (size_t)n /*beta*/ + 2 * (size_t)(n + 1) /*nbval,prow*/ +
               2 * (size_t)n /*l,u*/ + 34 /*red_d*/;
    size_t i = (size_t)n + (n + 1) + 2 * (size_t)(2 * n + 1) + 34;
    size_t b = (size_t)n;
    return d * 8 + ((i * 4 + 7) / 8) * 8 + ((b + 7) / 8) * 8;
}

__device__ inline void tab_carve(Tab& t, int n, int ld, unsigned char* smem) {
    t.n = n; t.ld = ld;
    double* d = reinterpret_cast<double*>(smem);
    t.T = d;      d += (size_t)ld * (n + 1);
    t.beta = d;   d += n;
    t.nbval = d;  d += n + 1;
    t.prow = d;   d += n + 1;
    t.l = d;      d += n;
    t.u = d;      d += n;
    t.red_d = d;  d += 34;
    int* ip = reinterpret_cast<int*>(d);
    t.rowvar = ip; ip += n;
    t.colvar = ip; ip += n + 1;
    t.rowof = ip;  ip += 2 * n + 1;
    t.colof = ip;  ip += 2 * n + 1;
    t.red_i = ip;  ip += 34;
    size_t ib = (size_t)(ip - reinterpret_cast<int*>(d));
    t.zst = reinterpret_cast<int8_t*>(reinterpret_cast<unsigned char*>(d) + ((ib * 4 + 7) / 8) * 8);
    t.pivots = 0;
}

// ---- block-wide arg-best: larger value wins, ties -> lower index --------------------
// Every thread receives the winner.  idx < 0 marks "no candidate" and loses to anything.
__device__ inline void block_argmax(const Tab& t, double& v, int& idx) {
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double v2 = __shfl_xor_sync(full, v, o);
        int i2 = __shfl_xor_sync(full, idx, o);
        bool take = (i2 >= 0) && (idx < 0 || v2 > v || (v2 == v && i2 < idx));
        if (take) { v = v2; idx = i2; }
    }
    const int nw = blockDim.x >> 5;
    if (nw > 1) {
        const int w = threadIdx.x >> 5;
        QPN_SYNC();                        // red_* free for reuse
        if ((threadIdx.x & 31) == 0) { t.red_d[w] = v; t.red_i[w] = idx; }
        QPN_SYNC();
        v = t.red_d[0]; idx = t.red_i[0];
        for (int k = 1; k < nw; ++k) {
            double v2 = t.red_d[k]; int i2 = t.red_i[k];
            bool take = (i2 >= 0) && (idx < 0 || v2 > v || (v2 == v && i2 < idx));
            if (take) { v = v2; idx = i2; }
        }
    }
}

__device__ inline double block_min(const Tab& t, double v) {
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(full, v, o));
    const int nw = blockDim.x >> 5;
    if (nw > 1) {
        const int w = threadIdx.x >> 5;
        QPN_SYNC();
        if ((threadIdx.x & 31) == 0) t.red_d[w] = v;
        QPN_SYNC();
        v = t.red_d[0];
        for (int k = 1; k < nw; ++k) v = fmin(v, t.red_d[k]);
    }
    return v;
}

__device__ inline void var_bounds(const Tab& t, int var, double& lo, double& up) {
    const int n = t.n;
    if (var == 2 * n) { lo = 0.0; up = 1.0; return; }
    if (var < n) { lo = t.l[var]; up = t.u[var]; return; }
    const int k = var - n;
    if (t.l[k] == t.u[k]) { lo = -QPN_INF; up = QPN_INF; return; }
    const int8_t s = t.zst[k];
    if (s == AT_L) { lo = 0.0; up = QPN_INF; return; }
    if (s == AT_U) { lo = -QPN_INF; up = 0.0; return; }
    lo = 0.0; up = 0.0;  // z_k basic or floating: w_k is an artificial fixed at 0
}

__device__ inline bool artificial_row(const Tab& t, int i) {
    const int v = t.rowvar[i], n = t.n;
    if (v < n || v == 2 * n) return false;
    const int k = v - n;
    if (t.l[k] == t.u[k]) return false;
    const int8_t s = t.zst[k];
    return s == FLOATING || s == BASIC;
}

// ---- start of the normal-map path (avi_scratch.jl:17-50) -----------------------------
// Expects T[:, 0:n] = -M, t.l, t.u filled, q and z0 readable (shared or global), and
// zb = t.prow as scratch.  Ends with a barrier.
__device__ inline void tab_start(Tab& t, const double* q, const double* z0) {
    const int n = t.n, i = threadIdx.x;
    double* zb = t.prow;
    if (i < n) zb[i] = fmin(fmax(z0[i], t.l[i]), t.u[i]);
    QPN_SYNC();
    if (i < n) {
        double acc = 0.0;
        for (int j = 0; j < n; ++j) {
            const double mij = -t.T[(size_t)j * t.ld + i];
            if (mij != 0.0) acc = fma(mij, zb[j], acc);
        }
        const double zi = z0[i], zbi = zb[i];
        const double r = ((acc + q[i]) + zi) - zbi;
        t.T[(size_t)n * t.ld + i] = -r;
        t.beta[i] = zbi - zi;
        t.rowvar[i] = n + i;
        t.zst[i] = (zi <= t.l[i]) ? AT_L : (zi >= t.u[i]) ? AT_U : FLOATING;
        t.colvar[i] = i;
        t.nbval[i] = zbi;
        t.rowof[i] = -1; t.colof[i] = i;
        t.rowof[n + i] = i; t.colof[n + i] = -1;
    }
    if (i == 0) {
        t.colvar[n] = 2 * n; t.nbval[n] = 0.0;
        t.rowof[2 * n] = -1; t.colof[2 * n] = n;
    }
    t.pivots = 0;
    QPN_SYNC();
}

// ---- rank-1 pivot on the compact tableau (avi_scratch.jl:2-7) --------------------------
__device__ inline void pivot(Tab& t, int rho, int c) {
    const int n = t.n, nc = n + 1, ld = t.ld, i = threadIdx.x;
    double* T = t.T;
    const double p = T[(size_t)c * ld + rho];
    for (int j = i; j < nc; j += blockDim.x)
        t.prow[j] = (j == c) ? (1.0 / p) : T[(size_t)j * ld + rho] / p;
    const double d = (i < n) ? T[(size_t)c * ld + i] : 0.0;
    QPN_SYNC();
    if (i < n) {
        if (i == rho) {
            for (int j = 0; j < nc; ++j) T[(size_t)j * ld + i] = t.prow[j];
        } else if (d != 0.0) {
            const double nd = -d;
#pragma unroll 4
            for (int j = 0; j < nc; ++j) {
                const double pj = t.prow[j];
                if (j == c) T[(size_t)j * ld + i] = fma(nd, pj, 0.0);
                else if (pj != 0.0) T[(size_t)j * ld + i] = fma(nd, pj, T[(size_t)j * ld + i]);
            }
        }
    }
    if (i == 0) {
        const int ev = t.colvar[c], lv = t.rowvar[rho];
        t.rowvar[rho] = ev; t.colvar[c] = lv;
        t.rowof[ev] = rho; t.colof[ev] = -1;
        t.rowof[lv] = -1;  t.colof[lv] = c;
        const double tmp = t.beta[rho]; t.beta[rho] = t.nbval[c]; t.nbval[c] = tmp;
    }
    t.pivots++;
    QPN_SYNC();
}

__device__ inline int best_artificial_row(const Tab& t, int c) {
    const int i = threadIdx.x;
    double a = 0.0; int idx = -1;
    if (i < t.n && artificial_row(t, i)) {
        a = fabs(t.T[(size_t)c * t.ld + i]);
        if (a > 0.0) idx = i;
    }
    block_argmax(t, a, idx);
    return (idx >= 0 && a > PIV_TOL) ? idx : -1;
}

// ---- ratio test over the finite bounds of the basics (avi_scratch.jl:65-77) ------------
// Returns the step of the blocking row (INF if none); all threads get the same answer.
__device__ inline double ratio_test(const Tab& t, int c, double sigma, int& rho, int& which) {
    const int n = t.n, i = threadIdx.x;
    double r = QPN_INF, a = 0.0;
    bool is_t = false;
    if (i < n) {
        const double ci = t.T[(size_t)c * t.ld + i];
        const double d = sigma * ci;
        const int v = t.rowvar[i];
        double lo, up;
        var_bounds(t, v, lo, up);
        if (d > D_TOL && lo > -QPN_INF) r = fmax((t.beta[i] - lo) / d, 0.0);
        else if (d < -D_TOL && up < QPN_INF) r = fmax((up - t.beta[i]) / (-d), 0.0);
        a = fabs(ci);
        is_t = (v == 2 * n);
    }
    const double theta = block_min(t, r);
    rho = -1; which = 0;
    if (theta == QPN_INF) { QPN_SYNC(); return QPN_INF; }   // every exit ends with a barrier
    const double cut = theta + TIE_TOL * (1.0 + theta);
    // among ties: t first (so the path terminates), then largest |d|, then lowest row
    double key = -1.0; int idx = -1;
    if (i < n && r <= cut) { key = is_t ? QPN_INF : a; idx = i; }
    block_argmax(t, key, idx);
    rho = idx;
    // the winner publishes its own ratio
    if (blockDim.x > 32) QPN_SYNC();
    if (i == rho) t.red_d[32] = r;
    QPN_SYNC();
    const double th = t.red_d[32];
    which = (sigma * t.T[(size_t)c * t.ld + rho] > 0.0) ? -1 : +1;
    return th;
}

__device__ inline void move(Tab& t, int c, double sigma, double theta) {
    if (theta == 0.0) return;
    const int i = threadIdx.x;
    if (i < t.n) {
        const double ci = t.T[(size_t)c * t.ld + i];
        if (ci != 0.0) t.beta[i] = fma(-(sigma * theta), ci, t.beta[i]);
    }
    if (i == 0) t.nbval[c] = fma(sigma, theta, t.nbval[c]);
    QPN_SYNC();
}

__device__ inline void leave_at(Tab& t, int rho, int which) {
    if (threadIdx.x == 0) {
        double lo, up;
        var_bounds(t, t.rowvar[rho], lo, up);
        t.beta[rho] = which < 0 ? lo : up;
    }
    QPN_SYNC();
}

__device__ inline void set_zst(Tab& t, int k, int8_t s) {
    QPN_SYNC();
    if (threadIdx.x == 0) t.zst[k] = s;
    QPN_SYNC();
}

__device__ inline bool try_exchange(Tab& t, int var) {
    const int c = t.colof[var];
    const int rho = best_artificial_row(t, c);
    if (rho < 0) return false;
    pivot(t, rho, c);
    return true;
}

// ---- phase 1: bring interior / free variables into the basis ----------------------------
__device__ inline void crash(Tab& t) {
    const int n = t.n;
    for (int i = 0; i < n; ++i) {
        if (t.zst[i] != FLOATING) continue;
        const int c = t.colof[i];
        const int rho = best_artificial_row(t, c);
        if (rho >= 0) { pivot(t, rho, c); set_zst(t, i, BASIC); continue; }
        // dependent column: walk towards an extreme point (Cao-Ferris stage 2)
        int rb[2], wb[2]; double th[2], own[2], step[2];
        own[0] = t.u[i] - t.nbval[c];        // read before the ratio tests (see lemke)
        own[1] = t.nbval[c] - t.l[i];
        for (int s = 0; s < 2; ++s) {
            const double sigma = s == 0 ? 1.0 : -1.0;
            th[s] = ratio_test(t, c, sigma, rb[s], wb[s]);
            step[s] = fmin(th[s], own[s]);
        }
        const int s = step[0] <= step[1] ? 0 : 1;
        const double sigma = s == 0 ? 1.0 : -1.0;
        if (step[s] == QPN_INF) continue;          // lineality direction: stays parked
        if (own[s] <= th[s]) {
            move(t, c, sigma, own[s]);
            QPN_SYNC();
            if (threadIdx.x == 0) { t.nbval[c] = s == 0 ? t.u[i] : t.l[i]; t.zst[i] = s == 0 ? AT_U : AT_L; }
            QPN_SYNC();
            continue;
        }
        move(t, c, sigma, th[s]);
        leave_at(t, rb[s], wb[s]);
        const int lv = t.rowvar[rb[s]];
        pivot(t, rb[s], c);
        set_zst(t, i, BASIC);
        if (lv < n) {
            set_zst(t, lv, wb[s] < 0 ? AT_L : AT_U);
            if (t.rowof[n + lv] < 0) try_exchange(t, n + lv);
        } else {
            const int k = lv - n;
            if (try_exchange(t, k)) set_zst(t, k, BASIC);
            else try_exchange(t, n + k);
        }
    }
}

__device__ inline void repair(Tab& t) {
    const int n = t.n;
    bool progress = true;
    while (progress) {
        progress = false;
        int any = 0;
        if (threadIdx.x < n) any = artificial_row(t, threadIdx.x) ? 1 : 0;
        any = QPN_SYNC_OR(any);
        if (!any) return;
        for (int k = 0; k < n; ++k) {
            const int8_t s = t.zst[k];
            if ((s == AT_L || s == AT_U) && t.l[k] != t.u[k] && t.rowof[k] < 0 && t.rowof[n + k] < 0) {
                if (try_exchange(t, n + k)) progress = true;
                else if (try_exchange(t, k)) { set_zst(t, k, BASIC); progress = true; }
            }
        }
        for (int k = 0; k < n; ++k)
            if (t.zst[k] == FLOATING && t.rowof[k] < 0)
                if (try_exchange(t, k)) { set_zst(t, k, BASIC); progress = true; }
    }
}

// ---- phase 2: complementary pivoting (avi_scratch.jl:59-132) ------------------------------
__device__ inline int lemke(Tab& t, int max_pivots) {
    const int n = t.n;
    int ent = 2 * n; double sigma = 1.0;
    for (;;) {
        if (t.pivots > max_pivots) return ST_MAX_ITERS;
        const int c = t.colof[ent];
        int rb, wb;
        // Read everything the branch below depends on BEFORE the ratio test: its barriers then
        // separate these reads from the writes in move() (a read after it would race with them).
        const double own = ent == 2 * n ? 1.0 - t.nbval[c] : ent < n ? (t.u[ent] - t.l[ent]) : QPN_INF;
        const double th = ratio_test(t, c, sigma, rb, wb);
        if (own == QPN_INF && th == QPN_INF) return ST_RAY_TERM;
        if (own <= th) {
            move(t, c, sigma, own);
            QPN_SYNC();
            if (ent == 2 * n) { if (threadIdx.x == 0) t.nbval[c] = 1.0; QPN_SYNC(); return ST_SUCCESS; }
            if (threadIdx.x == 0) { t.nbval[c] = sigma > 0 ? t.u[ent] : t.l[ent]; t.zst[ent] = sigma > 0 ? AT_U : AT_L; }
            QPN_SYNC();
            ent = n + ent; sigma = -sigma;
            continue;
        }
        move(t, c, sigma, th);
        leave_at(t, rb, wb);
        const int lv = t.rowvar[rb];
        const bool was_art = artificial_row(t, rb);
        pivot(t, rb, c);
        if (ent < n) set_zst(t, ent, BASIC);
        if (lv == 2 * n) return wb > 0 ? ST_SUCCESS : ST_RAY_TERM;
        if (lv < n) {
            set_zst(t, lv, wb < 0 ? AT_L : AT_U);
            if (t.rowof[n + lv] >= 0) return ST_FAILURE;
            ent = n + lv; sigma = wb < 0 ? 1.0 : -1.0;
        } else {
            const int k = lv - n;
            const int8_t s = t.zst[k];
            if (was_art || t.rowof[k] >= 0 || !(s == AT_L || s == AT_U)) return ST_FAILURE;
            ent = k; sigma = s == AT_L ? 1.0 : -1.0;
        }
    }
}

// Runs crash + repair + path following on a started tableau.  On return thread i < n holds
// z_i in *zi and its basis code in *code (1 lower, 2 basic/interior, 3 upper, 4 fixed).
__device__ inline int avi_pivot_run(Tab& t, int max_pivots, double* zi, int8_t* code) {
    crash(t);
    repair(t);
    const int st = lemke(t, max_pivots);
    const int i = threadIdx.x;
    if (i < t.n) {
        const int r = t.rowof[i];
        *zi = r >= 0 ? t.beta[r] : t.nbval[t.colof[i]];
        const int8_t s = t.zst[i];
        *code = (t.l[i] == t.u[i]) ? 4 : (r >= 0 || s == FLOATING) ? 2 : (s == AT_L ? 1 : 3);
    }
    return st;
}

// check_avi_solution (avi.jl:148-156) contribution of one index.
__device__ inline int check_avi_index(double r, double z, double l, double u, double tol) {
    int bad = 0;
    if (r > tol && fabs(z - l) > tol) bad++;
    if (r < -tol && fabs(z - u) > tol) bad++;
    if (z - l < -tol) bad++;
    if (z - u > tol) bad++;
    return bad;
}

}  // namespace qpn
