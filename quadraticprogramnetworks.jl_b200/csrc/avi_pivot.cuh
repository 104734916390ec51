// Batched bounded-variable complementary pivoting for affine variational
// inequalities on sm_100a: one CTA per instance, the fp64 compact tableau in shared
// memory, one thread per tableau row.
//
// Replaces PATHSolver.solve_mcp as called by solve_avi
// (/root/reference/src/avi.jl:63-77); the pivotal method is the one sketched in
// /root/reference/src/deprecated/avi_scratch.jl:2-134 (normal-map start :17-50, ratio
// test :65-77, rank-1 pivot :2-7, complementary entering rule :105-131) completed with a
// crash / extreme-point phase.  The arithmetic is, operation for operation, the one the
// CPU oracle's specification prescribes (explicit fma, sequential dot products), so
// discrete outputs (basis, status, pivot counts) and z agree bit for bit with it.
//
// Layout: T is ROW-major with row stride ldr = 2*odd doubles.  Thread i owns row i and
// walks it with 128-bit shared-memory accesses (two columns per LDS/STS, conflict-free
// across the warp); the scaled pivot row is broadcast from shared memory.  Columns of
// slack variables that can never re-enter (w_k of a free z_k) are swapped out of the live
// range [0, ncol) as soon as they appear, so later pivots touch fewer columns.
//
// Synchronisation rule: a value that steers control flow is read from
// shared memory BEFORE the last barrier that precedes any write to it.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <math_constants.h>

namespace qpn {

constexpr double PIV_TOL = 1e-9;   // smallest |pivot| accepted in a crash exchange
constexpr double D_TOL = 1e-10;    // |direction entry| treated as zero in ratio tests
constexpr double TIE_TOL = 1e-10;  // ratios within this (relative) of the minimum tie

enum : int8_t { AT_L = 0, AT_U = 1, FLOATING = 2, BASIC = 3, FROZEN = 4 };
// FROZEN: a free variable that phase 0 of the crash made basic.  It never blocks a ratio test, never leaves the
// basis and its row is no candidate of the crash, so no decision reads that row: the row stays as it was after
// phase 0 (it is left out of every later pivot sweep) and the variable's final value is evaluated from it once,
// at the end (oracle/avi_pivot.py: freeze / solution).  Everywhere else FROZEN reads like BASIC.
enum : int { ST_SUCCESS = 1, ST_RAY_TERM = 2, ST_MAX_ITERS = 3, ST_FAILURE = 4 };

#define QPN_INF CUDART_INF

// Debug build only (-DQPN_TRACE): barrier wrappers that verify that every warp of the CTA
// arrived at the same source line with all 32 lanes active; the first mismatch of a CTA is
// recorded in a host-mapped buffer (16 ints per CTA), readable after a device fault.
#ifdef QPN_TRACE
__device__ int* qpn_trace_ptr = nullptr;
__device__ __forceinline__ void qpn_dbg_record(int slot, int value) {
    if (qpn_trace_ptr) ((volatile int*)qpn_trace_ptr)[blockIdx.x * 16 + slot] = value;
}
__device__ __forceinline__ void qpn_dbg_presync(int line) {
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
#else
#define QPN_SYNC() __syncthreads()
#define QPN_SYNC_OR(x) __syncthreads_or(x)
#endif

// Row stride for `cols` columns: even (rows stay 16-byte aligned) with ldr/2 odd, so that the
// 32 lanes of a warp, each reading 16 bytes of its own row, fall into distinct bank groups.
__host__ __device__ __forceinline__ int row_stride(int cols) {
    int e = (cols + 1) & ~1;
    if ((e >> 1) % 2 == 0) e += 2;
    return e;
}

// All dynamic shared memory of every kernel in this library.
extern __shared__ __align__(16) unsigned char qpn_smem[];

// Bytes of a workspace with room for `nmax` rows, a tableau buffer of `tdoubles` doubles and
// rows of up to `ldrmax` doubles.  One workspace serves several solves of different shapes
// (tab_shape) inside one kernel.
// fz = 1: room for the freeze record (column variables / nonbasic values after phase 0); the global-memory
// engine keeps that record in its slot instead (fz = 0).
__host__ __device__ __forceinline__ size_t tab_smem_bytes_ex(int nmax, size_t tdoubles, int ldrmax, int fz = 1) {
    tdoubles = (tdoubles + 1) & ~(size_t)1;
    size_t d = tdoubles + (2 + (size_t)fz) * (size_t)ldrmax + (4 + (size_t)fz) * (size_t)nmax + 36;
    size_t i = (size_t)nmax + (1 + (size_t)fz) * ldrmax + 2 * (size_t)(2 * nmax + 1) + 36;
    size_t b = (size_t)nmax;
    return d * 8 + ((i * 4 + 15) / 16) * 16 + ((b + 15) / 16) * 16;
}
__host__ __device__ __forceinline__ size_t tab_smem_bytes(int n, int cap) {
    return tab_smem_bytes_ex(n, (size_t)n * row_stride(cap), row_stride(cap));
}

// Shared-memory workspace of one instance.  Only a base offset and three sizes are stored; the
// arrays are addressed by offset arithmetic (keeping a dozen pointers alive cost the level kernel
// a local-memory stack frame and a load per access).
struct Tab {
    int base;       // byte offset of the workspace in qpn_smem
    int nmax, ldrmax, td;   // capacity: rows, longest row, tableau doubles (even)
    int n;          // rows of the current solve
    int ldr;        // row stride of T in doubles for the current solve
    int ncol;       // live columns [0, ncol); uniform across the CTA
    int pivots;     // uniform across the CTA
    int fz;         // 1: the freeze record lives in this workspace (0: elsewhere, global-memory engine)
    int ncol0;      // live columns when the rows of free basics were frozen (uniform)
    int own_frozen; // PER THREAD: the row this thread owns (row threadIdx.x) is frozen -- fixed from the freeze to the end of a solve

    __device__ __forceinline__ int dl() const { return (2 + fz) * ldrmax; }               // doubles of the per-column vectors
    __device__ __forceinline__ int il() const { return (1 + fz) * ldrmax; }               // ints of the per-column vectors
    __device__ __forceinline__ double* dbl(int off) const { return reinterpret_cast<double*>(qpn_smem + base) + off; }
    __device__ __forceinline__ double* T() const { return dbl(0); }                       // n x ldr, row-major, negated
    __device__ __forceinline__ double* prow() const { return dbl(td); }                   // ldr  scaled pivot row
    __device__ __forceinline__ double* nbval() const { return dbl(td + ldrmax); }         // ldr  nonbasic values
    __device__ __forceinline__ double* nbval0() const { return dbl(td + 2 * ldrmax); }    // ldr  (fz) nonbasic values at the freeze
    __device__ __forceinline__ double* beta() const { return dbl(td + dl()); }            // n    basic values
    __device__ __forceinline__ double* l() const { return dbl(td + dl() + nmax); }
    __device__ __forceinline__ double* u() const { return dbl(td + dl() + 2 * nmax); }
    __device__ __forceinline__ double* rr() const { return dbl(td + dl() + 3 * nmax); }   // residual r at the start
    __device__ __forceinline__ double* birv() const { return dbl(td + dl() + 4 * nmax); }  // n  (fz) (B^-1 r)_i of the frozen rows a plan start did not copy
    __device__ __forceinline__ double* red_d() const { return dbl(td + dl() + (4 + fz) * nmax); }  // 36: [0,32) per warp, [32] ratio test, [33..35] below
    // Where the frozen rows are (written by thread 0 at the start of a solve, before a barrier).  A plan that exported the
    // swept rows first (PlanDesc::nact) leaves T() with rows 0 .. nact-1 only; rows nact .. n-1 are then read from the
    // plan at the end (same row stride).  Null: every row is in T().
    __device__ __forceinline__ const double** frozen_src() const { return reinterpret_cast<const double**>(red_d() + 33); }
    __device__ __forceinline__ int* frozen_hdr() const { return reinterpret_cast<int*>(red_d() + 34); }   // [0] nact, [1] column of t in the plan
    __device__ __forceinline__ int* ints() const { return reinterpret_cast<int*>(dbl(td + dl() + (4 + fz) * nmax + 36)); }
    __device__ __forceinline__ int* rowvar() const { return ints(); }                     // n     z_i = i, w_i = n+i, t = 2n
    __device__ __forceinline__ int* colvar() const { return ints() + nmax; }              // ldr
    __device__ __forceinline__ int* colvar0() const { return ints() + nmax + ldrmax; }    // ldr  (fz) column variables at the freeze
    __device__ __forceinline__ int* rowof() const { return ints() + nmax + il(); }        // 2n+1  row of a variable or -1
    __device__ __forceinline__ int* colof() const { return ints() + nmax + il() + 2 * nmax + 1; }   // 2n+1 (-1: basic or dead)
    __device__ __forceinline__ int* red_i() const { return ints() + nmax + il() + 2 * (2 * nmax + 1); }   // 36
    __device__ __forceinline__ int8_t* zst() const {                                      // n  AT_L / AT_U / FLOATING / BASIC / FROZEN
        const int ib = nmax + il() + 2 * (2 * nmax + 1) + 36;
        return reinterpret_cast<int8_t*>(reinterpret_cast<unsigned char*>(ints()) + ((ib * 4 + 15) / 16) * 16);
    }
    // entry (i, j) of a frozen row: in place (the frozen rows of the shared-memory tableau are simply never swept)
    __device__ __forceinline__ double frozen_entry(int i, int j) const { return T()[(size_t)i * ldr + j]; }
};

__device__ __forceinline__ void tab_carve_ex(Tab& t, int nmax, size_t tdoubles, int ldrmax, int base_off, int fz = 1) {
    t.base = base_off; t.nmax = nmax; t.ldrmax = ldrmax; t.td = (int)((tdoubles + 1) & ~(size_t)1);
    t.n = nmax; t.ldr = ldrmax; t.ncol = 0; t.pivots = 0; t.fz = fz; t.ncol0 = 0; t.own_frozen = 0;
}
__device__ __forceinline__ void tab_carve(Tab& t, int n, int cap, int base_off) {
    tab_carve_ex(t, n, (size_t)n * row_stride(cap), row_stride(cap), base_off);
}
// Shape of the next solve inside a carved workspace.
__device__ __forceinline__ void tab_shape(Tab& t, int n, int cap) { t.n = n; t.ldr = row_stride(cap); }

// ---- block-wide reductions on non-negative doubles (+inf allowed, no NaN) ---------------------
// The IEEE bit pattern of a non-negative double orders like an unsigned integer, so a 64-bit
// max / min is two 32-bit REDUX instructions instead of a five-round shuffle butterfly.
__device__ __forceinline__ void warp_max_bits(unsigned& hi, unsigned& lo) {
    const unsigned full = 0xffffffffu;
    const unsigned mh = __reduce_max_sync(full, hi);
    const unsigned ml = __reduce_max_sync(full, hi == mh ? lo : 0u);
    hi = mh; lo = ml;
}
__device__ __forceinline__ void warp_min_bits(unsigned& hi, unsigned& lo) {
    const unsigned full = 0xffffffffu;
    const unsigned mh = __reduce_min_sync(full, hi);
    const unsigned ml = __reduce_min_sync(full, hi == mh ? lo : 0xffffffffu);
    hi = mh; lo = ml;
}

// arg-max over the rows: larger value wins, ties -> lower row.  `valid` lanes carry v >= 0.
// Every thread receives (v, idx); idx = -1 when no lane is valid.
__device__ __forceinline__ void block_argmax(const Tab& t, bool valid, double& v, int& idx) {
    const unsigned full = 0xffffffffu;
    unsigned hi = valid ? (unsigned)__double2hiint(v) : 0u, lo = valid ? (unsigned)__double2loint(v) : 0u;
    const unsigned mine_hi = hi, mine_lo = lo;
    warp_max_bits(hi, lo);
    const unsigned ball = __ballot_sync(full, valid && mine_hi == hi && mine_lo == lo);
    int widx = ball ? (int)(threadIdx.x & ~31u) + (__ffs(ball) - 1) : -1;
    double wv = __hiloint2double((int)hi, (int)lo);
    const int nw = blockDim.x >> 5;
    if (nw > 1) {
        const int w = threadIdx.x >> 5;
        QPN_SYNC();                        // red_* free for reuse
        if ((threadIdx.x & 31) == 0) { t.red_d()[w] = wv; t.red_i()[w] = widx; }
        QPN_SYNC();
        wv = t.red_d()[0]; widx = t.red_i()[0];
        #pragma unroll 1
        for (int k = 1; k < nw; ++k) {
            const double v2 = t.red_d()[k]; const int i2 = t.red_i()[k];
            if (i2 >= 0 && (widx < 0 || v2 > wv)) { wv = v2; widx = i2; }      // warps are in row order: ties keep the lower row
        }
    }
    v = wv; idx = widx;
}

// Same reduction when the candidates are not the tableau rows: every thread offers (v >= 0, idx);
// larger v wins, ties -> lower idx.
__device__ __forceinline__ void block_argmax_idx(const Tab& t, bool valid, double& v, int& idx) {
    const unsigned full = 0xffffffffu;
    unsigned hi = valid ? (unsigned)__double2hiint(v) : 0u, lo = valid ? (unsigned)__double2loint(v) : 0u;
    const unsigned mine_hi = hi, mine_lo = lo;
    warp_max_bits(hi, lo);
    const bool win = valid && mine_hi == hi && mine_lo == lo;
    int widx = (int)__reduce_min_sync(full, win ? (unsigned)idx : 0x7fffffffu);
    if (widx == 0x7fffffff) widx = -1;
    double wv = __hiloint2double((int)hi, (int)lo);
    const int nw = blockDim.x >> 5;
    if (nw > 1) {
        const int w = threadIdx.x >> 5;
        QPN_SYNC();
        if ((threadIdx.x & 31) == 0) { t.red_d()[w] = wv; t.red_i()[w] = widx; }
        QPN_SYNC();
        wv = t.red_d()[0]; widx = t.red_i()[0];
        #pragma unroll 1
        for (int k = 1; k < nw; ++k) {
            const double v2 = t.red_d()[k]; const int i2 = t.red_i()[k];
            if (i2 >= 0 && (widx < 0 || v2 > wv || (v2 == wv && i2 < widx))) { wv = v2; widx = i2; }
        }
    }
    v = wv; idx = widx;
}

__device__ __forceinline__ double block_min(const Tab& t, double v) {   // v >= 0 or +inf
    unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    warp_min_bits(hi, lo);
    double m = __hiloint2double((int)hi, (int)lo);
    const int nw = blockDim.x >> 5;
    if (nw > 1) {
        const int w = threadIdx.x >> 5;
        QPN_SYNC();
        if ((threadIdx.x & 31) == 0) t.red_d()[w] = m;
        QPN_SYNC();
        m = t.red_d()[0];
        #pragma unroll 1
        for (int k = 1; k < nw; ++k) m = fmin(m, t.red_d()[k]);
    }
    return m;
}

template <class TT>
__device__ __forceinline__ bool is_free_var(const TT& t, int k) { return t.l()[k] == -QPN_INF && t.u()[k] == QPN_INF; }

// Does row i hold a frozen free variable?  (reads rowvar[i] and zst: call it before the barrier that precedes a
// rewrite of either)
template <class TT>
__device__ __forceinline__ bool frozen_row(const TT& t, int i) {
    const int v = t.rowvar()[i];
    return v < t.n && t.zst()[v] == FROZEN;
}

template <class TT>
__device__ __forceinline__ void var_bounds(const TT& t, int var, double& lo, double& up) {
    const int n = t.n;
    if (var == 2 * n) { lo = 0.0; up = 1.0; return; }
    if (var < n) { lo = t.l()[var]; up = t.u()[var]; return; }
    const int k = var - n;
    if (t.l()[k] == t.u()[k]) { lo = -QPN_INF; up = QPN_INF; return; }
    const int8_t s = t.zst()[k];
    if (s == AT_L) { lo = 0.0; up = QPN_INF; return; }
    if (s == AT_U) { lo = -QPN_INF; up = 0.0; return; }
    lo = 0.0; up = 0.0;  // z_k basic or floating: w_k is an artificial fixed at 0
}

template <class TT>
__device__ __forceinline__ bool artificial_row(const TT& t, int i) {
    const int v = t.rowvar()[i], n = t.n;
    if (v < n || v == 2 * n) return false;
    const int k = v - n;
    if (t.l()[k] == t.u()[k]) return false;
    const int8_t s = t.zst()[k];
    return s >= FLOATING;                  // FLOATING, BASIC or FROZEN
}

// ---- start of the normal-map path (avi_scratch.jl:17-50) -----------------------------
// Expects T[i][0:n] = -M[i][:] (thread i wrote its own row), t.l(), t.u() filled, q and z0
// readable.  zb = t.prow() is scratch (needs ldr >= n; every caller has cap >= n+1).
// Ends with a barrier.
__device__ __noinline__ void tab_start_core(Tab t, const double* q, const double* z0) {
    const int n = t.n, i = threadIdx.x;
    double* zb = t.prow();
    if (i < n) zb[i] = fmin(fmax(z0[i], t.l()[i]), t.u()[i]);
    QPN_SYNC();
    if (i < n) {
        double* row = t.T() + (size_t)i * t.ldr;
        double acc = 0.0;
        #pragma unroll 4
        for (int j = 0; j < n; ++j) {
            const double mij = -row[j];
            if (mij != 0.0) acc = fma(mij, zb[j], acc);
        }
        const double zi = z0[i], zbi = zb[i];
        const double r = ((acc + q[i]) + zi) - zbi;
        row[n] = -r;                                      // the homotopy column goes last
        t.rr()[i] = r;
        t.beta()[i] = zbi - zi;
        t.rowvar()[i] = n + i;
        t.zst()[i] = (zi <= t.l()[i]) ? AT_L : (zi >= t.u()[i]) ? AT_U : FLOATING;
        t.rowof()[i] = -1; t.colof()[i] = i;
        t.rowof()[n + i] = i; t.colof()[n + i] = -1;
        t.colvar()[i] = i;
        t.nbval()[i] = zbi;
    }
    if (i == 0) {
        t.colvar()[n] = 2 * n; t.nbval()[n] = 0.0;
        t.rowof()[2 * n] = -1; t.colof()[2 * n] = n;
        *t.frozen_src() = nullptr; t.frozen_hdr()[0] = n; t.frozen_hdr()[1] = -1;
    }
    QPN_SYNC();
}
__device__ __forceinline__ void tab_start(Tab& t, const double* q, const double* z0) {
    tab_start_core(t, q, z0);
    t.ncol = t.n + 1;
    t.pivots = 0;
    t.own_frozen = 0;
}

// ---- rank-1 pivot on the compact tableau (avi_scratch.jl:2-7) --------------------------
// Not inlined (one copy keeps the engine's code small enough for the instruction cache: with every
// helper inlined at each call site the robust_avoid kernel stalled mostly on instruction fetch).
// Returns the new live-column count.
__device__ __noinline__ int pivot_core(Tab t, int rho, int c, bool compact) {
    const int n = t.n, ldr = t.ldr, i = threadIdx.x;
    const int ncol = t.ncol;
    const int nce = (ncol + 1) & ~1;                      // even: the update runs two columns at a time
    double* T = t.T();
    const double* prho = T + (size_t)rho * ldr;
    const double p = prho[c];
    // The scaled pivot row goes to prow AND straight back into row rho (element j by the thread that scaled it;
    // element c after the barrier, because every thread still reads p = T[rho][c] above): the update below then
    // needs no per-element "is this the pivot row" select -- row rho runs the same fma with a zero multiplier.
    #pragma unroll 1
    for (int j = i; j < nce; j += blockDim.x) {
        const double v = (j >= ncol) ? 0.0 : (j == c) ? (1.0 / p) : prho[j] / p;
        t.prow()[j] = v;
        if (j < ncol && j != c) T[(size_t)rho * ldr + j] = v;
    }
    const bool live = (i < n) && !t.own_frozen;             // frozen rows are never swept
    const double d = live ? T[(size_t)i * ldr + c] : 0.0;
    const int lv = t.rowvar()[rho];                         // leaving variable (read before the barrier)
    // A slack of a free variable never comes back: its column leaves the live range.
    const bool dead = compact && lv >= n && lv < 2 * n && is_free_var(t, lv - n);
    const int last = ncol - 1;
    QPN_SYNC();
    if (live) {
        double* row = T + (size_t)i * ldr;
        const bool isrho = (i == rho);
        row[c] = isrho ? t.prow()[c] : 0.0;               // then column c follows the common formula
        const double nd = isrho ? 0.0 : -d;               // fma(0, p_j, prow_j) = prow_j: the pivot row stays as written
#pragma unroll 2
        for (int j = 0; j < nce; j += 2) {
            const double2 pj = *reinterpret_cast<const double2*>(t.prow() + j);
            if (pj.x == 0.0 && pj.y == 0.0) continue;     // uniform: both columns untouched by this pivot
            double2 tv = *reinterpret_cast<double2*>(row + j);
            tv.x = fma(nd, pj.x, tv.x);
            tv.y = fma(nd, pj.y, tv.y);
            *reinterpret_cast<double2*>(row + j) = tv;
        }
        if (dead && c != last) row[c] = row[last];        // own row only: no barrier needed
    }
    if (i == 0) {
        const int ev = t.colvar()[c];
        const double vent = t.nbval()[c], vlv = t.beta()[rho];
        t.rowvar()[rho] = ev; t.rowof()[ev] = rho; t.colof()[ev] = -1; t.rowof()[lv] = -1;
        t.beta()[rho] = vent;
        if (dead) {
            t.colof()[lv] = -1;
            if (c != last) {
                const int mv = t.colvar()[last];
                t.colvar()[c] = mv; t.nbval()[c] = t.nbval()[last]; t.colof()[mv] = c;
            }
        } else {
            t.colvar()[c] = lv; t.colof()[lv] = c; t.nbval()[c] = vlv;
        }
    }
    QPN_SYNC();
    return dead ? last : ncol;
}
__device__ __forceinline__ void pivot(Tab& t, int rho, int c, bool compact = true) {
    t.ncol = pivot_core(t, rho, c, compact);
    t.pivots++;
}

__device__ __noinline__ int best_artificial_row(const Tab t, int c) {
    const int i = threadIdx.x;
    double a = 0.0; bool valid = false;
    if (i < t.n && artificial_row(t, i)) {
        a = fabs(t.T()[(size_t)i * t.ldr + c]);
        valid = a > 0.0;
    }
    int idx;
    block_argmax(t, valid, a, idx);
    return (idx >= 0 && a > PIV_TOL) ? idx : -1;
}

// ---- ratio test over the finite bounds of the basics (avi_scratch.jl:65-77) ------------
// Returns the step of the blocking row (INF if none); all threads get the same answer.
// Every exit ends with a barrier.
struct RatioResult { double th; int rho; int which; };
__device__ __noinline__ RatioResult ratio_test_core(const Tab t, int c, double sigma) {
    int rho, which;
    const int n = t.n, i = threadIdx.x;
    double r = QPN_INF, a = 0.0;
    bool is_t = false;
    if (i < n && !t.own_frozen) {                           // a frozen row never blocks (and may not be in T() at all)
        const double ci = t.T()[(size_t)i * t.ldr + c];
        const double d = sigma * ci;
        const int v = t.rowvar()[i];
        double lo, up;
        var_bounds(t, v, lo, up);
        if (d > D_TOL && lo > -QPN_INF) r = fmax((t.beta()[i] - lo) / d, 0.0);
        else if (d < -D_TOL && up < QPN_INF) r = fmax((up - t.beta()[i]) / (-d), 0.0);
        a = fabs(ci);
        is_t = (v == 2 * n);
    }
    const double theta = block_min(t, r);
    rho = -1; which = 0;
    if (theta == QPN_INF) { QPN_SYNC(); return RatioResult{QPN_INF, -1, 0}; }   // every exit ends with a barrier
    const double cut = theta + TIE_TOL * (1.0 + theta);
    // among ties: t first (so the path terminates), then largest |d|, then lowest row
    const bool cand = (i < n) && (r <= cut);
    double key = is_t ? QPN_INF : a;
    int idx;
    block_argmax(t, cand, key, idx);
    rho = idx;
    if (blockDim.x > 32) QPN_SYNC();
    if (i == rho) t.red_d()[32] = r;                        // the winner publishes its own ratio
    QPN_SYNC();
    const double th = t.red_d()[32];
    which = (sigma * t.T()[(size_t)rho * t.ldr + c] > 0.0) ? -1 : +1;
    return RatioResult{th, rho, which};
}
__device__ __forceinline__ double ratio_test(const Tab& t, int c, double sigma, int& rho, int& which) {
    const RatioResult r = ratio_test_core(t, c, sigma);
    rho = r.rho; which = r.which;
    return r.th;
}

__device__ __forceinline__ void move(Tab& t, int c, double sigma, double theta) {
    if (theta == 0.0) return;
    const int i = threadIdx.x;
    if (i < t.n && !t.own_frozen) {                         // a frozen row keeps the basic value it had at the freeze
        const double ci = t.T()[(size_t)i * t.ldr + c];
        if (ci != 0.0) t.beta()[i] = fma(-(sigma * theta), ci, t.beta()[i]);
    }
    if (i == 0) t.nbval()[c] = fma(sigma, theta, t.nbval()[c]);
    QPN_SYNC();
}

template <class TT>
__device__ __forceinline__ void leave_at(TT& t, int rho, int which) {
    if (threadIdx.x == 0) {
        double lo, up;
        var_bounds(t, t.rowvar()[rho], lo, up);
        t.beta()[rho] = which < 0 ? lo : up;
    }
    QPN_SYNC();
}

template <class TT>
__device__ __forceinline__ void set_zst(TT& t, int k, int8_t s) {
    QPN_SYNC();
    if (threadIdx.x == 0) t.zst()[k] = s;
    QPN_SYNC();
}

template <class TT>
__device__ __forceinline__ bool try_exchange(TT& t, int var) {
    const int c = t.colof()[var];
    if (c < 0) return false;
    const int rho = best_artificial_row(t, c);
    if (rho < 0) return false;
    pivot(t, rho, c);
    return true;
}

// Largest |T[i][c]| over rows still holding the slack of a FREE variable (artificial whatever
// the start point is).
__device__ __forceinline__ int best_free_row(const Tab& t, int c) {
    const int i = threadIdx.x, n = t.n;
    double a = 0.0; bool valid = false;
    if (i < n) {
        const int v = t.rowvar()[i];
        if (v >= n && v < 2 * n && is_free_var(t, v - n)) {
            a = fabs(t.T()[(size_t)i * t.ldr + c]);
            valid = a > 0.0;
        }
    }
    int idx;
    block_argmax(t, valid, a, idx);
    return (idx >= 0 && a > PIV_TOL) ? idx : -1;
}

// T[:, t] = B^-1 r rebuilt from the slack columns: the column of a nonbasic w_k is -B^-1 e_k, a
// w_k basic in row rho means B^-1 e_k = -e_rho.  Sequential fma over k.  Needs every slack
// column still in place (no compaction yet).  Ends with a barrier.
__device__ __forceinline__ void recompute_tcol(Tab& t) {
    const int n = t.n, i = threadIdx.x;
    const int tc = t.colof()[2 * n];
    if (i < n) {
        double* row = t.T() + (size_t)i * t.ldr;
        double acc = 0.0;
        #pragma unroll 4
        for (int k = 0; k < n; ++k) {
            const int ck = t.colof()[n + k];
            const double pik = ck >= 0 ? -row[ck] : (t.rowof()[n + k] == i ? -1.0 : 0.0);
            if (pik != 0.0) acc = fma(pik, t.rr()[k], acc);
        }
        row[tc] = acc;
    }
    QPN_SYNC();
}

// Drop every dead column (slack of a free variable) from the live range at once.
__device__ __forceinline__ void compact_dead(Tab& t) {
    const int n = t.n, i = threadIdx.x;
    int* map = reinterpret_cast<int*>(t.prow());            // scratch: prow is free between pivots
    if (i == 0) {
        int nl = 0;
        #pragma unroll 1
        for (int j = 0; j < t.ncol; ++j) {
            const int v = t.colvar()[j];
            if (v >= n && v < 2 * n && is_free_var(t, v - n)) t.colof()[v] = -1;
            else map[nl++] = j;
        }
        #pragma unroll 1
        for (int d = 0; d < nl; ++d) {
            const int sidx = map[d];
            if (sidx != d) { const int v = t.colvar()[sidx]; t.colvar()[d] = v; t.nbval()[d] = t.nbval()[sidx]; t.colof()[v] = d; }
        }
        t.red_i()[32] = nl;
    }
    QPN_SYNC();
    const int nl = t.red_i()[32];
    if (i < n) {
        double* row = t.T() + (size_t)i * t.ldr;
        #pragma unroll 1
        for (int d = 0; d < nl; ++d) { const int sidx = map[d]; if (sidx != d) row[d] = row[sidx]; }   // sidx >= d: in place
    }
    t.ncol = nl;
    QPN_SYNC();
}

// after the marks of freeze(): thread i keeps "my row is frozen" in a register for the rest of the solve
__device__ __forceinline__ void freeze_hook(Tab& t) { t.own_frozen = (threadIdx.x < t.n) && frozen_row(t, threadIdx.x); }

// ---- freeze (oracle/avi_pivot.py: freeze) ------------------------------------------------
// Called where phase 0 of the crash has just ended (in this solve or in its plan): marks the free basics FROZEN
// and records which variable sits in which live column and at what value.  Ends with a barrier.
template <class TT>
__device__ __forceinline__ void freeze(TT& t) {
    const int n = t.n;
    #pragma unroll 1
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int v = t.rowvar()[i];
        if (v < n && is_free_var(t, v)) t.zst()[v] = FROZEN;
    }
    #pragma unroll 1
    for (int j = threadIdx.x; j < t.ncol; j += blockDim.x) { t.colvar0()[j] = t.colvar()[j]; t.nbval0()[j] = t.nbval()[j]; }
    t.ncol0 = t.ncol;
    freeze_hook(t);
    QPN_SYNC();
}

// frozen_values (below) for the shared-memory workspace: the same sum, with the frozen rows read from the plan when
// the instance never copied them (t.T0f); their homotopy entry (B^-1 r)_i was left in birv() by the plan start.
// Ends with a barrier.
__device__ __noinline__ void frozen_values(Tab t, double* dx) {
    const int n = t.n, nc0 = t.ncol0, ldr = t.ldr;
    const double* T0f = *t.frozen_src();
    const int nact = t.frozen_hdr()[0], tcol0 = t.frozen_hdr()[1];
    #pragma unroll 1
    for (int j = threadIdx.x; j < nc0; j += blockDim.x) {
        const int v = t.colvar0()[j];
        const int r = t.rowof()[v], c = t.colof()[v];
        dx[j] = (r >= 0 ? t.beta()[r] : c >= 0 ? t.nbval()[c] : t.nbval0()[j]) - t.nbval0()[j];
    }
    QPN_SYNC();
    #pragma unroll 1
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (!frozen_row(t, i)) continue;
        double acc = t.beta()[i];
        if (T0f && i >= nact) {
            const double* row = T0f + (size_t)i * ldr;
            const double bir = t.birv()[i];               // written by the plan start, in the same pass as the swept rows' entries
            #pragma unroll 4
            for (int j = 0; j < nc0; ++j) acc = fma(-(j == tcol0 ? bir : row[j]), dx[j], acc);
        } else {
            const double* row = t.T() + (size_t)i * ldr;
            #pragma unroll 4
            for (int j = 0; j < nc0; ++j) acc = fma(-row[j], dx[j], acc);
        }
        t.beta()[i] = acc;                                // nobody reads a frozen row's beta in this pass
    }
    QPN_SYNC();
}

// Final values of the frozen variables from their phase-0 rows:
//   x_B[i] = beta[i] - sum_j T0[i][j] * (x(colvar0[j]) - nbval0[j]),  sequential fma over the live columns of the freeze.
// (Columns that were already dead then carry a variable that stays at its value: they add fma(., 0, acc) = acc in the
// specification and are simply absent here.)  `dx`: ncol0 doubles of scratch.  Ends with a barrier.
template <class TT>
__device__ __forceinline__ void frozen_values(TT& t, double* dx) {
    const int n = t.n, nc0 = t.ncol0;
    #pragma unroll 1
    for (int j = threadIdx.x; j < nc0; j += blockDim.x) {
        const int v = t.colvar0()[j];
        const int r = t.rowof()[v], c = t.colof()[v];
        // (a variable of a live column of the freeze is no slack of a free variable, so it cannot have retired since;
        // the guard only keeps a corrupted state from indexing out of range)
        dx[j] = (r >= 0 ? t.beta()[r] : c >= 0 ? t.nbval()[c] : t.nbval0()[j]) - t.nbval0()[j];
    }
    QPN_SYNC();
    #pragma unroll 1
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (!frozen_row(t, i)) continue;
        double acc = t.beta()[i];
        #pragma unroll 4
        for (int j = 0; j < nc0; ++j) acc = fma(-t.frozen_entry(i, j), dx[j], acc);
        t.beta()[i] = acc;                                // nobody reads a frozen row's beta in this pass
    }
    QPN_SYNC();
}

// ---- crash: bring interior / free variables into the basis ----------------------------
template <class TT>
__device__ __forceinline__ void crash(TT& t, bool from_plan) {
    const int n = t.n;
    // phase 0: free variables exchange against rows of free variables only.  Nothing here depends
    // on the start point or on q, so a plan (shared matrix) has done it already: running it again
    // would retry the variables that found no pivot at their turn, which the specification does not.
    if (!from_plan) {
        const int piv0 = t.pivots;
        #pragma unroll 1
        for (int i = 0; i < n; ++i) {
            if (!is_free_var(t, i)) continue;
            const int c = t.colof()[i];
            const int rho = best_free_row(t, c);
            if (rho >= 0) { pivot(t, rho, c, false); set_zst(t, i, BASIC); }
        }
        if (t.pivots > piv0) { recompute_tcol(t); compact_dead(t); }
    }
    freeze(t);
    // phase 1: everything still floating, against any artificial row
    #pragma unroll 1
    for (int i = 0; i < n; ++i) {
        if (t.zst()[i] != FLOATING) continue;
        const int c = t.colof()[i];
        const int rho = best_artificial_row(t, c);
        if (rho >= 0) { pivot(t, rho, c); set_zst(t, i, BASIC); continue; }
        // dependent column: walk towards an extreme point (Cao-Ferris stage 2)
        int rb[2], wb[2]; double th[2], own[2], step[2];
        own[0] = t.u()[i] - t.nbval()[c];        // read before the ratio tests (synchronisation rule)
        own[1] = t.nbval()[c] - t.l()[i];
        #pragma unroll 1
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
            if (threadIdx.x == 0) { t.nbval()[c] = s == 0 ? t.u()[i] : t.l()[i]; t.zst()[i] = s == 0 ? AT_U : AT_L; }
            QPN_SYNC();
            continue;
        }
        move(t, c, sigma, th[s]);
        leave_at(t, rb[s], wb[s]);
        const int lv = t.rowvar()[rb[s]];
        pivot(t, rb[s], c);
        set_zst(t, i, BASIC);
        if (lv < n) {
            set_zst(t, lv, wb[s] < 0 ? AT_L : AT_U);
            if (t.rowof()[n + lv] < 0) try_exchange(t, n + lv);
        } else {
            const int k = lv - n;
            if (try_exchange(t, k)) set_zst(t, k, BASIC);
            else try_exchange(t, n + k);
        }
    }
}

// Is any row still artificial?  (block-uniform; contains a barrier)
__device__ __forceinline__ bool any_artificial_row(const Tab& t) {
    int any = 0;
    if (threadIdx.x < t.n) any = artificial_row(t, threadIdx.x) ? 1 : 0;
    return QPN_SYNC_OR(any) != 0;
}

template <class TT>
__device__ __forceinline__ void repair(TT& t) {
    const int n = t.n;
    bool progress = true;
    #pragma unroll 1
    while (progress) {
        progress = false;
        if (!any_artificial_row(t)) return;
        #pragma unroll 1
        for (int k = 0; k < n; ++k) {
            const int8_t s = t.zst()[k];
            if ((s == AT_L || s == AT_U) && t.l()[k] != t.u()[k] && t.rowof()[k] < 0 && t.rowof()[n + k] < 0) {
                if (try_exchange(t, n + k)) progress = true;
                else if (try_exchange(t, k)) { set_zst(t, k, BASIC); progress = true; }
            }
        }
        #pragma unroll 1
        for (int k = 0; k < n; ++k)
            if (t.zst()[k] == FLOATING && t.rowof()[k] < 0)
                if (try_exchange(t, k)) { set_zst(t, k, BASIC); progress = true; }
    }
}

// ---- phase 2: complementary pivoting (avi_scratch.jl:59-132) ------------------------------
template <class TT>
__device__ __forceinline__ int lemke(TT& t, int max_pivots) {
    const int n = t.n;
    int ent = 2 * n; double sigma = 1.0;
    #pragma unroll 1
    for (;;) {
        if (t.pivots > max_pivots) return ST_MAX_ITERS;
        const int c = t.colof()[ent];
        int rb, wb;
        // Read everything the branch below depends on BEFORE the ratio test: its barriers then
        // separate these reads from the writes in move() (a read after it would race with them).
        const double own = ent == 2 * n ? 1.0 - t.nbval()[c] : ent < n ? (t.u()[ent] - t.l()[ent]) : QPN_INF;
        const double th = ratio_test(t, c, sigma, rb, wb);
        if (own == QPN_INF && th == QPN_INF) return ST_RAY_TERM;
        if (own <= th) {
            move(t, c, sigma, own);
            QPN_SYNC();
            if (ent == 2 * n) { if (threadIdx.x == 0) t.nbval()[c] = 1.0; QPN_SYNC(); return ST_SUCCESS; }
            if (threadIdx.x == 0) { t.nbval()[c] = sigma > 0 ? t.u()[ent] : t.l()[ent]; t.zst()[ent] = sigma > 0 ? AT_U : AT_L; }
            QPN_SYNC();
            ent = n + ent; sigma = -sigma;
            continue;
        }
        move(t, c, sigma, th);
        leave_at(t, rb, wb);
        const int lv = t.rowvar()[rb];
        const bool was_art = artificial_row(t, rb);
        pivot(t, rb, c);
        if (ent < n) set_zst(t, ent, BASIC);
        if (lv == 2 * n) return wb > 0 ? ST_SUCCESS : ST_RAY_TERM;
        if (lv < n) {
            set_zst(t, lv, wb < 0 ? AT_L : AT_U);
            if (t.rowof()[n + lv] >= 0) return ST_FAILURE;
            ent = n + lv; sigma = wb < 0 ? 1.0 : -1.0;
        } else {
            const int k = lv - n;
            const int8_t s = t.zst()[k];
            if (was_art || t.rowof()[k] >= 0 || !(s == AT_L || s == AT_U)) return ST_FAILURE;
            ent = k; sigma = s == AT_L ? 1.0 : -1.0;
        }
    }
}

// Runs crash + repair + path following on a started tableau.  Thread i < n gets z_i and its
// basis code (1 lower, 2 basic/interior, 3 upper, 4 fixed).  Deliberately NOT inlined: a kernel
// that solves several AVIs per instance then holds one copy of the engine, and the workspace
// descriptor travels in registers instead of a local-memory struct.
struct PivotResult { double zi; int st; int pivots; int code; };
__device__ __noinline__ PivotResult avi_pivot_run(Tab t, int max_pivots, bool from_plan) {
    crash(t, from_plan);
    repair(t);
    PivotResult out;
    out.st = lemke(t, max_pivots);
    frozen_values(t, t.prow());                           // prow is free between pivots
    out.zi = 0.0; out.code = 0; out.pivots = t.pivots;
    const int i = threadIdx.x;
    if (i < t.n) {
        const int r = t.rowof()[i];
        out.zi = r >= 0 ? t.beta()[r] : t.nbval()[t.colof()[i]];
        const int8_t s = t.zst()[i];
        out.code = (t.l()[i] == t.u()[i]) ? 4 : (r >= 0 || s == FLOATING) ? 2 : (s == AT_L ? 1 : 3);
    }
    return out;
}

// check_avi_solution (avi.jl:148-156) contribution of one index.
__device__ __forceinline__ int check_avi_index(double r, double z, double l, double u, double tol) {
    int bad = 0;
    if (r > tol && fabs(z - l) > tol) bad++;
    if (r < -tol && fabs(z - u) > tol) bad++;
    if (z - l < -tol) bad++;
    if (z - u > tol) bad++;
    return bad;
}

}  // namespace qpn
