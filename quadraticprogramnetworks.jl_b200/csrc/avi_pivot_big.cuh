// The same complementary-pivoting engine (avi_pivot.cuh) for AVIs whose compact tableau does
// not fit shared memory (n > ~166) or has more rows than a CTA may have threads: one CTA per
// instance, the fp64 tableau in GLOBAL memory (row-major, one workspace slot per resident CTA,
// L2 / HBM streamed), every per-row / per-column vector still in shared memory.
//
// Reference sizes this covers: robust_avoid levels 1-2 when pieces make
// them wide, the synthetic 3-level chain (d1 = 128 / 256 / 384) and the n = 256, m = 512
// monotone stress AVI (lifted n = 1,280 or 1,536); /root/reference/src/avi.jl:63-77 is still the
// call site being replaced.
//
// Work split: a pivot is one pass over the live tableau -- warps take rows (stride = warps per
// CTA), lanes take pairs of columns (128-bit accesses, fully coalesced); the scaled pivot row
// and the entering column are staged in shared memory first.  Rows whose entering-column entry
// is zero and column pairs whose pivot-row entries are zero are skipped without being read, so a
// sparse tableau costs its non-zero work only.  The control flow (crash, repair, path following)
// is the very same template code as the shared-memory engine, and each tableau entry is still
// updated by exactly one fma(-d_i, prow_j, T_ij): results are bit-identical to the CPU oracle.
//
// Delayed updates: a pivot does not sweep the tableau at once.  Its scaled pivot row and its entering column are
// queued (QPN_BIG_PEND of them, in the CTA's slot, L1 / L2 resident); the entering column and the pivot row the
// NEXT pivot needs are read from the stale tableau and brought up to date on the fly (one fma per queued update);
// when the queue is full, ONE pass over the tableau applies all queued rank-1 updates, in order, from registers.
// Every entry still sees the same sequence of fma(-d_i, prow_j, T_ij) -- only later -- so the bits do not
// change, while the tableau crosses HBM once per QPN_BIG_PEND pivots instead of once per pivot.
//
// Roofline: HBM / L2 bandwidth.  Algorithmic bytes: 16 * n * ncol per QPN_BIG_PEND pivots (read + write of the
// live tableau), + 8 * (n + ncol) per pivot for the staged column and row.
#pragma once
#include "avi_pivot.cuh"

#ifdef QPN_BIG_DEBUG
#include <cstdio>
#define BIGCHK(cond, what, a, b)                                                                                   \
    do {                                                                                                           \
        if (!(cond)) {                                                                                             \
            if (threadIdx.x == 0) printf("BIGCHK block %d line %d: %s (%d, %d) n=%d ncol=%d piv=%d\n", blockIdx.x, __LINE__, what, (int)(a), (int)(b), t.n, t.ncol, t.pivots); \
            __syncthreads();                                                                                       \
            __trap();                                                                                              \
        }                                                                                                          \
    } while (0)
#else
#define BIGCHK(cond, what, a, b) do { } while (0)
#endif

#ifndef QPN_BIG_PEND
#define QPN_BIG_PEND 4
#endif
#ifndef QPN_BIG_INFLIGHT
#define QPN_BIG_INFLIGHT 4      // predicated 128-bit tableau loads a lane keeps in flight in the flush (5 and 6 measured: within noise)
#endif

namespace qpn {

struct BigTab {
    Tab v;              // shared-memory vectors (carved with no tableau inside: td = 0)
    double* Tg;         // this CTA's tableau slot in global memory: n x ldr, row-major, negated (stale by `npend` pivots)
    double* Pd;         // queued entering columns: QPN_BIG_PEND x nmax   (same slot, after the tableau)
    double* Pp;         // queued scaled pivot rows: QPN_BIG_PEND x ldrmax
    int dcol_off;       // byte offset in qpn_smem of the entering-column cache (nmax doubles)
    int n, ldr, ncol, pivots;
    int cc, cpiv;       // which column the cache holds and at which pivot count it was read
    int stage_off, stage_cap;   // spare dynamic shared memory (byte offset, doubles): the queue's pivot rows are staged there for a flush
    int npend;          // queued pivots (block-uniform)
    int prw[QPN_BIG_PEND], pcl[QPN_BIG_PEND];    // their pivot rows / entering columns
    double* nbv0;       // freeze record (avi_pivot.cuh: freeze), in the slot: nonbasic values ...
    int* cv0;           // ... and column variables at the freeze
    int ncol0;
    // Started from a plan that exported the swept rows first: only rows 0 .. nact-1 are in the slot, the frozen rows
    // nact .. n-1 are read from the plan itself (T0f, same row stride; their homotopy entry is (B^-1 r)_i from PTf).
    int nact, tcol0;
    const double* T0f;
    const double* PTf;

    __device__ __forceinline__ double* prow() const { return v.prow(); }
    __device__ __forceinline__ double* nbval() const { return v.nbval(); }
    __device__ __forceinline__ double* beta() const { return v.beta(); }
    __device__ __forceinline__ double* l() const { return v.l(); }
    __device__ __forceinline__ double* u() const { return v.u(); }
    __device__ __forceinline__ double* rr() const { return v.rr(); }
    __device__ __forceinline__ int* rowvar() const { return v.rowvar(); }
    __device__ __forceinline__ int* colvar() const { return v.colvar(); }
    __device__ __forceinline__ int* rowof() const { return v.rowof(); }
    __device__ __forceinline__ int* colof() const { return v.colof(); }
    __device__ __forceinline__ int8_t* zst() const { return v.zst(); }
    __device__ __forceinline__ double* dcol() const { return reinterpret_cast<double*>(qpn_smem + dcol_off); }
    __device__ __forceinline__ double* nbval0() const { return nbv0; }
    __device__ __forceinline__ int* colvar0() const { return cv0; }
    // a frozen row is never swept and is never part of a queued update: its slot entries are those of the freeze
    __device__ __forceinline__ double frozen_entry(int i, int j) const { return Tg[(size_t)i * ldr + j]; }
};

// Shared-memory bytes of the vectors of a big tableau with up to nmax rows (ldr of the full
// n x (n+1) shape) plus the column cache.
__host__ __device__ __forceinline__ size_t big_smem_bytes(int nmax) {
    return tab_smem_bytes_ex(nmax, 0, row_stride(nmax + 1), 0) + 8 * (size_t)((nmax + 1) & ~1);
}
// Doubles of one workspace slot whose tableau area holds `tcap` doubles (even): tableau, the queue, the freeze record.
__host__ __device__ __forceinline__ size_t big_slot_doubles_ex(int nmax, size_t tcap) {
    return tcap + (size_t)QPN_BIG_PEND * ((size_t)((nmax + 1) & ~1) + row_stride(nmax + 1)) +
           row_stride(nmax + 1) + (size_t)row_stride(nmax + 1) / 2 + 1;          // stride/2 is odd: the slot stays a multiple of 16 bytes
}
// ... with room for the full n x (n+1) tableau (a solve without a plan).
__host__ __device__ __forceinline__ size_t big_slot_doubles(int nmax) {
    return big_slot_doubles_ex(nmax, (size_t)nmax * row_stride(nmax + 1));
}

// Returns the byte offset just past the workspace.
// tcap: doubles of the slot's tableau area (0: the full nmax x (nmax+1) shape).
__device__ __forceinline__ int big_carve(BigTab& t, int nmax, double* slot, int base_off, size_t tcap = 0) {
    tab_carve_ex(t.v, nmax, 0, row_stride(nmax + 1), base_off, 0);
    t.dcol_off = base_off + (int)tab_smem_bytes_ex(nmax, 0, row_stride(nmax + 1), 0);
    t.Tg = slot;
    t.Pd = slot + (tcap ? tcap : (size_t)nmax * row_stride(nmax + 1));
    t.Pp = t.Pd + (size_t)QPN_BIG_PEND * ((nmax + 1) & ~1);
    t.nbv0 = t.Pp + (size_t)QPN_BIG_PEND * row_stride(nmax + 1);
    t.cv0 = reinterpret_cast<int*>(t.nbv0 + row_stride(nmax + 1));
    t.ncol0 = 0; t.nact = nmax; t.tcol0 = -1; t.T0f = nullptr; t.PTf = nullptr;
    t.n = nmax; t.ldr = row_stride(nmax + 1); t.ncol = 0; t.pivots = 0; t.cc = -1; t.cpiv = -1; t.npend = 0;
    t.stage_off = 0; t.stage_cap = 0;
    return base_off + (int)big_smem_bytes(nmax);
}
// Whatever dynamic shared memory the launch got beyond the `used` bytes of the kernel's own layout.
__device__ __forceinline__ void big_stage_carve(BigTab& t, int used) {
    unsigned total;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(total));
    t.stage_off = (used + 15) & ~15;
    t.stage_cap = (int)total > t.stage_off ? ((int)total - t.stage_off) / 8 : 0;
}
// A column cached before the freeze still holds the entries of the rows that are frozen now.
__device__ __forceinline__ void freeze_hook(BigTab& t) { t.cc = -1; }
// Shape of the next solve; whatever was queued belongs to a tableau that is about to be overwritten.
__device__ __forceinline__ void big_shape(BigTab& t, int n, int cap) {
    t.n = n; t.ldr = row_stride(cap); t.cc = -1; t.npend = 0; t.nact = n; t.T0f = nullptr; t.PTf = nullptr; t.tcol0 = -1;
}
__device__ __forceinline__ int big_pd_stride(const BigTab& t) { return (t.v.nmax + 1) & ~1; }

// One queued update applied to the tableau entry (row r, column j) whose current value is v.
__device__ __forceinline__ double big_apply(const BigTab& t, int l, int r, int j, double v, double dlr, double plj) {
    if (r == t.prw[l]) return plj;                       // the pivot row was replaced by its scaled self
    if (dlr == 0.0) return v;                            // rows with a zero entering entry are not touched
    return fma(-dlr, plj, j == t.pcl[l] ? 0.0 : v);      // column c restarts from 0 (avi_pivot.cuh: row[c] = 0)
}

// ---- entering column into shared memory ------------------------------------------------------
// Block-uniform; ends with a barrier when it had to read.
__device__ __forceinline__ void big_col(BigTab& t, int c) {
    BIGCHK(c >= 0 && c < t.ncol, "big_col: column out of range", c, t.ncol);
    if (t.cc == c && t.cpiv == t.pivots) return;
    const int n = t.n, ldr = t.ldr, np = t.npend, pds = big_pd_stride(t), lds = t.v.ldrmax;
    double* d = t.dcol();
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
        // A frozen row (avi_pivot.cuh: FROZEN) is out of every sweep: a zero entering entry keeps it out of the ratio
        // test, of move() and -- through the queued column -- of the flush, and it is not even read.
        if (frozen_row(t, r)) { d[r] = 0.0; continue; }
        double v = t.Tg[(size_t)r * ldr + c];
        for (int l = 0; l < np; ++l)                     // bring the stale entry up to date, oldest update first
            v = big_apply(t, l, r, c, v, t.Pd[(size_t)l * pds + r], t.Pp[(size_t)l * lds + c]);
        d[r] = v;
    }
    t.cc = c; t.cpiv = t.pivots;
    QPN_SYNC();
}

// Apply every queued update to the tableau in one pass (in order, from registers).  When the LAST queued pivot
// retires its column (`dead_c` >= 0: the slack of a free variable left the basis), column `dead_last` is moved
// into its place after the updates, as the immediate form does (avi_pivot.cuh: row[c] = row[last]).
// Ends with a barrier.
__device__ __noinline__ void big_flush(BigTab& t, int dead_c = -1, int dead_last = -1) {
    const int np = t.npend;
    if (np == 0) return;
    const int n = t.n, ldr = t.ldr, nce = (t.ncol + 1) & ~1, pds = big_pd_stride(t);
    // the queued pivot rows are read once per tableau row: from shared memory when the launch has room to stage them
    const bool staged = np * nce <= t.stage_cap;
    const int lds = staged ? nce : t.v.ldrmax;
    const double* PP = t.Pp;
    if (staged) {
        double* st = reinterpret_cast<double*>(qpn_smem + t.stage_off);
        for (int e = threadIdx.x; e < np * nce; e += blockDim.x) { const int l = e / nce, j = e - l * nce; st[e] = t.Pp[(size_t)l * t.v.ldrmax + j]; }
        PP = st;
        QPN_SYNC();
    }
    // column pairs no queued update touches are neither read nor written
    unsigned char* touched = reinterpret_cast<unsigned char*>(t.prow());
    for (int q = threadIdx.x; q < nce / 2; q += blockDim.x) {
        int any = 0;
        for (int l = 0; l < np; ++l) {
            const double2 pj = *reinterpret_cast<const double2*>(PP + (size_t)l * lds + 2 * q);
            any |= (pj.x != 0.0 || pj.y != 0.0);
        }
        touched[q] = (unsigned char)any;
    }
    QPN_SYNC();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const bool mv = dead_c >= 0 && dead_c != dead_last;
    for (int i = w; i < n; i += nw) {
        if (frozen_row(t, i)) continue;                  // never swept (warp-uniform; two shared-memory reads instead of the queue's column)
        double nd[QPN_BIG_PEND];
        int kind[QPN_BIG_PEND];                          // 0: row untouched by update l, 1: fma, 2: row replaced (pivot row)
        bool any = false;
#pragma unroll
        for (int l = 0; l < QPN_BIG_PEND; ++l) {
            const double dl = l < np ? t.Pd[(size_t)l * pds + i] : 0.0;
            nd[l] = -dl;
            kind[l] = l >= np ? 0 : (i == t.prw[l]) ? 2 : (dl != 0.0) ? 1 : 0;
            any |= kind[l] != 0;
        }
        double* row = t.Tg + (size_t)i * ldr;
        if (any) {
            for (int j0 = 2 * lane; j0 < nce; j0 += 64 * QPN_BIG_INFLIGHT) {
                double2 tv[QPN_BIG_INFLIGHT];
                bool act[QPN_BIG_INFLIGHT];
#pragma unroll
                for (int q = 0; q < QPN_BIG_INFLIGHT; ++q) {
                    const int j = j0 + 64 * q;
                    act[q] = j < nce && touched[j >> 1];
                    if (act[q]) tv[q] = *reinterpret_cast<const double2*>(row + j);
                }
#pragma unroll
                for (int l = 0; l < QPN_BIG_PEND; ++l) {
                    if (kind[l] == 0) continue;           // warp-uniform
                    const double* ppl = PP + (size_t)l * lds;
                    const int cz = t.pcl[l];
                    if (kind[l] == 2) {
#pragma unroll
                        for (int q = 0; q < QPN_BIG_INFLIGHT; ++q)
                            if (act[q]) tv[q] = *reinterpret_cast<const double2*>(ppl + j0 + 64 * q);
                    } else {
#pragma unroll
                        for (int q = 0; q < QPN_BIG_INFLIGHT; ++q) {
                            if (!act[q]) continue;
                            const int j = j0 + 64 * q;
                            const double2 pj = *reinterpret_cast<const double2*>(ppl + j);
                            if (j == cz) tv[q].x = 0.0;
                            if (j + 1 == cz) tv[q].y = 0.0;
                            tv[q].x = fma(nd[l], pj.x, tv[q].x);
                            tv[q].y = fma(nd[l], pj.y, tv[q].y);
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < QPN_BIG_INFLIGHT; ++q)
                    if (act[q]) *reinterpret_cast<double2*>(row + j0 + 64 * q) = tv[q];
            }
        }
        if (mv) {                                         // (frozen rows, which keep the column layout of the freeze, were skipped above)
            __syncwarp();
            if (lane == 0) row[dead_c] = row[dead_last];
        }
    }
    t.npend = 0;
    QPN_SYNC();
}

// ---- start of the normal-map path (avi_scratch.jl:17-50) -------------------------------------
// Expects Tg[i][0:n] = -M[i][:], t.l(), t.u() filled, q and z0 readable (shared or global).
// Md (may be null): the dense column-major matrix itself (coalesced row products).  Ends with a barrier.
__device__ __noinline__ void big_start(BigTab& t, const double* Md, const double* q, const double* z0) {
    const int n = t.n, ldr = t.ldr;
    double* zb = t.prow();
    for (int i = threadIdx.x; i < n; i += blockDim.x) zb[i] = fmin(fmax(z0[i], t.l()[i]), t.u()[i]);
    QPN_SYNC();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double* row = t.Tg + (size_t)i * ldr;
        double acc = 0.0;
        if (Md) {
            for (int j = 0; j < n; ++j) {
                const double mij = Md[(size_t)j * n + i];
                if (mij != 0.0) acc = fma(mij, zb[j], acc);
            }
        } else {
            for (int j = 0; j < n; ++j) {
                const double mij = -row[j];
                if (mij != 0.0) acc = fma(mij, zb[j], acc);
            }
        }
        const double zi = z0[i], zbi = zb[i];
        const double r = ((acc + q[i]) + zi) - zbi;
        row[n] = -r;
        if (n + 1 < ldr) row[n + 1] = 0.0;
        t.rr()[i] = r;
        t.beta()[i] = zbi - zi;
        t.rowvar()[i] = n + i;
        t.zst()[i] = (zi <= t.l()[i]) ? AT_L : (zi >= t.u()[i]) ? AT_U : FLOATING;
        t.rowof()[i] = -1; t.colof()[i] = i;
        t.rowof()[n + i] = i; t.colof()[n + i] = -1;
        t.colvar()[i] = i;
    }
    QPN_SYNC();
    for (int i = threadIdx.x; i < n; i += blockDim.x) t.nbval()[i] = zb[i];      // zb aliases prow, nbval is separate
    if (threadIdx.x == 0) {
        t.colvar()[n] = 2 * n; t.nbval()[n] = 0.0;
        t.rowof()[2 * n] = -1; t.colof()[2 * n] = n;
    }
    t.ncol = n + 1; t.pivots = 0; t.cc = -1; t.npend = 0;
    QPN_SYNC();
}

// rank-1 pivot (avi_scratch.jl:2-7), queued: the scaled pivot row and the entering column go to the slot's queue,
// the basis bookkeeping happens now, the tableau sweep when the queue is full -- or at once when this pivot
// retires its column (the slack of a free variable never comes back).  Ends with a barrier.
__device__ __noinline__ void big_pivot(BigTab& t, int rho, int c, bool compact) {
    BIGCHK(rho >= 0 && rho < t.n, "big_pivot: row out of range", rho, c);
    big_col(t, c);                                        // current entering column (up to date) in dcol
    const int n = t.n, ldr = t.ldr, ncol = t.ncol, nce = (ncol + 1) & ~1;
    const int np = t.npend, pds = big_pd_stride(t), lds = t.v.ldrmax;
    const double* dc = t.dcol();
    const double p = dc[rho];
    double* pp = t.Pp + (size_t)np * lds;
    double* pd = t.Pd + (size_t)np * pds;
    for (int j = threadIdx.x; j < nce; j += blockDim.x) {
        double v = 0.0;
        if (j < ncol) {
            double raw = t.Tg[(size_t)rho * ldr + j];     // the pivot row, brought up to date entry by entry
            for (int l = 0; l < np; ++l)
                raw = big_apply(t, l, rho, j, raw, t.Pd[(size_t)l * pds + rho], t.Pp[(size_t)l * lds + j]);
            v = (j == c) ? (1.0 / p) : raw / p;
        }
        pp[j] = v;
    }
    for (int r = threadIdx.x; r < n; r += blockDim.x) pd[r] = dc[r];
    const int lv = t.rowvar()[rho];
    const bool dead = compact && lv >= n && lv < 2 * n && is_free_var(t, lv - n);
    const int last = ncol - 1;
    QPN_SYNC();                                           // every thread has read rowvar[rho] before thread 0 rewrites it
    if (threadIdx.x == 0) {
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
    t.prw[np] = rho; t.pcl[np] = c;
    t.npend = np + 1;
    t.pivots++;
    QPN_SYNC();
    if (dead) { big_flush(t, c, last); t.ncol = last; t.cc = -1; }
    else if (t.npend == QPN_BIG_PEND) big_flush(t);
}
__device__ __forceinline__ void pivot(BigTab& t, int rho, int c, bool compact = true) { big_pivot(t, rho, c, compact); }

// Largest |T[i][c]| over the rows selected by `pick`; ties -> lowest row.  -1 when below PIV_TOL.
template <class Pick>
__device__ __forceinline__ int big_best_row(BigTab& t, int c, Pick pick) {
    big_col(t, c);
    const double* dc = t.dcol();
    double a = 0.0; int idx = -1;
    for (int r = threadIdx.x; r < t.n; r += blockDim.x) {
        if (!pick(r)) continue;
        const double v = fabs(dc[r]);
        if (v > a) { a = v; idx = r; }
    }
    block_argmax_idx(t.v, idx >= 0, a, idx);
    return (idx >= 0 && a > PIV_TOL) ? idx : -1;
}
__device__ __noinline__ int best_artificial_row(BigTab& t, int c) {
    return big_best_row(t, c, [&](int r) { return artificial_row(t, r); });
}
__device__ __noinline__ int best_free_row(BigTab& t, int c) {
    const int n = t.n;
    return big_best_row(t, c, [&](int r) { const int v = t.rowvar()[r]; return v >= n && v < 2 * n && is_free_var(t, v - n); });
}

__device__ __forceinline__ bool any_artificial_row(const BigTab& t) {
    int any = 0;
    for (int r = threadIdx.x; r < t.n; r += blockDim.x) any |= artificial_row(t, r) ? 1 : 0;
    return QPN_SYNC_OR(any) != 0;
}

// ---- ratio test (avi_scratch.jl:65-77); every exit ends with a barrier ---------------------------
__device__ __forceinline__ double big_ratio(const BigTab& t, const double* dc, int r, double sigma) {
    const double d = sigma * dc[r];
    double lo, up;
    var_bounds(t, t.rowvar()[r], lo, up);
    if (d > D_TOL && lo > -QPN_INF) return fmax((t.beta()[r] - lo) / d, 0.0);
    if (d < -D_TOL && up < QPN_INF) return fmax((up - t.beta()[r]) / (-d), 0.0);
    return QPN_INF;
}
__device__ __noinline__ double ratio_test(BigTab& t, int c, double sigma, int& rho, int& which) {
    big_col(t, c);
    const int n = t.n;
    const double* dc = t.dcol();
    double rmin = QPN_INF;
    for (int r = threadIdx.x; r < n; r += blockDim.x) rmin = fmin(rmin, big_ratio(t, dc, r, sigma));
    const double theta = block_min(t.v, rmin);
    rho = -1; which = 0;
    if (theta == QPN_INF) { QPN_SYNC(); return QPN_INF; }
    const double cut = theta + TIE_TOL * (1.0 + theta);
    // among ties: t first (so the path terminates), then largest |d|, then lowest row
    double key = 0.0; int idx = -1;
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
        if (!(big_ratio(t, dc, r, sigma) <= cut)) continue;
        const double k = (t.rowvar()[r] == 2 * n) ? QPN_INF : fabs(dc[r]);
        if (idx < 0 || k > key) { key = k; idx = r; }
    }
    block_argmax_idx(t.v, idx >= 0, key, idx);
    rho = idx;
    BIGCHK(rho >= 0 && rho < n, "ratio_test: no tie winner", rho, c);
    QPN_SYNC();
    if (threadIdx.x == 0) t.v.red_d()[32] = big_ratio(t, dc, rho, sigma);
    QPN_SYNC();
    const double th = t.v.red_d()[32];
    which = (sigma * dc[rho] > 0.0) ? -1 : +1;
    return th;
}

__device__ __forceinline__ void move(BigTab& t, int c, double sigma, double theta) {
    if (theta == 0.0) return;
    big_col(t, c);
    const double* dc = t.dcol();
    for (int r = threadIdx.x; r < t.n; r += blockDim.x) {
        const double ci = dc[r];
        if (ci != 0.0) t.beta()[r] = fma(-(sigma * theta), ci, t.beta()[r]);
    }
    if (threadIdx.x == 0) t.nbval()[c] = fma(sigma, theta, t.nbval()[c]);
    QPN_SYNC();
}

// T[:, t] = B^-1 r from the slack columns (see avi_pivot.cuh).  Ends with a barrier.
__device__ __noinline__ void recompute_tcol(BigTab& t) {
    big_flush(t);
    const int n = t.n, ldr = t.ldr;
    const int tc = t.colof()[2 * n];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double* row = t.Tg + (size_t)i * ldr;
        double acc = 0.0;
        for (int k = 0; k < n; ++k) {
            const int ck = t.colof()[n + k];
            const double pik = ck >= 0 ? -row[ck] : (t.rowof()[n + k] == i ? -1.0 : 0.0);
            if (pik != 0.0) acc = fma(pik, t.rr()[k], acc);
        }
        row[tc] = acc;
    }
    t.cc = -1;
    QPN_SYNC();
}

// Drop every dead column (slack of a free variable) from the live range at once.
__device__ __noinline__ void compact_dead(BigTab& t) {
    big_flush(t);
    const int n = t.n, ldr = t.ldr;
    int* map = reinterpret_cast<int*>(t.prow());
    if (threadIdx.x == 0) {
        int nl = 0;
        for (int j = 0; j < t.ncol; ++j) {
            const int v = t.colvar()[j];
            if (v >= n && v < 2 * n && is_free_var(t, v - n)) t.colof()[v] = -1;
            else map[nl++] = j;
        }
        for (int d = 0; d < nl; ++d) {
            const int sidx = map[d];
            if (sidx != d) { const int v = t.colvar()[sidx]; t.colvar()[d] = v; t.nbval()[d] = t.nbval()[sidx]; t.colof()[v] = d; }
        }
        t.v.red_i()[32] = nl;
    }
    QPN_SYNC();
    const int nl = t.v.red_i()[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = w; i < n; i += nw) {
        double* row = t.Tg + (size_t)i * ldr;
        for (int d0 = 0; d0 < nl; d0 += 32) {           // map[d] >= d: a chunk only reads at or beyond itself
            const int d = d0 + lane;
            double val = 0.0;
            if (d < nl) val = row[map[d]];
            __syncwarp();
            if (d < nl) row[d] = val;
            __syncwarp();
        }
        if (lane == 0 && (nl & 1)) row[nl] = 0.0;
    }
    t.ncol = nl; t.cc = -1;
    QPN_SYNC();
}

// frozen_values (avi_pivot.cuh) for the global-memory engine: the same sum, with the frozen rows read from the plan
// when the instance never copied them (t.T0f); their homotopy entry is (B^-1 r)_i, rebuilt here exactly as the start
// of a solve builds it for the swept rows.  Ends with a barrier.
__device__ __noinline__ void frozen_values(BigTab& t, double* dx) {
    const int n = t.n, nc0 = t.ncol0, ldr = t.ldr;
    for (int j = threadIdx.x; j < nc0; j += blockDim.x) {
        const int v = t.colvar0()[j];
        const int r = t.rowof()[v], c = t.colof()[v];
        dx[j] = (r >= 0 ? t.beta()[r] : c >= 0 ? t.nbval()[c] : t.nbval0()[j]) - t.nbval0()[j];
    }
    QPN_SYNC();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (!frozen_row(t, i)) continue;
        double acc = t.beta()[i];
        if (t.T0f && i >= t.nact) {
            const double* row = t.T0f + (size_t)i * ldr;
            double bir = 0.0;
            for (int k = 0; k < n; ++k) {
                const double pik = t.PTf[(size_t)k * n + i];
                if (pik != 0.0) bir = fma(pik, t.rr()[k], bir);
            }
            for (int j = 0; j < nc0; ++j) acc = fma(-(j == t.tcol0 ? bir : row[j]), dx[j], acc);
        } else {
            const double* row = t.Tg + (size_t)i * ldr;
            for (int j = 0; j < nc0; ++j) acc = fma(-row[j], dx[j], acc);
        }
        t.beta()[i] = acc;
    }
    QPN_SYNC();
}

// Runs crash + repair + path following on a started big tableau; z and the basis codes go to
// zs / code (shared or global, n entries each; code may be null).  Ends with a barrier.
__device__ __noinline__ int avi_pivot_run_big(BigTab& t, int max_pivots, bool from_plan, double* zs, int8_t* code) {
    crash(t, from_plan);
    repair(t);
    const int st = lemke(t, max_pivots);
    t.npend = 0;                                          // z lives in beta / nbval: the queued sweeps are never needed
    QPN_SYNC();
    frozen_values(t, t.prow());                           // prow is free between pivots (ldrmax >= ncol0 doubles)
    for (int i = threadIdx.x; i < t.n; i += blockDim.x) {
        const int r = t.rowof()[i];
        zs[i] = r >= 0 ? t.beta()[r] : t.nbval()[t.colof()[i]];
        if (code) {
            const int8_t s = t.zst()[i];
            code[i] = (t.l()[i] == t.u()[i]) ? 4 : (r >= 0 || s == FLOATING) ? 2 : (s == AT_L ? 1 : 3);
        }
    }
    QPN_SYNC();
    return st;
}

}  // namespace qpn
