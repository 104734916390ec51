// Warp-cooperative form of net/vertex_enum.h (the `get_verts` of expand, /root/reference/src/avi_solutions.jl:252-255):
// the same decisions and the same arithmetic per candidate basis as qpn_multiplier_vertices -- the host build and the
// oracle keep the serial routine, the parity tests compare the two bit for bit -- but spread over the 32 lanes of one warp:
//   * the active rows by ballot, the matrix of the slice (G) and its elimination copy (W) in shared memory;
//   * the rank-revealing elimination with full pivoting as a warp arg-max (ties to the first entry in row-major order,
//     as the serial scan takes them) and one entry per lane in the update;
//   * the C(a, r) candidate bases QPN_VE_LANES at a time: lane t unranks the (base + t)-th combination in lexicographic order and
//     solves its r x r system in its own stretch of shared memory; lane 0 then appends the admissible candidates in
//     combination order, so the vertices come out in the serial order and the same ones are dropped as duplicates.
#pragma once
#include "net/vertex_enum.h"

namespace qpn {

// doubles of shared memory the routine needs behind V (QPN_VE_MAXV x QPN_VE_MAXA) for nodes with at most nd decision variables
#define QPN_VE_LANES 8      // candidate bases solved at a time (each needs its own stretch of shared memory)
__host__ __device__ __forceinline__ int ve_lane_stride(int nd) { return (nd * nd + nd + QPN_VE_MAXA) | 1; }
__host__ __device__ __forceinline__ size_t ve_scratch_bytes(int nd) {
    if (nd > QPN_VE_MAXND) nd = 0;
    // V, G, W, fE | per-lane Mx, y, cand | idxA, sgn, rowperm, colperm, rows, hdr
    return 8 * ((size_t)QPN_VE_MAXV * QPN_VE_MAXA + 2 * QPN_VE_MAXND * QPN_VE_MAXA + QPN_VE_MAXND + QPN_VE_LANES * (size_t)ve_lane_stride(nd)) +
           4 * (3 * QPN_VE_MAXA + 2 * QPN_VE_MAXND + 8);
}

__device__ __forceinline__ int ve_binom(int n, int k) {
    if (k < 0 || k > n) return 0;
    if (k > n - k) k = n - k;
    int c = 1;
    for (int j = 1; j <= k; ++j) c = c * (n - k + j) / j;       // exact: c holds C(n - k + j, j) after step j
    return c;
}

struct VeSmem {
    double *V, *G, *W, *fE, *lanes;
    int *idxA, *sgn, *rowperm, *colperm, *rows, *hdr;
    int stride;
};
__device__ __forceinline__ VeSmem ve_carve(double* base, int nd) {
    VeSmem s;
    s.V = base;
    s.G = s.V + QPN_VE_MAXV * QPN_VE_MAXA;
    s.W = s.G + QPN_VE_MAXND * QPN_VE_MAXA;
    s.fE = s.W + QPN_VE_MAXND * QPN_VE_MAXA;
    s.lanes = s.fE + QPN_VE_MAXND;
    s.stride = ve_lane_stride(nd > QPN_VE_MAXND ? 0 : nd);
    s.idxA = reinterpret_cast<int*>(s.lanes + QPN_VE_LANES * s.stride);
    s.sgn = s.idxA + QPN_VE_MAXA;
    s.rowperm = s.sgn + QPN_VE_MAXA;
    s.colperm = s.rowperm + QPN_VE_MAXND;
    s.rows = s.colperm + QPN_VE_MAXA;
    s.hdr = s.rows + QPN_VE_MAXND;              // [0] vertices, [1] active rows
    return s;
}

// Called by the 32 lanes of ONE warp (converged).  Returns the number of new vertices (uniform over the warp); the active
// rows are in s.idxA[0 .. *a_out), vertex k in s.V[k * QPN_VE_MAXA + j].
__device__ __forceinline__ int multiplier_vertices_warp(const VeSmem& s, int nd, int m, const double* __restrict__ A, const int* __restrict__ dec,
                                                        const double* __restrict__ l, const double* __restrict__ u, const double* ax, const double* qt,
                                                        const double* lam, int max_new, int* a_out) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    *a_out = 0;
    if (nd > QPN_VE_MAXND || max_new <= 0) return 0;
    if (max_new > QPN_VE_MAXV) max_new = QPN_VE_MAXV;
    // ---- active rows (serial lines 39-48: any inactive row with a multiplier, or more than MAXA active rows: nothing) ----
    int a = 0, fail = 0;
    #pragma unroll 1
    for (int base = 0; base < m; base += 32) {
        const int i = base + lane;
        bool lo = false, up = false, bad = false;
        if (i < m) {
            lo = fabs(ax[i] - l[i]) <= 1e-6; up = fabs(ax[i] - u[i]) <= 1e-6;
            bad = !(lo || up) && fabs(lam[i]) > 1e-6;
        }
        const unsigned act = __ballot_sync(FULL, lo || up);
        if (__any_sync(FULL, bad)) fail = 1;
        const int pos = a + __popc(act & ((1u << lane) - 1u));
        if ((lo || up) && pos < QPN_VE_MAXA) { s.idxA[pos] = i; s.sgn[pos] = (lo && up) ? 0 : (lo ? 1 : -1); }
        a += __popc(act);
    }
    if (fail || a > QPN_VE_MAXA) return 0;
    *a_out = a;
    if (a == 0) return 0;
    __syncwarp();
    // ---- G = A_d' on the active rows (nd x a), W its elimination copy ------------------------------------------------------
    #pragma unroll 1
    for (int idx = lane; idx < nd * a; idx += 32) {
        const int e = idx / a, k = idx - e * a;
        const double v = A[(size_t)dec[e] * m + s.idxA[k]];
        s.G[e * QPN_VE_MAXA + k] = v; s.W[e * QPN_VE_MAXA + k] = v;
    }
    if (lane < nd) s.rowperm[lane] = lane;
    if (lane < a) s.colperm[lane] = lane;
    __syncwarp();
    {   // the point itself must satisfy the slice's equalities at the piece tolerance
        bool bad = false;
        #pragma unroll 1
        for (int e = lane; e < nd; e += 32) {
            double sum = 0.0;
            #pragma unroll 1
            for (int k = 0; k < a; ++k) sum = fma(s.G[e * QPN_VE_MAXA + k], lam[s.idxA[k]], sum);
            if (fabs(sum - qt[e]) > 1e-6) bad = true;
        }
        if (__any_sync(FULL, bad)) return 0;
    }
    // ---- rank and an independent row set by elimination with full pivoting -----------------------------------------------
    int r = 0;
    const int lim = nd < a ? nd : a;
    #pragma unroll 1
    for (; r < lim; ++r) {
        const int wr = a - r, cnt = (nd - r) * wr;
        double best = 1e-9;
        int bidx = -1;
        #pragma unroll 1
        for (int idx = lane; idx < cnt; idx += 32) {
            const int e = r + idx / wr, k = r + idx % wr;
            const double v = fabs(s.W[s.rowperm[e] * QPN_VE_MAXA + s.colperm[k]]);
            if (v > best) { best = v; bidx = idx; }
        }
        {
            // warp arg-max by REDUX: the bit pattern of a non-negative double orders like an unsigned integer (two 32-bit
            // maxima), then the lowest row-major index among the lanes that hold the maximum
            const unsigned hi = bidx >= 0 ? (unsigned)__double2hiint(best) : 0u, lo = bidx >= 0 ? (unsigned)__double2loint(best) : 0u;
            const unsigned mh = __reduce_max_sync(FULL, hi);
            const unsigned ml = __reduce_max_sync(FULL, hi == mh ? lo : 0u);
            const bool mine = bidx >= 0 && hi == mh && lo == ml;
            const unsigned widx = __reduce_min_sync(FULL, mine ? (unsigned)bidx : 0xffffffffu);
            bidx = widx == 0xffffffffu ? -1 : (int)widx;
        }
        if (bidx < 0) break;
        __syncwarp();
        if (lane == 0) {
            const int pe = r + bidx / wr, pk = r + bidx % wr;
            int t = s.rowperm[r]; s.rowperm[r] = s.rowperm[pe]; s.rowperm[pe] = t;
            t = s.colperm[r]; s.colperm[r] = s.colperm[pk]; s.colperm[pk] = t;
        }
        __syncwarp();
        const double piv = s.W[s.rowperm[r] * QPN_VE_MAXA + s.colperm[r]];
        #pragma unroll 1
        for (int e = r + 1 + lane; e < nd; e += 32) s.fE[e] = s.W[s.rowperm[e] * QPN_VE_MAXA + s.colperm[r]] / piv;
        __syncwarp();
        const int ucnt = (nd - r - 1) * wr;
        #pragma unroll 1
        for (int idx = lane; idx < ucnt; idx += 32) {
            const int e = r + 1 + idx / wr, k = r + idx % wr;
            const double f = s.fE[e];
            if (f == 0.0) continue;
            double* w = s.W + s.rowperm[e] * QPN_VE_MAXA + s.colperm[k];
            *w = fma(-f, s.W[s.rowperm[r] * QPN_VE_MAXA + s.colperm[k]], *w);
        }
        __syncwarp();
    }
    if (a <= r) return 0;                        // independent columns: the polytope is the point lam itself
    if (lane == 0) {                             // the r equations in ascending order
        #pragma unroll 1
        for (int e = 0; e < r; ++e) s.rows[e] = s.rowperm[e];
        #pragma unroll 1
        for (int i = 1; i < r; ++i) { const int t = s.rows[i]; int j = i - 1; while (j >= 0 && s.rows[j] > t) { s.rows[j + 1] = s.rows[j]; --j; } s.rows[j + 1] = t; }
    }
    __syncwarp();
    // free multipliers (equality rows) must be basic: a polyhedron with a line among them has no vertices
    unsigned freemask = 0;
    #pragma unroll 1
    for (int k = 0; k < a; ++k) if (s.sgn[k] == 0) freemask |= 1u << k;
    if (__popc(freemask) > r) return 0;
    // ---- the candidate bases, QPN_VE_LANES at a time ------------------------------------------------------------------------------
    const int total = ve_binom(a, r);
    double* Mx = s.lanes + lane * s.stride;      // r x r, row stride r
    double* y = Mx + nd * nd;
    double* cand = y + nd;
    int nv_found = 0;
    #pragma unroll 1
    for (int base = 0; base < total && nv_found < max_new; base += QPN_VE_LANES) {
        int c = base + lane;
        bool ok = c < total && lane < QPN_VE_LANES;
        unsigned comb = 0, combmask = 0;         // comb[i] in nibble i
        if (ok) {
            int x = 0;
            #pragma unroll 1
            for (int i = 0; i < r; ++i) {
                #pragma unroll 1
                while (true) {
                    const int skip = ve_binom(a - x - 1, r - i - 1);
                    if (skip <= c) { c -= skip; ++x; } else break;
                }
                comb |= (unsigned)x << (4 * i); combmask |= 1u << x;
                ++x;
            }
            ok = (freemask & ~combmask) == 0;
        }
        if (ok) {
            // solve G[rows, comb] y = qt[rows] (Gaussian elimination with partial pivoting; every multiply-add an fma)
            #pragma unroll 1
            for (int i = 0; i < r; ++i) {
                #pragma unroll 1
                for (int j = 0; j < r; ++j) Mx[i * r + j] = s.G[s.rows[i] * QPN_VE_MAXA + ((comb >> (4 * j)) & 15)];
                y[i] = qt[s.rows[i]];
            }
            #pragma unroll 1
            for (int cc = 0; cc < r && ok; ++cc) {
                int p = cc;
                #pragma unroll 1
                for (int i = cc + 1; i < r; ++i) if (fabs(Mx[i * r + cc]) > fabs(Mx[p * r + cc])) p = i;
                if (fabs(Mx[p * r + cc]) < 1e-9) { ok = false; break; }
                if (p != cc) {
                    #pragma unroll 1
                    for (int j = 0; j < r; ++j) { const double t = Mx[p * r + j]; Mx[p * r + j] = Mx[cc * r + j]; Mx[cc * r + j] = t; }
                    const double t = y[p]; y[p] = y[cc]; y[cc] = t;
                }
                #pragma unroll 1
                for (int i = cc + 1; i < r; ++i) {
                    const double f = Mx[i * r + cc] / Mx[cc * r + cc];
                    if (f == 0.0) continue;
                    #pragma unroll 1
                    for (int j = cc; j < r; ++j) Mx[i * r + j] = fma(-f, Mx[cc * r + j], Mx[i * r + j]);
                    y[i] = fma(-f, y[cc], y[i]);
                }
            }
        }
        if (ok) {
            #pragma unroll 1
            for (int i = r - 1; i >= 0; --i) {
                double sum = y[i];
                #pragma unroll 1
                for (int j = i + 1; j < r; ++j) sum = fma(-Mx[i * r + j], y[j], sum);
                y[i] = sum / Mx[i * r + i];
            }
            #pragma unroll 1
            for (int k = 0; k < a; ++k) cand[k] = 0.0;
            #pragma unroll 1
            for (int i = 0; i < r; ++i) cand[(comb >> (4 * i)) & 15] = y[i];
            #pragma unroll 1
            for (int k = 0; k < a && ok; ++k) if (s.sgn[k] != 0 && s.sgn[k] * cand[k] < -1e-6) ok = false;
            #pragma unroll 1
            for (int e = 0; e < nd && ok; ++e) {      // every stationarity equation, not only the r chosen ones
                double sum = 0.0;
                #pragma unroll 1
                for (int k = 0; k < a; ++k) sum = fma(s.G[e * QPN_VE_MAXA + k], cand[k], sum);
                if (fabs(sum - qt[e]) > 1e-6) ok = false;
            }
            if (ok) {                                 // QuantizedVector (avi_solutions.jl:23-32): equal to the point at 5 digits
                bool same = true;
                #pragma unroll 1
                for (int k = 0; k < a; ++k) same &= (rint(cand[k] * 1e5) == rint(lam[s.idxA[k]] * 1e5));
                if (same) ok = false;
            }
        }
        unsigned good = __ballot_sync(FULL, ok);
        if (lane == 0) {
            // in combination order: new at 5 digits against the vertices found so far
            #pragma unroll 1
            while (good && nv_found < max_new) {
                const int t = __ffs(good) - 1;
                good &= good - 1;
                const double* ct = s.lanes + t * s.stride + nd * nd + nd;
                bool dup = false;
                #pragma unroll 1
                for (int q = 0; q < nv_found && !dup; ++q) {
                    bool eq = true;
                    #pragma unroll 1
                    for (int k = 0; k < a; ++k) eq &= (rint(ct[k] * 1e5) == rint(s.V[q * QPN_VE_MAXA + k] * 1e5));
                    dup = eq;
                }
                if (!dup) {
                    #pragma unroll 1
                    for (int k = 0; k < a; ++k) s.V[nv_found * QPN_VE_MAXA + k] = ct[k];
                    ++nv_found;
                }
            }
        }
        nv_found = __shfl_sync(FULL, nv_found, 0);
        __syncwarp();
    }
    return nv_found;
}

}  // namespace qpn
