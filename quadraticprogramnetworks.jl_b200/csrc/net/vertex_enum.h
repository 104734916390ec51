// Vertices of a node's multiplier polytope at the current primal point -- the `get_verts(simplify(poly_slice(piece,
// slice_recipe)))` of `expand` (/root/reference/src/avi_solutions.jl:252-255, src/sets.jl:439-453).
//
// For the single-node GAVI of process_solution_graph (avi.jl:447-475), slicing a local piece at the primal part of z
// and at w leaves a polyhedron in the multipliers alone: stationarity  A_d' lam = qt  (qt = Q_d x + q_d), a sign per
// multiplier whose row is active, lam_i = 0 for every other row.  The piece of the recipe with the fewest "inactive"
// choices slices to the whole polytope
//     Lambda(x) = { lam : A_d' lam = qt,  lam_i >= 0 (row at its lower bound), <= 0 (upper), free (both), 0 (inactive) },
// every other admissible recipe to a face of it, so the vertices `collect` can ever queue are the vertices of Lambda(x).
// The reference enumerates them by double description (Polyhedra.jl); here the rows fixed at zero are eliminated and
// the basic solutions of the remaining a-column system are enumerated directly: a <= 12 columns, rank r, C(a, r) bases
// in lexicographic order, each solved by Gaussian elimination with partial pivoting (every multiply-add an explicit fma,
// so the device and the host build round alike).  Compiled for the device
// (net_verify_kernel), for the host state machine's checker build, and restated in the Python mirror (solgraph.py).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define QPN_VE_HD __host__ __device__
#else
#define QPN_VE_HD
#endif

#define QPN_VE_MAXA 12       // active rows handled (more: no exploration for that point)
#define QPN_VE_MAXND 8       // decision variables handled (larger nodes: no exploration)
#define QPN_VE_MAXV 15       // vertices returned at most

// A: m x nv column-major, dec: nd indices into x.  ax = A x, qt = Q_d x + q_d, lam = the multipliers verify_solution returned.
// Out: idxA[0..a) the active rows, V[k * QPN_VE_MAXA + j] = multiplier of row idxA[j] at new vertex k.
// Returns the number of NEW vertices (different from lam itself at 5 digits), at most max_new.
QPN_VE_HD inline int qpn_multiplier_vertices(int nd, int m, int nv, const double* A, const int* dec, const double* l, const double* u,
                                             const double* ax, const double* qt, const double* lam, int max_new, int* idxA, int* a_out,
                                             double* V) {
    *a_out = 0;
    if (nd > QPN_VE_MAXND || max_new <= 0) return 0;
    if (max_new > QPN_VE_MAXV) max_new = QPN_VE_MAXV;
    int a = 0;
    int sgn[QPN_VE_MAXA];
    for (int i = 0; i < m; ++i) {
        const bool lo = fabs(ax[i] - l[i]) <= 1e-6, up = fabs(ax[i] - u[i]) <= 1e-6;      // infinite bounds never match
        if (lo || up) {
            if (a == QPN_VE_MAXA) return 0;
            idxA[a] = i; sgn[a] = (lo && up) ? 0 : (lo ? 1 : -1);
            ++a;
        } else if (fabs(lam[i]) > 1e-6) {
            return 0;                            // the point is in no piece at the slice's tolerance: expand yields no vertices
        }
    }
    *a_out = a;
    if (a == 0) return 0;
    // G = A_d' restricted to the active rows (nd x a); rank and an independent row set by elimination with full pivoting
    double G[QPN_VE_MAXND * QPN_VE_MAXA], W[QPN_VE_MAXND * QPN_VE_MAXA];
    for (int e = 0; e < nd; ++e)
        for (int k = 0; k < a; ++k) { G[e * QPN_VE_MAXA + k] = A[(size_t)dec[e] * m + idxA[k]]; W[e * QPN_VE_MAXA + k] = G[e * QPN_VE_MAXA + k]; }
    for (int e = 0; e < nd; ++e) {               // the point itself must satisfy the slice's equalities at the piece tolerance
        double s = 0.0;
        for (int k = 0; k < a; ++k) s = fma(G[e * QPN_VE_MAXA + k], lam[idxA[k]], s);
        if (fabs(s - qt[e]) > 1e-6) return 0;
    }
    int rows[QPN_VE_MAXND], rowperm[QPN_VE_MAXND], colperm[QPN_VE_MAXA];
    for (int e = 0; e < nd; ++e) rowperm[e] = e;
    for (int k = 0; k < a; ++k) colperm[k] = k;
    int r = 0;
    const int lim = nd < a ? nd : a;
    for (; r < lim; ++r) {
        int pe = -1, pk = -1;
        double best = 1e-9;
        for (int e = r; e < nd; ++e)
            for (int k = r; k < a; ++k) {
                const double v = fabs(W[rowperm[e] * QPN_VE_MAXA + colperm[k]]);
                if (v > best) { best = v; pe = e; pk = k; }
            }
        if (pe < 0) break;
        int t = rowperm[r]; rowperm[r] = rowperm[pe]; rowperm[pe] = t;
        t = colperm[r]; colperm[r] = colperm[pk]; colperm[pk] = t;
        const double piv = W[rowperm[r] * QPN_VE_MAXA + colperm[r]];
        for (int e = r + 1; e < nd; ++e) {
            const double f = W[rowperm[e] * QPN_VE_MAXA + colperm[r]] / piv;
            if (f == 0.0) continue;
            for (int k = r; k < a; ++k)
                W[rowperm[e] * QPN_VE_MAXA + colperm[k]] = fma(-f, W[rowperm[r] * QPN_VE_MAXA + colperm[k]], W[rowperm[e] * QPN_VE_MAXA + colperm[k]]);
        }
    }
    if (a <= r) return 0;                        // independent columns: the polytope is the point lam itself
    for (int e = 0; e < r; ++e) rows[e] = rowperm[e];
    // the r equations in ascending order (the order only matters for reproducibility of the elimination below)
    for (int i = 1; i < r; ++i) { const int t = rows[i]; int j = i - 1; while (j >= 0 && rows[j] > t) { rows[j + 1] = rows[j]; --j; } rows[j + 1] = t; }
    // free multipliers (equality rows) must be basic: a polyhedron with a line among them has no vertices
    int nfree = 0;
    for (int k = 0; k < a; ++k) nfree += (sgn[k] == 0);
    if (nfree > r) return 0;
    int nv_found = 0;
    int comb[QPN_VE_MAXND];
    for (int i = 0; i < r; ++i) comb[i] = i;
    long guard = 0;
    while (true) {
        if (++guard > 20000) break;
        bool has_free = true;
        for (int k = 0; k < a && has_free; ++k)
            if (sgn[k] == 0) { bool in = false; for (int i = 0; i < r; ++i) in |= (comb[i] == k); has_free = in; }
        if (has_free) {
            // solve G[rows, comb] y = qt[rows]
            double Mx[QPN_VE_MAXND * QPN_VE_MAXND], y[QPN_VE_MAXND];
            for (int i = 0; i < r; ++i) { for (int j = 0; j < r; ++j) Mx[i * QPN_VE_MAXND + j] = G[rows[i] * QPN_VE_MAXA + comb[j]]; y[i] = qt[rows[i]]; }
            bool ok = true;
            for (int c = 0; c < r && ok; ++c) {
                int p = c;
                for (int i = c + 1; i < r; ++i) if (fabs(Mx[i * QPN_VE_MAXND + c]) > fabs(Mx[p * QPN_VE_MAXND + c])) p = i;
                if (fabs(Mx[p * QPN_VE_MAXND + c]) < 1e-9) { ok = false; break; }
                if (p != c) {
                    for (int j = 0; j < r; ++j) { const double t = Mx[p * QPN_VE_MAXND + j]; Mx[p * QPN_VE_MAXND + j] = Mx[c * QPN_VE_MAXND + j]; Mx[c * QPN_VE_MAXND + j] = t; }
                    const double t = y[p]; y[p] = y[c]; y[c] = t;
                }
                for (int i = c + 1; i < r; ++i) {
                    const double f = Mx[i * QPN_VE_MAXND + c] / Mx[c * QPN_VE_MAXND + c];
                    if (f == 0.0) continue;
                    for (int j = c; j < r; ++j) Mx[i * QPN_VE_MAXND + j] = fma(-f, Mx[c * QPN_VE_MAXND + j], Mx[i * QPN_VE_MAXND + j]);
                    y[i] = fma(-f, y[c], y[i]);
                }
            }
            if (ok) {
                for (int i = r - 1; i >= 0; --i) {
                    double s = y[i];
                    for (int j = i + 1; j < r; ++j) s = fma(-Mx[i * QPN_VE_MAXND + j], y[j], s);
                    y[i] = s / Mx[i * QPN_VE_MAXND + i];
                }
                double cand[QPN_VE_MAXA];
                for (int k = 0; k < a; ++k) cand[k] = 0.0;
                for (int i = 0; i < r; ++i) cand[comb[i]] = y[i];
                for (int k = 0; k < a && ok; ++k) if (sgn[k] != 0 && sgn[k] * cand[k] < -1e-6) ok = false;
                for (int e = 0; e < nd && ok; ++e) {      // every stationarity equation, not only the r chosen ones
                    double s = 0.0;
                    for (int k = 0; k < a; ++k) s = fma(G[e * QPN_VE_MAXA + k], cand[k], s);
                    if (fabs(s - qt[e]) > 1e-6) ok = false;
                }
                if (ok) {                                 // QuantizedVector (avi_solutions.jl:23-32): equal at 5 digits = the same vertex
                    bool same = true;
                    for (int k = 0; k < a; ++k) same &= (rint(cand[k] * 1e5) == rint(lam[idxA[k]] * 1e5));
                    bool dup = same;
                    for (int q = 0; q < nv_found && !dup; ++q) {
                        bool eq = true;
                        for (int k = 0; k < a; ++k) eq &= (rint(cand[k] * 1e5) == rint(V[q * QPN_VE_MAXA + k] * 1e5));
                        dup = eq;
                    }
                    if (!dup) {
                        for (int k = 0; k < a; ++k) V[nv_found * QPN_VE_MAXA + k] = cand[k];
                        if (++nv_found == max_new) return nv_found;
                    }
                }
            }
        }
        // next combination of r out of a, lexicographic
        int i = r - 1;
        while (i >= 0 && comb[i] == a - r + i) --i;
        if (i < 0) break;
        ++comb[i];
        for (int j = i + 1; j < r; ++j) comb[j] = comb[j - 1] + 1;
    }
    return nv_found;
}
