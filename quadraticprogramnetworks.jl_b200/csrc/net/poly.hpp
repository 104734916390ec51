// Polyhedral sets of the native network solver (host side, C++17).
//
// Restates the set operations the reference performs on the host around the numeric hot path
// (/root/reference/src/sets.jl): Slice normalisation (:76-89), BasicPoly with 5-digit slice equality
// (:104-112,123-125,141-146), simplify (:255-305), poly_slice (:532-542), exemplar / isempty (:591-655),
// issubset (:377-407), remove_subsets (:889-902), complement (:918-930), poly_intersect (:936-968), and the
// projection of sets.jl:501-523 (Polyhedra.jl double description there; elimination through equality rows,
// Fourier-Motzkin combination and LP redundancy removal here).  Every LP goes to the numeric backend as a GAVI
// solve (avi.jl:79-128), exactly as the Python host mirror does (polyhedra.py); the arithmetic below follows that
// mirror operation for operation (compile with -ffp-contract=off) so both produce the same rows bit for bit.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <unordered_set>
#include <utility>
#include <vector>

namespace qpnnet {

static const double INF = std::numeric_limits<double>::infinity();

// numpy.round(x, p): rint(x * 10^p) / 10^p (ties to even); infinities pass through.
static inline double round_dec(double x, double scale) {
    if (std::isinf(x)) return x;
    return std::nearbyint(x * scale) / scale;
}

struct Poly {
    int d = 0;
    std::vector<double> A;            // m x d, row-major
    std::vector<double> l, u;
    std::vector<uint8_t> rl, ru;      // 1 = strict '<'
    std::vector<std::string> keys;    // per row: the 5-digit slice key (sets.jl:104-112)
    int m() const { return (int)l.size(); }
    const double* row(int i) const { return A.data() + (size_t)i * d; }
    double* row(int i) { return A.data() + (size_t)i * d; }
};

// Rows as they are handed around before a Poly is built.
struct Rows {
    int d = 0;
    std::vector<double> A, l, u;
    std::vector<uint8_t> rl, ru;
    int m() const { return (int)l.size(); }
    void add(const double* a, double lo, double up, uint8_t sl = 0, uint8_t su = 0) {
        A.insert(A.end(), a, a + d);
        l.push_back(lo); u.push_back(up); rl.push_back(sl); ru.push_back(su);
    }
};

static inline std::string slice_key(const double* a, int d, double l, double u, uint8_t rl, uint8_t ru) {
    std::string k((size_t)(d + 4) * 8, '\0');
    double* out = reinterpret_cast<double*>(&k[0]);
    for (int j = 0; j < d; ++j) out[j] = round_dec(a[j], 1e5) + 0.0;
    out[d] = round_dec(l, 1e5) + 0.0;
    out[d + 1] = round_dec(u, 1e5) + 0.0;
    out[d + 2] = (double)rl;
    out[d + 3] = (double)ru;
    return k;
}

// model.Poly.__init__: every row through normalize_slice (sets.jl:76-89) when `normalize`, then equal slices stored
// once, in first-seen order.
static inline Poly make_poly(Rows r, bool normalize = true) {
    const int d = r.d, m = r.m();
    if (normalize) {
        for (int i = 0; i < m; ++i) {
            double* a = r.A.data() + (size_t)i * d;
            double ss = 0.0;
            for (int j = 0; j < d; ++j) { if (std::fabs(a[j]) <= 1e-8) a[j] = 0.0; ss += a[j] * a[j]; }
            if (std::sqrt(ss) <= 1e-8) { for (int j = 0; j < d; ++j) a[j] = 0.0; continue; }
            int first = 0;
            while (first < d && a[first] == 0.0) ++first;
            const double lead = a[first], n = std::fabs(lead);
            if (lead >= 0) {
                for (int j = 0; j < d; ++j) a[j] = a[j] / n;
                r.l[i] = r.l[i] / n; r.u[i] = r.u[i] / n;
            } else {
                for (int j = 0; j < d; ++j) a[j] = -(a[j] / n);
                const double lo = -(r.u[i] / n), up = -(r.l[i] / n);
                r.l[i] = lo; r.u[i] = up;
                std::swap(r.rl[i], r.ru[i]);
            }
        }
    }
    Poly P;
    P.d = d;
    std::unordered_set<std::string> seen;
    for (int i = 0; i < m; ++i) {
        const double* a = r.A.data() + (size_t)i * d;
        std::string k = slice_key(a, d, r.l[i], r.u[i], r.rl[i], r.ru[i]);
        if (!seen.insert(k).second) continue;
        P.A.insert(P.A.end(), a, a + d);
        P.l.push_back(r.l[i]); P.u.push_back(r.u[i]); P.rl.push_back(r.rl[i]); P.ru.push_back(r.ru[i]);
        P.keys.push_back(std::move(k));
    }
    return P;
}

static inline Rows rows_of(const Poly& P) {
    Rows r;
    r.d = P.d; r.A = P.A; r.l = P.l; r.u = P.u; r.rl = P.rl; r.ru = P.ru;
    return r;
}

// Exact content (row order and every bit): the key of memos whose value depends on the rows themselves.
static inline std::string exact_key(const Poly& P) {
    std::string k;
    k.reserve(16 + P.A.size() * 8 + P.l.size() * 18);
    int32_t hdr[2] = {P.d, P.m()};
    k.append(reinterpret_cast<const char*>(hdr), sizeof hdr);
    k.append(reinterpret_cast<const char*>(P.A.data()), P.A.size() * 8);
    k.append(reinterpret_cast<const char*>(P.l.data()), P.l.size() * 8);
    k.append(reinterpret_cast<const char*>(P.u.data()), P.u.size() * 8);
    k.append(reinterpret_cast<const char*>(P.rl.data()), P.rl.size());
    k.append(reinterpret_cast<const char*>(P.ru.data()), P.ru.size());
    return k;
}

// The unordered set of slice keys as one string: equal iff the polyhedra are equal in the reference's sense
// (sets.jl:141-146).
static inline std::string set_key(const Poly& P) {
    std::vector<const std::string*> ks;
    for (const auto& k : P.keys) ks.push_back(&k);
    std::sort(ks.begin(), ks.end(), [](const std::string* a, const std::string* b) { return *a < *b; });
    std::string out;
    int32_t d = P.d;
    out.append(reinterpret_cast<const char*>(&d), 4);
    for (auto* k : ks) out += *k;
    return out;
}

// sets.jl:820-853 (host form; the batched form runs on the numeric backend)
static inline bool contains(const Poly& P, const double* x, double tol = 1e-6, bool closed = false) {
    for (int i = 0; i < P.m(); ++i) {
        const double* a = P.row(i);
        double ax = 0.0;
        for (int j = 0; j < P.d; ++j) ax += a[j] * x[j];
        const bool ls = P.rl[i] && !closed, us = P.ru[i] && !closed;
        const bool lo = ls ? (P.l[i] - tol < ax) : (P.l[i] - tol <= ax);
        const bool up = us ? (ax - tol < P.u[i]) : (ax - tol <= P.u[i]);
        if (!(lo && up)) return false;
    }
    return true;
}

// poly_intersect (sets.jl:936-968): the conjunction of all slices, rows of `a` first.
static inline Poly intersect(const Poly& a, const Poly& b) {
    Rows r = rows_of(a);
    r.A.insert(r.A.end(), b.A.begin(), b.A.end());
    r.l.insert(r.l.end(), b.l.begin(), b.l.end());
    r.u.insert(r.u.end(), b.u.begin(), b.u.end());
    r.rl.insert(r.rl.end(), b.rl.begin(), b.rl.end());
    r.ru.insert(r.ru.end(), b.ru.begin(), b.ru.end());
    return make_poly(std::move(r), false);
}

// sets.jl:918-930: one open half-space per finite bound.
static inline std::vector<Poly> complement(const Poly& P) {
    std::vector<Poly> out;
    for (int i = 0; i < P.m(); ++i) {
        if (!std::isinf(P.l[i])) {
            Rows r; r.d = P.d;
            r.add(P.row(i), -INF, P.l[i], 1, !P.rl[i]);
            out.push_back(make_poly(std::move(r), false));
        }
        if (!std::isinf(P.u[i])) {
            Rows r; r.d = P.d;
            r.add(P.row(i), P.u[i], INF, !P.ru[i], 1);
            out.push_back(make_poly(std::move(r), false));
        }
    }
    return out;
}

// sets.jl:255-305: merge slices with the same normal, keeping the tighter bounds.
static inline Poly simplify(const Poly& P, double tol = 1e-6) {
    const int m = P.m(), d = P.d;
    Rows out; out.d = d;
    if (m == 0) return make_poly(std::move(out));
    struct Kept { int row; double l, u; uint8_t rl, ru; };
    std::vector<Kept> keep;
    for (int i = 0; i < m; ++i) {
        const double* a = P.row(i);
        double l = P.l[i], u = P.u[i];
        uint8_t rl = P.rl[i], ru = P.ru[i];
        Kept* k = nullptr;
        for (auto& kk : keep) {
            const double* b = P.row(kk.row);
            double ss = 0.0;
            for (int j = 0; j < d; ++j) { const double e = a[j] - b[j]; ss += e * e; }
            if (std::sqrt(ss) <= tol) { k = &kk; break; }
        }
        if (k) {
            double nl, nu; uint8_t nrl, nru;
            if (k->l > l + tol) { nl = k->l; nrl = k->rl; }
            else if (l > k->l + tol) { nl = l; nrl = rl; }
            else { nl = (std::isinf(k->l) && std::isinf(l)) ? l : 0.5 * (k->l + l); nrl = k->rl ? 1 : rl; }
            if (k->u < u - tol) { nu = k->u; nru = k->ru; }
            else if (u < k->u - tol) { nu = u; nru = ru; }
            else { nu = (std::isinf(k->u) && std::isinf(u)) ? u : 0.5 * (k->u + u); nru = k->ru ? 1 : ru; }
            k->l = nl; k->u = nu; k->rl = nrl; k->ru = nru;
        } else {
            double ss = 0.0;
            for (int j = 0; j < d; ++j) ss += a[j] * a[j];
            if (std::sqrt(ss) > tol) keep.push_back({i, l, u, rl, ru});
        }
    }
    for (auto& k : keep) out.add(P.row(k.row), k.l, k.u, k.rl, k.ru);
    return make_poly(std::move(out));
}

// sets.jl:532-542: fix some coordinates (fixed[j] finite or NaN = free), drop them.
static inline Poly poly_slice(const Poly& P, const std::vector<double>& fixed) {
    Rows r;
    int keepn = 0;
    for (int j = 0; j < P.d; ++j) if (std::isnan(fixed[j])) ++keepn;
    r.d = keepn;
    std::vector<double> a(keepn > 0 ? keepn : 1);
    for (int i = 0; i < P.m(); ++i) {
        const double* row = P.row(i);
        double shift = 0.0;
        int c = 0;
        for (int j = 0; j < P.d; ++j) {
            if (std::isnan(fixed[j])) a[c++] = row[j];
            else shift += row[j] * fixed[j];
        }
        r.add(a.data(), P.l[i] - shift, P.u[i] - shift, P.rl[i], P.ru[i]);
    }
    return make_poly(std::move(r));
}

// A generalized AVI with dense column-major blocks (struct GAVI, avi.jl:29-39; the layout of qpn_gavi in
// include/qpn_cuda.h).
struct GaviData {
    int d1 = 0, d2 = 0, np = 0;
    std::vector<double> M, N, o, l1, u1, A, B, l2, u2;
};

// The numeric backend as the geometry sees it: one GAVI solve (presolve on), synchronous.
struct LPBackend {
    virtual ~LPBackend() {}
    // z0, z: d1 + d2 entries.  Returns the StatusCode (1 = SUCCESS).
    virtual int gavi_solve_one(const GaviData& g, const double* w, const double* z0, double* z) = 0;
    // The same GAVI for `batch` parameter vectors in one call (W: np x batch, Z0 / Z: (d1 + d2) x batch, column-major).
    // true: a batched call costs about what a single solve costs (a GPU): callers then prefer a few wide calls over many
    // narrow ones even when that solves LPs a serial scan would have skipped
    virtual bool wide_batches() const { return false; }
    virtual void gavi_solve_many(const GaviData& g, int batch, const double* W, const double* Z0, double* Z, int32_t* status) {
        const int dz = g.d1 + g.d2;
        for (int b = 0; b < batch; ++b) status[b] = gavi_solve_one(g, W + (size_t)b * g.np, Z0 + (size_t)b * dz, Z + (size_t)b * dz);
    }
};

struct LPResult { int status = 0; std::vector<double> x, lam; double obj = 0.0; };

// polyhedra.LPSolver.solve: min c'x s.t. l <= Ax <= u as the GAVI with M = [0 -A'], o = c.
static inline LPResult lp_solve(LPBackend& be, const std::vector<double>& c, const double* A /*m x n row-major*/, int m,
                                const double* l, const double* u, long* counter = nullptr) {
    const int n = (int)c.size();
    GaviData g;
    g.d1 = n; g.d2 = m; g.np = 0;
    const int dz = n + m;
    g.M.assign((size_t)n * dz, 0.0);
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) g.M[(size_t)(n + i) * n + j] = -A[(size_t)i * n + j];       // M[j, n+i] = -A[i, j]
    g.o = c;
    g.l1.assign(n, -INF); g.u1.assign(n, INF);
    g.A.assign((size_t)m * dz, 0.0);
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) g.A[(size_t)j * m + i] = A[(size_t)i * n + j];
    g.l2.assign(l, l + m); g.u2.assign(u, u + m);
    std::vector<double> z0(dz, 0.0), z(dz, 0.0);
    LPResult res;
    res.status = be.gavi_solve_one(g, nullptr, z0.data(), z.data());
    if (counter) ++*counter;
    res.x.assign(z.begin(), z.begin() + n);
    res.lam.assign(z.begin() + n, z.end());
    double obj = 0.0;
    for (int j = 0; j < n; ++j) obj += c[j] * res.x[j];
    res.obj = obj;
    return res;
}

static inline bool allclose(double a, double b, double tol) { return std::fabs(a - b) <= tol + tol * std::fabs(b); }

// Square dense solve, Gaussian elimination with partial pivoting; false on an exactly singular matrix.
static inline bool dense_solve(std::vector<double> A /*n x n row-major*/, std::vector<double> b, int n, std::vector<double>& x) {
    for (int k = 0; k < n; ++k) {
        int p = k;
        for (int i = k + 1; i < n; ++i) if (std::fabs(A[(size_t)i * n + k]) > std::fabs(A[(size_t)p * n + k])) p = i;
        if (A[(size_t)p * n + k] == 0.0) return false;
        if (p != k) { for (int j = 0; j < n; ++j) std::swap(A[(size_t)p * n + j], A[(size_t)k * n + j]); std::swap(b[p], b[k]); }
        for (int i = k + 1; i < n; ++i) {
            const double f = A[(size_t)i * n + k] / A[(size_t)k * n + k];
            if (f == 0.0) continue;
            for (int j = k; j < n; ++j) A[(size_t)i * n + j] -= f * A[(size_t)k * n + j];
            b[i] -= f * b[k];
        }
    }
    x.assign(n, 0.0);
    for (int i = n - 1; i >= 0; --i) {
        double s = b[i];
        for (int j = i + 1; j < n; ++j) s -= A[(size_t)i * n + j] * x[j];
        x[i] = s / A[(size_t)i * n + i];
    }
    return true;
}

// sets.jl:591-642.  Returns true iff the polyhedron is empty.
static inline bool exemplar_empty(const Poly& P, LPBackend& be, double tol, long* lp_count = nullptr) {
    const int n = P.m(), d = P.d;
    if (n == 0) return false;
    bool any_open = false, all_eq = true, any_inf_l = false;
    for (int i = 0; i < n; ++i) {
        if ((P.rl[i] && !std::isinf(P.l[i])) || (P.ru[i] && !std::isinf(P.u[i]))) any_open = true;
        if (!allclose(P.l[i], P.u[i], tol)) all_eq = false;
        if (std::isinf(P.l[i])) any_inf_l = true;
    }
    if (all_eq && !any_open && n == d && !any_inf_l) {
        std::vector<double> x;
        if (dense_solve(P.A, P.l, n, x)) {
            bool ok = true;
            for (int i = 0; i < n; ++i) {
                double ax = 0.0;
                for (int j = 0; j < d; ++j) ax += P.row(i)[j] * x[j];
                if (!allclose(ax, P.l[i], tol)) ok = false;
            }
            return !ok;
        }
    }
    // min eps  s.t.  A x + eps >= l,  -A x + eps >= -u   (rows with an infinite right-hand side constrain nothing)
    std::vector<double> AA, ll;
    std::vector<int> src;                   // row of the 2n-row system each kept row came from
    for (int half = 0; half < 2; ++half)
        for (int i = 0; i < n; ++i) {
            const double rhs = half == 0 ? P.l[i] : -P.u[i];
            if (std::isinf(rhs)) continue;
            for (int j = 0; j < d; ++j) AA.push_back(half == 0 ? P.row(i)[j] : -P.row(i)[j]);
            AA.push_back(1.0);
            ll.push_back(rhs);
            src.push_back(half * n + i);
        }
    const int mk = (int)ll.size();
    std::vector<double> c(d + 1, 0.0), uu(mk, INF);
    c[d] = 1.0;
    LPResult res = lp_solve(be, c, AA.data(), mk, ll.data(), uu.data(), lp_count);
    if (res.status != 1) return false;      // unbounded below: a whole cone of interior points
    const double eps = res.x[d];
    if (eps > tol) return true;
    if (eps > -tol) {
        for (int k = 0; k < mk; ++k) {
            if (!(std::fabs(res.lam[k]) > tol)) continue;
            const int s = src[k], i = s % n;
            if (s < n) { if (P.rl[i] && !std::isinf(P.l[i])) return true; }
            else { if (P.ru[i] && !std::isinf(P.u[i])) return true; }
        }
    }
    return false;
}

// sets.jl:377-407: for every finite bound of P2 minimise the bound's direction over P1.
static inline bool issubset(const Poly& P1, const Poly& P2, LPBackend& be, double tol = 1e-6, long* lp_count = nullptr) {
    std::vector<double> c(P2.d);
    for (int i = 0; i < P2.m(); ++i)
        for (int side = 0; side < 2; ++side) {
            const double bound = side == 0 ? P2.l[i] : P2.u[i], dirn = side == 0 ? 1.0 : -1.0;
            if (std::isinf(bound)) continue;
            if (P1.m() == 0) return false;
            for (int j = 0; j < P2.d; ++j) c[j] = dirn * P2.row(i)[j];
            LPResult res = lp_solve(be, c, P1.A.data(), P1.m(), P1.l.data(), P1.u.data(), lp_count);
            if (res.status != 1) return false;
            if (res.obj < dirn * bound - tol) return false;
        }
    return true;
}

// issubset(P1, P2) for several P2 at once (remove_subsets, sets.jl:889-902, asks for every pair of a list): the LPs of the
// bounds of all P2 share P1's rows, so they go out as batched GAVI solves with the cost vector as the parameter
// (M = [0 -A'], N = I, o = 0, w = +-a) instead of one launch per bound.  The reference stops at the first bound that
// fails; here every P2 still in the race contributes its next 1, 2, 4, ... bounds per call, so a pair that is no subset
// costs about as many LPs as the serial scan and the whole list a handful of launches.  out[k] = P1 subset of P2s[k].
static inline void issubset_many(const Poly& P1, const std::vector<const Poly*>& P2s, LPBackend& be, std::vector<char>& out,
                                 double tol = 1e-6, long* lp_count = nullptr, long* call_count = nullptr) {
    const size_t K = P2s.size();
    out.assign(K, 1);
    const int n = P1.d, m = P1.m();
    // the finite bounds of every P2 in the order of the serial scan
    struct Bound { int row; double dirn, rhs; };
    std::vector<std::vector<Bound>> bounds(K);
    for (size_t k = 0; k < K; ++k) {
        const Poly& P2 = *P2s[k];
        for (int i = 0; i < P2.m(); ++i)
            for (int side = 0; side < 2; ++side) {
                const double bound = side == 0 ? P2.l[i] : P2.u[i], dirn = side == 0 ? 1.0 : -1.0;
                if (std::isinf(bound)) continue;
                bounds[k].push_back({i, dirn, dirn * bound - tol});
            }
        if (m == 0 && !bounds[k].empty()) out[k] = 0;
    }
    if (m == 0) return;
    GaviData g;
    g.d1 = n; g.d2 = m; g.np = n;
    const int dz = n + m;
    g.M.assign((size_t)n * dz, 0.0);
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) g.M[(size_t)(n + i) * n + j] = -P1.A[(size_t)i * n + j];
    g.N.assign((size_t)n * n, 0.0);
    for (int j = 0; j < n; ++j) g.N[(size_t)j * n + j] = 1.0;
    g.o.assign(n, 0.0);
    g.l1.assign(n, -INF); g.u1.assign(n, INF);
    g.A.assign((size_t)m * dz, 0.0);
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) g.A[(size_t)j * m + i] = P1.A[(size_t)i * n + j];
    g.B.assign((size_t)m * n, 0.0);
    g.l2 = P1.l; g.u2 = P1.u;
    std::vector<size_t> next(K, 0);
    std::vector<double> W, Z0, Z;
    std::vector<int32_t> st;
    struct Item { int k; double rhs; };
    std::vector<Item> items;
    for (size_t chunk = 1;; chunk *= 2) {
        W.clear(); items.clear();
        for (size_t k = 0; k < K; ++k) {
            if (!out[k]) continue;
            const Poly& P2 = *P2s[k];
            for (size_t c = 0; c < chunk && next[k] < bounds[k].size(); ++c, ++next[k]) {
                const Bound& bd = bounds[k][next[k]];
                for (int j = 0; j < n; ++j) W.push_back(bd.dirn * P2.row(bd.row)[j]);
                items.push_back({(int)k, bd.rhs});
            }
        }
        if (items.empty()) return;
        const int B = (int)items.size();
        Z0.assign((size_t)dz * B, 0.0); Z.assign((size_t)dz * B, 0.0); st.assign(B, 0);
        be.gavi_solve_many(g, B, W.data(), Z0.data(), Z.data(), st.data());
        if (lp_count) *lp_count += B;
        if (call_count) ++*call_count;
        for (int b = 0; b < B; ++b) {
            if (st[b] != 1) { out[items[b].k] = 0; continue; }
            double obj = 0.0;
            for (int j = 0; j < n; ++j) obj += W[(size_t)b * n + j] * Z[(size_t)b * dz + j];
            if (obj < items[b].rhs) out[items[b].k] = 0;
        }
    }
}

// ---- projection (replaces sets.jl:501-523) ----------------------------------------------------------------------
struct Ineq { std::vector<double> a; double b; };

static inline void dedupe_rows(std::vector<Ineq>& rows, double tol = 1e-9) {
    std::vector<Ineq> out;
    std::unordered_set<std::string> seen;
    for (auto& r : rows) {
        double nrm = 0.0;
        for (double v : r.a) nrm = std::max(nrm, std::fabs(v));
        if (nrm <= tol) continue;           // 0 <= b rows carry nothing
        std::string key((r.a.size() + 1) * 8, '\0');
        double* kp = reinterpret_cast<double*>(&key[0]);
        for (size_t j = 0; j < r.a.size(); ++j) { r.a[j] = r.a[j] / nrm; kp[j] = round_dec(r.a[j], 1e9) + 0.0; }
        r.b = r.b / nrm;
        kp[r.a.size()] = round_dec(r.b, 1e9) + 0.0;
        if (!seen.insert(key).second) continue;
        out.push_back(std::move(r));
    }
    rows.swap(out);
}

// Drop every inequality implied by the others (max a'x over the rest <= b), one at a time in row order.
static inline void irredundant(const std::vector<Ineq>& eq, std::vector<Ineq>& ineq, LPBackend& be, int d, long* lp_count,
                               double tol = 1e-7) {
    size_t i = 0;
    while (i < ineq.size()) {
        const size_t mo = eq.size() + ineq.size() - 1;
        bool drop = false;
        if (mo > 0) {
            std::vector<double> A, l, u, c(d);
            A.reserve(mo * d);
            for (auto& e : eq) { A.insert(A.end(), e.a.begin(), e.a.end()); l.push_back(e.b); u.push_back(e.b); }
            for (size_t k = 0; k < ineq.size(); ++k) {
                if (k == i) continue;
                A.insert(A.end(), ineq[k].a.begin(), ineq[k].a.end());
                l.push_back(-INF); u.push_back(ineq[k].b);
            }
            for (int j = 0; j < d; ++j) c[j] = -ineq[i].a[j];
            LPResult res = lp_solve(be, c, A.data(), (int)mo, l.data(), u.data(), lp_count);
            drop = res.status == 1 && -res.obj <= ineq[i].b + tol;
        }
        if (drop) ineq.erase(ineq.begin() + i);
        else ++i;
    }
}

// Projection of the closed polyhedron P onto the coordinates keep (in that order).
static inline Poly project(const Poly& P, const std::vector<int>& keep, LPBackend& be, long* lp_count = nullptr, double tol = 1e-9) {
    const int d = P.d;
    std::vector<Ineq> eq, ineq;
    for (int i = 0; i < P.m(); ++i) {
        const double l = P.l[i], u = P.u[i];
        std::vector<double> a(P.row(i), P.row(i) + d);
        if (!std::isinf(l) && !std::isinf(u) && std::fabs(l - u) <= 1e-6) eq.push_back({a, u});
        else {
            if (!std::isinf(l)) { std::vector<double> na(d); for (int j = 0; j < d; ++j) na[j] = -a[j]; ineq.push_back({na, -l}); }
            if (!std::isinf(u)) ineq.push_back({a, u});
        }
    }
    std::vector<char> kept(d, 0);
    for (int j : keep) kept[j] = 1;
    for (int j = 0; j < d; ++j) {
        if (kept[j]) continue;
        int piv = -1;
        for (size_t k = 0; k < eq.size(); ++k)
            if (piv < 0 || std::fabs(eq[k].a[j]) > std::fabs(eq[piv].a[j])) piv = (int)k;
        if (piv >= 0 && std::fabs(eq[piv].a[j]) > tol) {
            Ineq e0 = std::move(eq[piv]);
            eq.erase(eq.begin() + piv);
            auto sub = [&](Ineq& r) {
                const double f = r.a[j] / e0.a[j];
                for (int c = 0; c < d; ++c) { const double t = f * e0.a[c]; r.a[c] = r.a[c] - t; }
                const double tb = f * e0.b;
                r.b = r.b - tb;
            };
            for (auto& r : eq) sub(r);
            for (auto& r : ineq) sub(r);
        } else {
            std::vector<Ineq> pos, neg, zer;
            for (auto& r : ineq) {
                if (r.a[j] > tol) pos.push_back(r);
                else if (r.a[j] < -tol) neg.push_back(r);
                else zer.push_back(r);
            }
            for (auto& p : pos)
                for (auto& q : neg) {
                    Ineq r;
                    r.a.resize(d);
                    for (int c = 0; c < d; ++c) { const double t1 = p.a[c] / p.a[j], t2 = q.a[c] / q.a[j]; r.a[c] = t1 - t2; }
                    const double b1 = p.b / p.a[j], b2 = q.b / q.a[j];
                    r.b = b1 - b2;
                    zer.push_back(std::move(r));
                }
            ineq.swap(zer);
        }
        for (auto& r : eq) r.a[j] = 0.0;
        for (auto& r : ineq) r.a[j] = 0.0;
        dedupe_rows(eq);
        dedupe_rows(ineq);
        if (ineq.size() > 24) irredundant(eq, ineq, be, d, lp_count);
    }
    irredundant(eq, ineq, be, d, lp_count);
    Rows out;
    out.d = (int)keep.size();
    std::vector<double> a(keep.size() ? keep.size() : 1);
    for (auto& r : eq) { for (size_t c = 0; c < keep.size(); ++c) a[c] = r.a[keep[c]]; out.add(a.data(), r.b, r.b); }
    for (auto& r : ineq) { for (size_t c = 0; c < keep.size(); ++c) a[c] = r.a[keep[c]]; out.add(a.data(), -INF, r.b); }
    return make_poly(std::move(out));
}

}  // namespace qpnnet
