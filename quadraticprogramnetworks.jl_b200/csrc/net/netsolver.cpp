// Native batched host state machine for networks with children -- see netsolver.hpp for the shape.
// Build: g++ -std=c++20 -O2 -ffp-contract=off (the geometry must round like the Python host mirror).
#include "netsolver.hpp"
#include "vertex_enum.h"

#include <cassert>
#include <condition_variable>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <set>
#include <thread>

namespace qpnnet {

enum { ERR_NONE = 0, ERR_CYCLE, ERR_AVI, ERR_DISAGREE, ERR_MAXIT, ERR_GRAPH_EMPTY, ERR_MASK, ERR_COMBINE, ERR_UNPOPULATED, ERR_NOPROJ };

// =================================================================================================================
// GeoCache
// =================================================================================================================
template <class Map, class Key, class F>
auto GeoCache::memo(Map& map, const Key& key, F&& compute) -> typename Map::mapped_type {
    {
        std::shared_lock<std::shared_mutex> lk(mu_);
        auto it = map.find(key);
        if (it != map.end()) return it->second;
    }
    auto value = compute();                      // outside the lock: pure, so a racing duplicate computes the same value
    std::unique_lock<std::shared_mutex> lk(mu_);
    auto ins = map.emplace(key, value);
    return ins.first->second;
}

int GeoCache::intern_poly(Poly&& P, Worker* w) {
    std::string ek = exact_key(P);
    {
        std::shared_lock<std::shared_mutex> lk(mu_);
        auto it = poly_by_exact_.find(ek);
        if (it != poly_by_exact_.end()) return it->second;
    }
    std::lock_guard<std::mutex> cl(create_mu_);
    int id;
    {
        std::shared_lock<std::shared_mutex> lk(mu_);
        auto it = poly_by_exact_.find(ek);
        if (it != poly_by_exact_.end()) return it->second;
        id = (int)polys_.size();
    }
    std::string sk = set_key(P);
    store_->new_piece(id, P, w);                 // resident before anybody can ask about it
    std::unique_lock<std::shared_mutex> lk(mu_);
    auto sit = set_by_key_.emplace(std::move(sk), (int)set_by_key_.size());
    polys_.push_back(std::move(P));
    set_ids_.push_back(sit.first->second);
    poly_by_exact_.emplace(std::move(ek), id);
    stats.pieces++;
    return id;
}

int GeoCache::intern_list(const std::vector<int>& ids) {
    {
        std::shared_lock<std::shared_mutex> lk(mu_);
        auto it = list_ids_.find(ids);
        if (it != list_ids_.end()) return it->second;
    }
    std::unique_lock<std::shared_mutex> lk(mu_);
    auto it = list_ids_.find(ids);
    if (it != list_ids_.end()) return it->second;
    const int id = (int)lists_.size();
    lists_.push_back(ids);
    list_ids_.emplace(ids, id);
    return id;
}

// ---- assembly (avi.jl:205-251,305-377,447-475; qp_processing.jl:57-66) --------------------------------------------
static void stack_polys(const GeoCache& c, const std::vector<int>& polys, int nv, std::vector<double>& A_rm, std::vector<double>& l,
                        std::vector<double>& u) {
    for (int id : polys) {
        const Poly& P = c.poly(id);
        A_rm.insert(A_rm.end(), P.A.begin(), P.A.end());
        l.insert(l.end(), P.l.begin(), P.l.end());
        u.insert(u.end(), P.u.begin(), P.u.end());
    }
    (void)nv;
}

int GeoCache::node(int pid, const std::vector<int>& pieces, Worker* w) {
    std::vector<int> key;
    key.push_back(pid);
    key.insert(key.end(), pieces.begin(), pieces.end());
    {
        std::shared_lock<std::shared_mutex> lk(mu_);
        auto it = node_ids_.find(key);
        if (it != node_ids_.end()) return it->second;
    }
    std::lock_guard<std::mutex> cl(create_mu_);
    int id;
    {
        std::shared_lock<std::shared_mutex> lk(mu_);
        auto it = node_ids_.find(key);
        if (it != node_ids_.end()) return it->second;
        id = (int)nodes_.size();
    }
    const NetData& net = net_;
    NodeInfo n;
    n.pid = pid; n.nv = net.nv;
    n.polys = net.base[pid];
    n.polys.insert(n.polys.end(), pieces.begin(), pieces.end());
    const std::vector<int>& dec = net.dec[pid];
    n.nd = (int)dec.size();
    n.dec.assign(dec.begin(), dec.end());
    std::vector<char> isdec(net.nv, 0);
    for (int d : dec) isdec[d] = 1;
    for (int j = 0; j < net.nv; ++j) if (!isdec[j]) n.par.push_back(j);
    std::vector<double> A, l, u;
    stack_polys(*this, n.polys, net.nv, A, l, u);
    const int m = (int)l.size(), nv = net.nv, nd = n.nd, np = (int)n.par.size();
    n.m = m;
    n.l = l; n.u = u;
    n.A.assign((size_t)m * nv, 0.0);
    for (int i = 0; i < m; ++i) for (int j = 0; j < nv; ++j) n.A[(size_t)j * m + i] = A[(size_t)i * nv + j];
    const std::vector<double>& Q = net.Q[pid];
    n.Qd.assign((size_t)nd * nv, 0.0);
    n.qd.resize(nd);
    for (int e = 0; e < nd; ++e) {
        for (int j = 0; j < nv; ++j) n.Qd[(size_t)j * nd + e] = Q[(size_t)dec[e] * nv + j];
        n.qd[e] = net.q[pid][dec[e]];
    }
    // single-node GAVI: (Q_dd x_d + Q_dp w + q_d - A_d' lam) comp. x_d free ; lam comp. l <= A_d x_d + A_p w <= u
    GaviData& g = n.g;
    g.d1 = nd; g.d2 = m; g.np = np;
    const int dz = nd + m;
    g.M.assign((size_t)nd * dz, 0.0);
    for (int e = 0; e < nd; ++e) {
        for (int c = 0; c < nd; ++c) g.M[(size_t)c * nd + e] = Q[(size_t)dec[e] * nv + dec[c]];
        for (int i = 0; i < m; ++i) g.M[(size_t)(nd + i) * nd + e] = -A[(size_t)i * nv + dec[e]];
    }
    g.N.assign((size_t)nd * np, 0.0);
    for (int e = 0; e < nd; ++e) for (int c = 0; c < np; ++c) g.N[(size_t)c * nd + e] = Q[(size_t)dec[e] * nv + n.par[c]];
    g.o = n.qd;
    g.l1.assign(nd, -INF); g.u1.assign(nd, INF);
    g.A.assign((size_t)m * dz, 0.0);
    for (int i = 0; i < m; ++i) for (int c = 0; c < nd; ++c) g.A[(size_t)c * m + i] = A[(size_t)i * nv + dec[c]];
    g.B.assign((size_t)m * np, 0.0);
    for (int i = 0; i < m; ++i) for (int c = 0; c < np; ++c) g.B[(size_t)c * m + i] = A[(size_t)i * nv + n.par[c]];
    g.l2 = l; g.u2 = u;
    store_->new_node(id, n, w);
    std::unique_lock<std::shared_mutex> lk(mu_);
    nodes_.push_back(std::move(n));
    node_ids_.emplace(std::move(key), id);
    stats.nodes++;
    return id;
}

int GeoCache::level_gavi(int level, const std::vector<int>& assignment, Worker* w) {
    std::vector<int> key;
    key.push_back(level);
    key.insert(key.end(), assignment.begin(), assignment.end());
    {
        std::shared_lock<std::shared_mutex> lk(mu_);
        auto it = gavi_ids_.find(key);
        if (it != gavi_ids_.end()) return it->second;
    }
    std::lock_guard<std::mutex> cl(create_mu_);
    int id;
    {
        std::shared_lock<std::shared_mutex> lk(mu_);
        auto it = gavi_ids_.find(key);
        if (it != gavi_ids_.end()) return it->second;
        id = (int)gavis_.size();
    }
    const NetData& net = net_;
    const std::vector<int>& players = net.levels[level];
    // children of the level, sorted: assignment[k] is the piece of the k-th child
    std::vector<int> kids;
    for (int p : players) kids.insert(kids.end(), net.children[p].begin(), net.children[p].end());
    std::sort(kids.begin(), kids.end());
    kids.erase(std::unique(kids.begin(), kids.end()), kids.end());
    std::map<int, int> piece_of;
    for (size_t k = 0; k < kids.size(); ++k) piece_of[kids[k]] = assignment[k];
    const int nv = net.nv;
    LevelGaviInfo L;
    L.level = level;
    std::vector<char> isdec(nv, 0);
    for (int p : players) for (int d : net.dec[p]) isdec[d] = 1;
    std::vector<int> dec, par, dpos(nv, -1);
    for (int j = 0; j < nv; ++j) { if (isdec[j]) { dpos[j] = (int)dec.size(); dec.push_back(j); } else par.push_back(j); }
    const int nd = (int)dec.size(), np = (int)par.size();
    struct View { std::vector<double> A, l, u; int k, m; };
    std::vector<View> views;
    int total_xi = 0, total_lam = 0;
    for (int p : players) {
        View v;
        std::vector<int> polys = net.base[p];
        for (int j : net.children[p]) polys.push_back(piece_of[j]);
        stack_polys(*this, polys, nv, v.A, v.l, v.u);
        v.k = (int)net.dec[p].size(); v.m = (int)v.l.size();
        total_xi += v.k; total_lam += v.m;
        views.push_back(std::move(v));
    }
    GaviData& g = L.g;
    const int d1 = nd + total_xi, d2 = total_lam, dz = d1 + d2;
    g.d1 = d1; g.d2 = d2; g.np = np;
    g.M.assign((size_t)d1 * dz, 0.0); g.N.assign((size_t)d1 * np, 0.0); g.o.assign(d1, 0.0);
    g.A.assign((size_t)d2 * dz, 0.0); g.B.assign((size_t)d2 * np, 0.0); g.l2.assign(d2, 0.0); g.u2.assign(d2, 0.0);
    g.l1.assign(d1, -INF); g.u1.assign(d1, INF);
    int xi_off = 0, lam_off = 0, row = nd;
    for (size_t pi = 0; pi < players.size(); ++pi) {
        const int p = players[pi];
        const View& v = views[pi];
        const std::vector<int>& decp = net.dec[p];
        const std::vector<double>& Q = net.Q[p];
        for (int e = 0; e < v.k; ++e) {
            for (int c = 0; c < nd; ++c) g.M[(size_t)c * d1 + row + e] = Q[(size_t)decp[e] * nv + dec[c]];
            for (int c = 0; c < np; ++c) g.N[(size_t)c * d1 + row + e] = Q[(size_t)decp[e] * nv + par[c]];
            g.o[row + e] = net.q[p][decp[e]];
            for (int i = 0; i < v.m; ++i) g.M[(size_t)(d1 + lam_off + i) * d1 + row + e] = -v.A[(size_t)i * nv + decp[e]];
            g.M[(size_t)(nd + xi_off + e) * d1 + dpos[decp[e]]] = 1.0;        // top rows: sum of the owners' xi = 0
        }
        for (int i = 0; i < v.m; ++i) {
            for (int c = 0; c < nd; ++c) g.A[(size_t)c * d2 + lam_off + i] = v.A[(size_t)i * nv + dec[c]];
            for (int c = 0; c < np; ++c) g.B[(size_t)c * d2 + lam_off + i] = v.A[(size_t)i * nv + par[c]];
            g.l2[lam_off + i] = v.l[i]; g.u2[lam_off + i] = v.u[i];
        }
        row += v.k; xi_off += v.k; lam_off += v.m;
    }
    L.dec.assign(dec.begin(), dec.end());
    L.par.assign(par.begin(), par.end());
    store_->new_gavi(id, L, w);
    std::unique_lock<std::shared_mutex> lk(mu_);
    gavis_.push_back(std::move(L));
    gavi_ids_.emplace(std::move(key), id);
    stats.gavis++;
    return id;
}

// ---- predicates ----------------------------------------------------------------------------------------------
static inline uint64_t pair_key(int a, int b) { return ((uint64_t)(uint32_t)a << 32) | (uint32_t)b; }

bool GeoCache::empty(int pid, double tol, Worker* w) {
    // two tolerances are in use (1e-4 everywhere on this path); fold the tolerance into the key
    const uint64_t key = pair_key(pid, tol == 1e-4 ? 0 : 1 + (int)std::lround(-std::log10(tol)));
    return memo(empty_, key, [&]() -> char {
        long n = 0;
        const bool e = exemplar_empty(poly(pid), *w, tol, &n);
        stats.lps += n; stats.lps_empty += n; stats.lp_calls += n;
        return (char)e;
    }) != 0;
}

bool GeoCache::subset(int p1, int p2, Worker* w) {
    return memo(subset_, pair_key(p1, p2), [&]() -> char {
        // the LPs of P2's bounds over P1 in batched calls of 1, 2, 4, ... (issubset_many): a pair that is no subset costs
        // about what the serial scan of sets.jl:377-407 costs, a pair that is one a handful of launches instead of 2 m
        long n = 0, calls = 0;
        std::vector<char> res;
        issubset_many(poly(p1), {&poly(p2)}, *w, res, 1e-6, &n, &calls);
        stats.lps += n; stats.lps_subset += n; stats.lp_calls += calls;
        return res[0];
    }) != 0;
}

int GeoCache::remove_subsets(int lid, Worker* w) {
    return memo(remove_subsets_, lid, [&]() -> int {
        const std::vector<int> ids = list(lid);
        const int k = (int)ids.size();
        if (w->wide_batches()) {
            // every pair the scan below may ask for, the pairs of one P1 in a handful of wide calls (the scan stops at the
            // first superset, so some of these LPs it would have skipped: on a GPU they ride along for free)
            for (int i = 0; i < k; ++i) {
                std::vector<int> todo;
                {
                    std::shared_lock<std::shared_mutex> lk(mu_);
                    for (int j = 0; j < k; ++j)
                        if (j != i && ids[j] != ids[i] && !subset_.count(pair_key(ids[i], ids[j])) &&
                            std::find(todo.begin(), todo.end(), ids[j]) == todo.end()) todo.push_back(ids[j]);
                }
                if (todo.size() < 2) continue;
                std::vector<const Poly*> P2s;
                for (int id : todo) P2s.push_back(&poly(id));
                std::vector<char> res;
                long n = 0, calls = 0;
                issubset_many(poly(ids[i]), P2s, *w, res, 1e-6, &n, &calls);
                stats.lps += n; stats.lps_subset += n; stats.lp_calls += calls;
                std::unique_lock<std::shared_mutex> lk(mu_);
                for (size_t q = 0; q < todo.size(); ++q) subset_.emplace(pair_key(ids[i], todo[q]), res[q]);
            }
        }
        std::vector<char> is_sub(k, 0);
        for (int i = 0; i < k; ++i)
            for (int j = 0; j < k; ++j)
                if (i != j && !is_sub[j] && subset(ids[i], ids[j], w)) { is_sub[i] = 1; break; }
        std::vector<int> out;
        for (int i = 0; i < k; ++i) if (!is_sub[i]) out.push_back(ids[i]);
        return intern_list(out);
    });
}

int GeoCache::intersect2(int a, int b, Worker* w) {
    return memo(intersect2_, pair_key(a, b), [&]() -> int { return intern_poly(intersect(poly(a), poly(b)), w); });
}

int GeoCache::intersect_all(const std::vector<int>& ids, Worker* w) {
    // ph.intersect(*polys): rows of the first, then of the second, ...; one poly alone is rebuilt as it is
    int cur = ids[0];
    for (size_t k = 1; k < ids.size(); ++k) {
        // intersect(intersect(p1, p2), p3) lists the rows in the order of intersect(p1, p2, p3)
        cur = memo(intersect2_, pair_key(cur, ids[k]) ^ 0x8000000000000000ull,
                   [&]() -> int { return intern_poly(intersect(poly(cur), poly(ids[k])), w); });
    }
    return cur;
}

const std::vector<int>& GeoCache::complement_of(int pid, Worker* w) {
    {
        std::shared_lock<std::shared_mutex> lk(mu_);
        auto it = complement_.find(pid);
        if (it != complement_.end()) return it->second;
    }
    std::vector<int> ids;
    for (Poly& c : complement(poly(pid))) ids.push_back(intern_poly(std::move(c), w));
    std::unique_lock<std::shared_mutex> lk(mu_);
    return complement_.emplace(pid, std::move(ids)).first->second;
}

// ---- local pieces (avi_solutions.jl:400-496, 79-90, 241-261) ---------------------------------------------------
static Poly build_local_piece(const GaviData& g, const std::string& K) {
    const int d1 = g.d1, d2 = g.d2, n = d1 + d2, m = g.np, d = n + m;
    Rows r;
    r.d = d;
    std::vector<double> row(d);
    std::vector<double> lo(2 * n), up(2 * n);
    for (int i = 0; i < n; ++i) {
        const int k = K[i];
        double a, b, c, e;
        if (i < d1) {
            const double o = g.o[i], l = g.l1[i], u = g.u1[i];
            switch (k) {
                case 1: a = -o; b = INF; c = l; e = l; break;
                case 2: a = -o; b = -o; c = l; e = u; break;
                case 3: a = -INF; b = -o; c = u; e = u; break;
                default: a = -INF; b = INF; c = l; e = u; break;
            }
        } else {
            const double l = g.l2[i - d1], u = g.u2[i - d1];
            switch (k) {
                case 1: a = 0.0; b = INF; c = l; e = l; break;
                case 2: a = 0.0; b = 0.0; c = l; e = u; break;
                case 3: a = -INF; b = 0.0; c = u; e = u; break;
                default: a = -INF; b = INF; c = l; e = u; break;
            }
        }
        lo[i] = a; up[i] = b; lo[n + i] = c; up[n + i] = e;
    }
    for (int i = 0; i < 2 * n; ++i) if (lo[i] > up[i]) lo[i] = up[i];
    for (int rr = 0; rr < 2 * n; ++rr) {
        std::fill(row.begin(), row.end(), 0.0);
        if (rr < d1) {                               // [M N]
            for (int c = 0; c < n; ++c) row[c] = g.M[(size_t)c * d1 + rr];
            for (int c = 0; c < m; ++c) row[n + c] = g.N[(size_t)c * d1 + rr];
        } else if (rr < n) {                         // [0 I2 0]
            row[d1 + (rr - d1)] = 1.0;
        } else if (rr < n + d1) {                    // [I1 0 0]
            row[rr - n] = 1.0;
        } else {                                     // [A B]
            const int i = rr - n - d1;
            for (int c = 0; c < n; ++c) row[c] = g.A[(size_t)c * d2 + i];
            for (int c = 0; c < m; ++c) row[n + c] = g.B[(size_t)c * d2 + i];
        }
        bool any = false;
        for (int c = 0; c < d; ++c) { if (std::fabs(row[c]) <= 1e-8) row[c] = 0.0; if (row[c] != 0.0) any = true; }
        if ((std::isinf(lo[rr]) && std::isinf(up[rr])) || !any) continue;
        r.add(row.data(), lo[rr], up[rr]);
    }
    if (r.m() == 0) return make_poly(std::move(r));
    return simplify(make_poly(std::move(r)));
}

int GeoCache::local_piece_id(int node, const std::string& K, Worker* w) {
    std::string key((const char*)&node, 4);
    key += K;
    return memo(local_piece_, key, [&]() -> int { return intern_poly(build_local_piece(node_info(node).g, K), w); });
}

int GeoCache::expand(int node, const std::string& K, Worker* w) {
    std::string key((const char*)&node, 4);
    key += K;
    return memo(expand_, key, [&]() -> int {
        const NodeInfo& n = node_info(node);
        const int lp = local_piece_id(node, K, w);
        const Poly& piece = poly(lp);
        if (piece.m() > 0 && empty(lp, 1e-4, w)) return -1;
        // project_and_permute: keep [z[0:nd]; w], scatter the columns to x's ordering
        const int d = piece.d, nd = n.nd, np = (int)n.par.size();
        std::vector<int> keep;
        for (int j = 0; j < nd; ++j) keep.push_back(j);
        for (int j = d - np; j < d; ++j) keep.push_back(j);
        long nl = 0;
        Poly pr = project(piece, keep, *w, &nl);
        stats.lps += nl; stats.lps_project += nl; stats.lp_calls += nl;
        Rows r;
        r.d = n.nv;
        std::vector<double> a(n.nv);
        for (int i = 0; i < pr.m(); ++i) {
            std::fill(a.begin(), a.end(), 0.0);
            for (int c = 0; c < nd; ++c) a[n.dec[c]] = pr.row(i)[c];
            for (int c = 0; c < np; ++c) a[n.par[c]] = pr.row(i)[nd + c];
            r.add(a.data(), pr.l[i], pr.u[i]);
        }
        return intern_poly(simplify(make_poly(std::move(r))), w);
    });
}

// all_Ks (avi_solutions.jl:200-215) in lexicographic order; false when an index belongs to no set
static bool all_Ks(const std::vector<int8_t>& mask, std::vector<std::string>& out) {
    const size_t n = mask.size();
    std::vector<std::vector<char>> choices(n);
    for (size_t i = 0; i < n; ++i) {
        for (int b = 0; b < 4; ++b) if ((mask[i] >> b) & 1) choices[i].push_back((char)(b + 1));
        if (choices[i].empty()) return false;
    }
    std::vector<size_t> idx(n, 0);
    std::string K(n, 0);
    while (true) {
        for (size_t i = 0; i < n; ++i) K[i] = choices[i][idx[i]];
        out.push_back(K);
        size_t i = n;
        while (i > 0) {
            --i;
            if (++idx[i] < choices[i].size()) break;
            idx[i] = 0;
            if (i == 0) return true;
        }
        if (n == 0) return true;
    }
}

int GeoCache::collect(int node, const std::vector<int8_t>& mask, const std::vector<int8_t>& extra, Worker* w, bool* bad_mask) {
    std::string key((const char*)&node, 4);
    key.append((const char*)mask.data(), mask.size());
    key.append((const char*)extra.data(), extra.size());
    *bad_mask = false;
    const int r = memo(collect_, key, [&]() -> int {
        stats.collect_miss++;
        // collect(LocalGAVISolutions), avi_solutions.jl:277-321: the recipes of the point first, then those of the
        // explored vertices that are new, each group in sorted order
        std::vector<std::string> Ks;
        if (!all_Ks(mask, Ks)) return -2;
        std::set<std::string> explored(Ks.begin(), Ks.end());
        std::set<std::string> fresh;
        const size_t dz = mask.size();
        for (size_t o = 0; dz > 0 && o + dz <= extra.size(); o += dz) {
            std::vector<int8_t> vm(extra.begin() + o, extra.begin() + o + dz);
            std::vector<std::string> Kv;
            if (!all_Ks(vm, Kv)) return -2;
            for (auto& K : Kv) if (!explored.count(K)) fresh.insert(K);
        }
        Ks.insert(Ks.end(), fresh.begin(), fresh.end());
        std::vector<int> out;
        std::set<int> seen_sets;
        for (const std::string& K : Ks) {
            const int p = expand(node, K, w);
            if (p < 0) continue;
            if (!seen_sets.insert(set_id(p)).second) continue;
            out.push_back(p);
        }
        return intern_list(out);
    });
    if (r == -2) { *bad_mask = true; return -1; }
    return r;
}

int GeoCache::leaves(const std::vector<int>& union_lists, const std::vector<int>& red_lengths, const std::vector<uint8_t>& in_bits,
                     Worker* w) {
    std::string key;
    key.append((const char*)union_lists.data(), union_lists.size() * 4);
    key.append((const char*)red_lengths.data(), red_lengths.size() * 4);
    key.append((const char*)in_bits.data(), in_bits.size());
    return memo(leaves_, key, [&]() -> int {
        stats.combine_miss++;
        const int n = (int)union_lists.size();
        std::vector<std::vector<int>> unions(n);
        std::vector<std::vector<uint8_t>> inside(n);
        size_t bit = 0;
        for (int k = 0; k < n; ++k) {
            unions[k] = list(union_lists[k]);
            for (int p : unions[k]) inside[k].push_back(poly(p).m() > 0 ? in_bits[bit++] : 1);
        }
        std::vector<int> out, idx(n, 0);
        std::function<void(int, int)> rec = [&](int depth, int cur) {
            if (depth == n) {
                bool red = true;
                for (int k = 0; k < n; ++k) if (!(idx[k] >= (int)unions[k].size() - red_lengths[k])) red = false;
                if (red) return;                     // the all-complements "red zone" (intersection.jl:123)
                out.push_back(cur);
                return;
            }
            for (int k = 0; k < (int)unions[depth].size(); ++k) {
                const int piece = unions[depth][k];
                if (!inside[depth][k]) continue;
                const int nxt = cur < 0 ? piece : intersect2(piece, cur, w);
                if (empty(nxt, 1e-4, w)) continue;
                idx[depth] = k;
                rec(depth + 1, nxt);
            }
        };
        rec(0, -1);
        return intern_list(out);
    });
}

// Iterators.product order: the FIRST iterator varies fastest (qp_processing.jl:169)
static void julia_product(const std::vector<int>& sizes, std::vector<std::vector<int>>& out) {
    const size_t n = sizes.size();
    for (int s : sizes) if (s == 0) return;
    std::vector<int> idx(n, 0);
    while (true) {
        out.push_back(idx);
        size_t i = 0;
        while (i < n && ++idx[i] == sizes[i]) { idx[i] = 0; ++i; }
        if (i == n) break;
    }
}

// ---- memoised transitions ---------------------------------------------------------------------------------------
// process_qp, first phase: every player of the level against every combination of its children's pieces
int GeoCache::verify_plan(int level, const std::vector<int>& S, Worker* w) {
    std::vector<int> key;
    key.reserve(S.size() + 1);
    key.push_back(level);
    key.insert(key.end(), S.begin(), S.end());
    {
        std::shared_lock<std::shared_mutex> lk(mu_);
        auto it = plan_ids_.find(key);
        if (it != plan_ids_.end()) return it->second;
    }
    const NetData& net = net_;
    VerifyPlan P;
    P.level = level; P.S = S;
    // solution graphs are built for every level but the first (qp_processing.jl:158); that is where vertices matter
    const bool gen = level != 0 || net.gen_solution_map;
    P.want = (gen && net.exploration_vertices > 1) ? std::min(net.exploration_vertices - 1, (int)QPN_VE_MAXV) : 0;
    for (int pid : net.levels[level]) {
        VerifyPlan::PV pv;
        pv.pid = pid; pv.first_req = (int)P.req_nodes.size();
        const std::vector<int>& ch = net.children[pid];
        if (!ch.empty()) {
            std::vector<int> sizes;
            for (int j : ch) {
                if (S[j] < 0 || list(S[j]).empty()) { P.error = ERR_UNPOPULATED; break; }
                sizes.push_back((int)list(S[j]).size());
            }
            if (P.error) break;
            julia_product(sizes, pv.combos);
            for (const auto& combo : pv.combos) {
                std::vector<int> pieces;
                for (size_t k = 0; k < ch.size(); ++k) pieces.push_back(list(S[ch[k]])[combo[k]]);
                P.req_nodes.push_back(node(pid, pieces, w));
            }
        } else {
            P.req_nodes.push_back(node(pid, {}, w));
        }
        P.pvs.push_back(std::move(pv));
    }
    P.rep_bytes = 1;
    for (int nid : P.req_nodes) { const NodeInfo& n = node_info(nid); P.rep_bytes += verify_rep_bytes(n.nd + n.m, n.m, P.want); }
    std::unique_lock<std::shared_mutex> lk(mu_);
    auto it = plan_ids_.find(key);
    if (it != plan_ids_.end()) return it->second;
    const int id = (int)plans_.size();
    plans_.push_back(std::move(P));
    plan_ids_.emplace(std::move(key), id);
    return id;
}

// algorithm.jl:84,104-116: the level's graphs (after remove_subsets where the level asks for it) join those from below
int GeoCache::finish_graphs(const VerifyPlan& P, const std::vector<int>& S_out, Worker* w) {
    std::vector<int> S = P.S;
    for (int pid : net_.levels[P.level]) {
        if (S_out[pid] >= 0 && net_.remove_subsets_at[P.level]) S[pid] = remove_subsets(S_out[pid], w);
        else S[pid] = S_out[pid];
    }
    return intern_list(S);
}

int GeoCache::verify_outcome(int plan_id, const uint8_t* rep, Worker* w) {
    const VerifyPlan& P = plan(plan_id);
    const NetData& net = net_;
    const size_t nr = P.req_nodes.size();
    // the key: the answers that matter (flags; masks and vertex masks where a node is at a solution)
    std::string key((const char*)&plan_id, 4);
    {
        const uint8_t* r = rep;
        for (size_t k = 0; k < nr; ++k) {
            const NodeInfo& info = node_info(P.req_nodes[k]);
            const int dz = info.nd + info.m, vbytes = (info.m + 1) / 2;
            key.push_back((char)r[0]);
            if (r[0]) {
                const int nvx = std::min((int)r[1 + dz], P.want);
                key.append((const char*)r + 1, (size_t)dz);
                key.push_back((char)nvx);
                key.append((const char*)r + 2 + dz, (size_t)nvx * vbytes);
            }
            r += verify_rep_bytes(dz, info.m, P.want);
        }
    }
    {
        std::shared_lock<std::shared_mutex> lk(mu_);
        auto it = outcome_ids_.find(key);
        if (it != outcome_ids_.end()) return it->second;
    }
    // ---- parse ---------------------------------------------------------------------------------------------------------
    std::vector<uint8_t> sol(nr, 0);
    std::vector<std::vector<int8_t>> masks(nr), vmasks(nr);
    for (size_t r = 0; r < nr; ++r) {
        const NodeInfo& info = node_info(P.req_nodes[r]);
        const int mrows = info.m, nd = info.nd, dz = nd + mrows, vbytes = (mrows + 1) / 2;
        const uint8_t* mask = rep + 1;
        const int nvx = rep[1 + dz];
        const uint8_t* vmask = rep + 2 + dz;
        sol[r] = rep[0];
        rep += verify_rep_bytes(dz, mrows, P.want);
        if (!sol[r]) continue;
        masks[r].assign((const int8_t*)mask, (const int8_t*)mask + dz);
        // a vertex keeps the primal part of the point: only the masks of the m multiplier rows change
        for (int q = 0; q < nvx && q < P.want; ++q) {
            const uint8_t* nib = vmask + (size_t)q * vbytes;
            std::vector<int8_t> vm = masks[r];
            for (int i = 0; i < mrows; ++i) vm[nd + i] = (int8_t)((nib[i >> 1] >> ((i & 1) * 4)) & 0xf);
            vmasks[r].insert(vmasks[r].end(), vm.begin(), vm.end());
        }
    }
    VerifyOutcome O;
    O.plan = plan_id;
    const int level = P.level;
    bool equilibrium = true;
    for (uint8_t s : sol) if (!s) equilibrium = false;
    auto failed = [&](int err) { O.kind = VerifyOutcome::FAIL; O.error = err; };
    // ---- process_qp, second phase, for one player whose every combination verified: its solution graph, or what is needed to
    // combine the graphs of its combinations (qp_processing.jl:189-218,243-291).  Returns an error code (0 = none).
    auto build_player = [&](const VerifyPlan::PV& pv, std::vector<int>& S_out, std::vector<VerifyOutcome::Comb>& combs) -> int {
        const int pid = pv.pid;
        const bool gen = level != 0 || net.gen_solution_map;
        if (!gen) return 0;
        const std::vector<int>& ch = net.children[pid];
        if (ch.empty()) {
            bool bad = false;
            const int lid = collect(P.req_nodes[pv.first_req], masks[pv.first_req], vmasks[pv.first_req], w, &bad);
            if (bad) return ERR_MASK;
            if (list(lid).empty()) return ERR_GRAPH_EMPTY;
            S_out[pid] = lid;
            return 0;
        }
        std::vector<int> sols;
        for (size_t k = 0; k < pv.combos.size(); ++k) {
            bool bad = false;
            const int lid = collect(P.req_nodes[pv.first_req + k], masks[pv.first_req + k], vmasks[pv.first_req + k], w, &bad);
            if (bad) return ERR_MASK;
            sols.push_back(remove_subsets(lid, w));
        }
        if (sols.size() == 1) { S_out[pid] = sols[0]; return 0; }
        VerifyOutcome::Comb cb;
        cb.pid = pid;
        int total = 0;
        for (size_t k = 0; k < pv.combos.size(); ++k) {
            std::vector<int> pieces;
            for (size_t q = 0; q < ch.size(); ++q) pieces.push_back(list(P.S[ch[q]])[pv.combos[k][q]]);
            const int region = intersect_all(pieces, w);
            const std::vector<int>& comp = complement_of(region, w);
            std::vector<int> combined = list(sols[k]);
            combined.insert(combined.end(), comp.begin(), comp.end());
            total += (int)combined.size();
            cb.red.push_back((int)comp.size());
            for (int p : combined) if (poly(p).m() > 0) cb.flat.push_back(p);
            cb.union_lists.push_back(intern_list(combined));
        }
        if (cb.union_lists.size() > 3 && total > 20) return ERR_COMBINE;
        combs.push_back(std::move(cb));
        return 0;
    };
    if (equilibrium) {
        O.S_out.assign(net.nplayers, -1);
        O.kind = VerifyOutcome::DONE;
        for (const VerifyPlan::PV& pv : P.pvs) {
            const int err = build_player(pv, O.S_out, O.combs);
            if (err) { failed(err); break; }
        }
        if (O.kind != VerifyOutcome::FAIL) {
            if (!O.combs.empty()) O.kind = VerifyOutcome::MEMBER;
            else O.S_new = finish_graphs(P, O.S_out, w);
        }
    } else {
        // ---- not an equilibrium: solve_qep with the offending child pieces (algorithm.jl:68-101).  The reference has by now
        // built the graph of every player that DID verify (process_qp does both phases per player) and ends the solve if that
        // raised or the player's combine failed (algorithm.jl:56-63,120-126), although the graphs themselves are thrown away:
        // the same checks here, the graphs memoised for the pass in which the level does verify.
        O.kind = VerifyOutcome::QEP;
        for (const VerifyPlan::PV& pv : P.pvs) {
            bool all_sol = true;
            const size_t nreq = pv.combos.empty() ? 1 : pv.combos.size();
            for (size_t k = 0; k < nreq; ++k) if (!sol[pv.first_req + k]) all_sol = false;
            if (!all_sol) continue;
            std::vector<int> S_tmp(net.nplayers, -1);
            std::vector<VerifyOutcome::Comb> combs_tmp;
            const int err = build_player(pv, S_tmp, combs_tmp);
            if (err) { failed(err); break; }
        }
    }
    if (!equilibrium && O.kind != VerifyOutcome::FAIL) {
        std::vector<int> kids;
        for (int p : net.levels[level]) kids.insert(kids.end(), net.children[p].begin(), net.children[p].end());
        std::sort(kids.begin(), kids.end());
        kids.erase(std::unique(kids.begin(), kids.end()), kids.end());
        std::vector<int> assignment(kids.size());
        for (size_t k = 0; k < kids.size(); ++k) assignment[k] = list(P.S[kids[k]])[0];
        for (const VerifyPlan::PV& pv : P.pvs) {
            const std::vector<int>& ch = net.children[pv.pid];
            if (ch.empty()) continue;
            for (size_t k = 0; k < pv.combos.size(); ++k) {
                if (sol[pv.first_req + k]) continue;
                for (size_t q = 0; q < ch.size(); ++q) {
                    const size_t pos = std::lower_bound(kids.begin(), kids.end(), ch[q]) - kids.begin();
                    assignment[pos] = list(P.S[ch[q]])[pv.combos[k][q]];
                }
                break;                           // the first combination that fails
            }
        }
        O.gavi = level_gavi(level, assignment, w);
    }
    std::unique_lock<std::shared_mutex> lk(mu_);
    auto it = outcome_ids_.find(key);
    if (it != outcome_ids_.end()) return it->second;
    const int id = (int)outcomes_.size();
    outcomes_.push_back(std::move(O));
    // (the lists of a stored outcome never move: the membership request points at them)
    VerifyOutcome& stored = const_cast<VerifyOutcome&>(outcomes_[id]);
    for (const VerifyOutcome::Comb& cb : stored.combs) stored.comb_lists.push_back(&cb.flat);
    outcome_ids_.emplace(std::move(key), id);
    return id;
}

int GeoCache::member_outcome(int outcome_id, const uint8_t* bits, Worker* w) {
    const VerifyOutcome& O = outcome(outcome_id);
    size_t np = 0;
    for (const VerifyOutcome::Comb& cb : O.combs) np += cb.flat.size();
    std::string key((const char*)&outcome_id, 4);
    key.append((const char*)bits, np);
    return memo(member_ids_, key, [&]() -> int {
        std::vector<int> S_out = O.S_out;
        const uint8_t* b = bits;
        for (const VerifyOutcome::Comb& cb : O.combs) {
            S_out[cb.pid] = leaves(cb.union_lists, cb.red, std::vector<uint8_t>(b, b + cb.flat.size()), w);
            b += cb.flat.size();
        }
        return finish_graphs(plan(O.plan), S_out, w);
    });
}

// =================================================================================================================
// the cohort state machine (algorithm.jl:1-127 + qp_processing.jl:151-291, explicit and copyable)
// =================================================================================================================
namespace {

struct Frame {                                   // one activation of solve_base! at a level
    int level = 0, it = 0;
    int plan = -1, outcome = -1;                 // memoised: what is being verified, what the answers led to
    std::vector<int> S;                          // per player: list id of its solution graph, -1 = none
};

enum Wait { W_NONE, W_VERIFY, W_MEMBER, W_QEP };

struct Cohort {
    Seg seg;                                     // its members in the worker's order array
    std::vector<Frame> stack;
    int cyc_level = 0, ncyc = 0;                 // cycle checks that ride with the next verify request (levels cyc_level ..)
    std::vector<int> level_iters;
    Wait wait = W_NONE;
    bool done = false, solved = false;
    int error = 0;
    std::vector<int> sol;
};

struct Machine {
    GeoCache& c;
    const NetData& net;
    Worker* w;
    std::vector<std::unique_ptr<Cohort>> ready;  // to be looked at by the driver
    int bottom_plan = -1;                        // what the first pass of the bottom level verifies (no graphs below it)

    Machine(GeoCache& cc, Worker* ww) : c(cc), net(cc.net()), w(ww) {
        bottom_plan = c.verify_plan(net.nlevels - 1, std::vector<int>(net.nplayers, -1), w);
    }

    void finish(Cohort& C, bool solved, int err) {
        C.done = true; C.solved = solved; C.error = err; C.wait = W_NONE;
        if (solved) C.sol = C.stack.front().S;
    }
    void fail(Cohort& C, int err) { finish(C, false, err); }

    // ---- algorithm.jl:13-39 --------------------------------------------------------------------------------------------
    // A loop pass begins with the cycle check of its level and, above the last level, with the full recursive solve of
    // the level below -- whose first pass begins the same way.  Nothing numeric happens until the bottom level's nodes
    // are verified and x does not change on the way down, so the checks of the whole chain are posted together with
    // that verify request; if one of them hits, the passes opened below it are taken back (apply()).
    void start_iter(Cohort& C) {
        while (true) {
            Frame& f = C.stack.back();
            if (f.it == net.max_iters) return fail(C, ERR_MAXIT);
            f.it++;
            C.level_iters[f.level]++;
            if (net.check_for_cycling) {
                if (net.num_projections == 0) return fail(C, ERR_NOPROJ);
                if (C.ncyc == 0) C.cyc_level = f.level;
                C.ncyc++;
            }
            f.S.assign(net.nplayers, -1);
            if (f.level + 1 >= net.nlevels) break;
            Frame child;                         // algorithm.jl:32-39: the full recursive solve of the level below
            child.level = f.level + 1;
            C.stack.push_back(std::move(child));
        }
        post_verify(C);
    }

    // ---- process_qp, first phase (memoised: GeoCache::verify_plan) -------------------------------------------------------
    void post_verify(Cohort& C) {
        Frame& f = C.stack.back();
        f.plan = c.verify_plan(f.level, f.S, w);
        const VerifyPlan& P = c.plan(f.plan);
        if (P.error) return fail(C, P.error);
        C.wait = W_VERIFY;
    }

    // algorithm.jl:104-116: the level is done; its graphs go up
    void finish_level(Cohort& C, int S_new) {
        C.stack.back().S = c.list(S_new);
        if (C.stack.size() == 1) return finish(C, true, 0);
        std::vector<int> S = std::move(C.stack.back().S);
        C.stack.pop_back();
        C.stack.back().S = std::move(S);         // S = ret_low.Sol; x = ret_low.x_opt (x is resident: nothing to copy)
        post_verify(C);
    }

    // ---- the request a waiting cohort posts, and what it does with a representative's answers ---------------------------
    Post make_post(Cohort& C) {
        Frame& f = C.stack.back();
        Post p;
        p.seg = C.seg;
        switch (C.wait) {
            case W_VERIFY: {
                const VerifyPlan& P = c.plan(f.plan);
                p.kind = POST_VERIFY; p.nodes = P.req_nodes.data(); p.nnodes = (int)P.req_nodes.size(); p.want_vertices = P.want;
                p.snap = f.level == 0; p.cyc_level = C.cyc_level; p.ncyc = C.ncyc;
                break;
            }
            case W_MEMBER: {
                const VerifyOutcome& O = c.outcome(f.outcome);
                p.kind = POST_MEMBER; p.piece_lists = O.comb_lists.data(); p.nlists = (int)O.comb_lists.size();
                break;
            }
            default: {
                p.kind = POST_QEP; p.gavi = c.outcome(f.outcome).gavi; p.snap = f.level == 0;
                // after a successful solve: start_iter() at this level, i.e. the cycle checks of levels f.level .. last and the
                // bottom level's verify request -- posted now, answered in the same round
                const VerifyPlan& P = c.plan(bottom_plan);
                if (!P.error) {
                    p.nodes = P.req_nodes.data(); p.nnodes = (int)P.req_nodes.size(); p.want_vertices = P.want;
                    p.vsnap = net.nlevels == 1;
                    if (net.check_for_cycling && net.num_projections > 0) { p.cyc_level = f.level; p.ncyc = net.nlevels - f.level; }
                }
                break;
            }
        }
        return p;
    }

    void apply(Cohort& P, const uint8_t* rep) {
        const Wait wt = P.wait;
        P.wait = W_NONE;
        Frame& f = P.stack.back();
        if (wt == W_MEMBER) return finish_level(P, c.member_outcome(f.outcome, rep, w));
        if (wt == W_QEP) {
            int32_t status;
            std::memcpy(&status, rep, 4);
            // the level and the solver's StatusCode ride in the upper bytes of the error word (triage of unsolved instances)
            if (status != 1) return fail(P, ERR_AVI | ((status & 0xff) << 8) | ((f.level & 0xff) << 16));
            if (rep[4] == 0) return fail(P, ERR_DISAGREE);           // algorithm.jl:96-97
            start_iter(P);
            // the verify request start_iter() arrived at was answered in the same round (make_post)
            if (P.done || P.wait != W_VERIFY || P.stack.back().plan != bottom_plan || c.plan(bottom_plan).error) return;
            P.wait = W_NONE;
            return apply_verify(P, rep + 8);
        }
        apply_verify(P, rep);
    }

    void apply_verify(Cohort& P, const uint8_t* rep) {
        Frame& f = P.stack.back();
        if (rep[0] != 0) {                       // a cycle check hit at level rep[0] - 1: the passes opened below it never began
            for (int l = rep[0]; l < P.cyc_level + P.ncyc; ++l) P.level_iters[l]--;
            return fail(P, ERR_CYCLE);
        }
        P.ncyc = 0;
        f.outcome = c.verify_outcome(f.plan, rep + 1, w);
        const VerifyOutcome& O = c.outcome(f.outcome);
        switch (O.kind) {
            case VerifyOutcome::FAIL: return fail(P, O.error);
            case VerifyOutcome::QEP: P.wait = W_QEP; return;
            case VerifyOutcome::MEMBER: P.wait = W_MEMBER; return;
            default: return finish_level(P, O.S_new);
        }
    }
};

}  // namespace

// =================================================================================================================
// NetSolver
// =================================================================================================================
NetSolver::NetSolver(NetData net, std::vector<Poly> polys, std::unique_ptr<Store> store)
    : net_(std::move(net)), store_(std::move(store)) {
    cache_.reset(new GeoCache(net_, store_.get()));
    workers_.emplace_back(store_->make_worker());
    std::vector<int> ids;
    for (Poly& P : polys) ids.push_back(cache_->intern_poly(std::move(P), workers_[0].get()));
    for (auto& b : net_.base) for (int& k : b) k = ids[k];
}
NetSolver::~NetSolver() { workers_.clear(); cache_.reset(); store_.reset(); }

void NetSolver::run_shard(int tid, int lo, int hi, const double* inits, double* x_out, int* result_of, std::vector<SolveOut>& results) {
    const int B = hi - lo, nv = net_.nv;
    if (B <= 0) return;
    Worker* w = workers_[tid].get();
    Stats& st = cache_->stats;
    auto now = []() { return std::chrono::steady_clock::now(); };
    auto ns = [](auto a, auto b) { return (long)std::chrono::duration_cast<std::chrono::nanoseconds>(b - a).count(); };
    auto t_host = now();
    w->set_batch(B, inits + (size_t)lo * nv);
    st.backend_ns += ns(t_host, now());
    t_host = now();
    Machine M(*cache_, w);
    {
        auto C = std::make_unique<Cohort>();
        C->seg = Seg{0, B};
        C->level_iters.assign(net_.nlevels, 0);
        C->stack.emplace_back();
        M.start_iter(*C);
        M.ready.push_back(std::move(C));
    }
    results.clear();                             // outcome per finished cohort; instances carry an index into it
    std::vector<std::unique_ptr<Cohort>> waiting;
    std::vector<Part> parts;
    while (true) {
        // ---- everything the host can decide on its own --------------------------------------------------------------
        while (!M.ready.empty()) {
            std::unique_ptr<Cohort> C = std::move(M.ready.back());
            M.ready.pop_back();
            if (C->done) {
                SolveOut o;
                o.solved = C->solved; o.error = C->error; o.level_iters = C->level_iters;
                if (C->solved) o.sol = C->sol; else o.sol.assign(net_.nplayers, -1);
                w->mark_done(C->seg, (int)results.size());
                results.push_back(std::move(o));
            } else {
                waiting.push_back(std::move(C));
            }
        }
        if (waiting.empty()) break;
        // ---- one round of the backend over the requests of all waiting cohorts -------------------------------------------
        long nreq = 0;
        for (auto& C : waiting) {
            const Post p = M.make_post(*C);
            nreq += (long)p.seg.n * (p.kind == POST_VERIFY ? p.nnodes : p.kind == POST_MEMBER ? p.nlists : 1);
            w->post(p);
        }
        auto t_back = now();
        st.host_ns += ns(t_host, t_back);
        st.rounds++;
        st.requests += nreq;
        st.calls += (long)waiting.size();
        w->finish_round(parts);
        t_host = now();
        st.backend_ns += ns(t_back, t_host);
        // ---- every part of a cohort goes on with a copy of the state and its representative's answers -----------------------
        std::vector<std::unique_ptr<Cohort>> got;
        got.swap(waiting);
        const auto t_apply = now();
        for (size_t i = 0; i < parts.size();) {
            size_t j = i;
            while (j < parts.size() && parts[j].cohort == parts[i].cohort) ++j;
            std::unique_ptr<Cohort>& C = got[parts[i].cohort];
            if (j - i > 1) st.cohorts += (long)(j - i - 1);
            for (size_t k = i + 1; k < j; ++k) {
                auto D = std::make_unique<Cohort>(*C);
                D->seg = parts[k].seg;
                M.apply(*D, parts[k].rep);
                M.ready.push_back(std::move(D));
            }
            C->seg = parts[i].seg;
            M.apply(*C, parts[i].rep);
            M.ready.push_back(std::move(C));
            i = j;
        }
        st.apply_ns += ns(t_apply, now());
    }
    st.host_ns += ns(t_host, now());
    t_host = now();
    std::vector<uint8_t> solved(results.size() + 1, 0);
    for (size_t r = 0; r < results.size(); ++r) solved[r] = results[r].solved;
    std::vector<int32_t> res_of(B, -1);
    w->download(x_out + (size_t)lo * nv, res_of.data(), solved.data(), (int)results.size());
    // (an instance no cohort finished -- cannot happen unless the backend failed -- reads as unsolved)
    int lost = -1;
    for (int b = 0; b < B; ++b) {
        if (res_of[b] >= 0 && res_of[b] < (int)results.size()) { result_of[b] = res_of[b]; continue; }
        if (lost < 0) {
            SolveOut o;
            o.level_iters.assign(net_.nlevels, 0); o.sol.assign(net_.nplayers, -1); o.error = ERR_MAXIT;
            lost = (int)results.size();
            results.push_back(std::move(o));
        }
        result_of[b] = lost;
    }
    st.backend_ns += ns(t_host, now());
}

void NetSolver::solve_batched(int B, const double* inits, double* x_out, std::vector<int>& result_of, std::vector<SolveOut>& results, int threads) {
    result_of.assign(B > 0 ? B : 0, -1);
    results.clear();
    if (B <= 0) return;
    if (threads < 1) threads = 1;
    if (threads > B) threads = B;
    while ((int)workers_.size() < threads) workers_.emplace_back(store_->make_worker());
    if (threads == 1) { run_shard(0, 0, B, inits, x_out, result_of.data(), results); return; }
    std::vector<std::vector<SolveOut>> shard_results(threads);
    std::vector<std::pair<int, int>> range(threads);
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) {
        const int base = B / threads, rem = B % threads;
        const int lo = t * base + std::min(t, rem), hi = lo + base + (t < rem ? 1 : 0);
        range[t] = {lo, hi};
        th.emplace_back([this, t, lo, hi, inits, x_out, &result_of, &shard_results]() {
            run_shard(t, lo, hi, inits, x_out, result_of.data() + lo, shard_results[t]);
        });
    }
    for (auto& t : th) t.join();
    for (int t = 0; t < threads; ++t) {          // one table: the shards' outcomes back to back
        const int off = (int)results.size();
        for (auto& r : shard_results[t]) results.push_back(std::move(r));
        if (off) for (int b = range[t].first; b < range[t].second; ++b) result_of[b] += off;
    }
}

}  // namespace qpnnet
