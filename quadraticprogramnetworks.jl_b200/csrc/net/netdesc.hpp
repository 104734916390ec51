// qpn_net_desc (include/qpn_cuda.h) -> NetData, and the C entry points shared by the CUDA library and the oracle build.
#pragma once
#include <algorithm>
#include <set>

#include "../../../include/qpn_cuda.h"
#include "netsolver.hpp"

namespace qpnnet {

// Builds NetData; the constraint polys are interned through `intern` in their given order (poly k gets id k).
inline NetData net_from_desc(const qpn_net_desc* d, std::vector<Poly>& polys_out) {
    NetData n;
    n.nv = d->nv; n.nplayers = d->nplayers; n.nlevels = d->nlevels;
    const int nv = d->nv, np = d->nplayers;
    n.Q.resize(np); n.q.resize(np); n.base.resize(np); n.children.resize(np); n.dec.resize(np);
    n.level_of.assign(d->level_of, d->level_of + np);
    n.levels.assign(d->nlevels, {});
    std::vector<std::vector<int>> own(np);
    for (int p = 0; p < np; ++p) {
        n.Q[p].assign(d->Q + (size_t)p * nv * nv, d->Q + (size_t)(p + 1) * nv * nv);
        n.q[p].assign(d->q + (size_t)p * nv, d->q + (size_t)(p + 1) * nv);
        own[p].assign(d->var_idx + d->var_ptr[p], d->var_idx + d->var_ptr[p + 1]);
        n.base[p].assign(d->con_idx + d->con_ptr[p], d->con_idx + d->con_ptr[p + 1]);
        n.children[p].assign(d->child_idx + d->child_ptr[p], d->child_idx + d->child_ptr[p + 1]);
        std::sort(n.children[p].begin(), n.children[p].end());
        n.levels[n.level_of[p]].push_back(p);
    }
    for (auto& l : n.levels) std::sort(l.begin(), l.end());
    // decision_inds (programs.jl:340-346): own variables and those of every reachable node
    for (int p = 0; p < np; ++p) {
        std::set<int> inds(own[p].begin(), own[p].end());
        std::vector<int> stack = n.children[p];
        std::vector<char> seen(np, 0);
        while (!stack.empty()) {
            const int j = stack.back();
            stack.pop_back();
            if (seen[j]) continue;
            seen[j] = 1;
            inds.insert(own[j].begin(), own[j].end());
            for (int k : n.children[j]) stack.push_back(k);
        }
        n.dec[p].assign(inds.begin(), inds.end());
    }
    n.max_iters = d->max_iters; n.num_projections = d->num_projections; n.exploration_vertices = d->exploration_vertices;
    n.gen_solution_map = d->gen_solution_map; n.check_for_cycling = d->check_for_cycling;
    n.remove_subsets_at.assign(d->nlevels, 1);
    if (d->remove_subsets_at) for (int l = 0; l < d->nlevels; ++l) n.remove_subsets_at[l] = (char)d->remove_subsets_at[l];
    if (d->num_projections > 0 && d->proj) n.proj.assign(d->proj, d->proj + (size_t)d->num_projections * nv);
    polys_out.clear();
    for (int k = 0; k < d->npolys; ++k) {
        Rows r;
        r.d = nv;
        for (int i = d->poly_ptr[k]; i < d->poly_ptr[k + 1]; ++i) r.add(d->poly_A + (size_t)i * nv, d->poly_l[i], d->poly_u[i]);
        polys_out.push_back(make_poly(std::move(r), false));
    }
    return n;
}

// One net object behind the C ABI.
struct NetObject {
    std::unique_ptr<NetSolver> solver;
    std::vector<int> result_of;                  // per instance of the last batch: index into `results`
    std::vector<SolveOut> results;
    const SolveOut* out(int b) const { return (b >= 0 && b < (int)result_of.size() && result_of[b] >= 0) ? &results[result_of[b]] : nullptr; }
    int threads = 2;                             // host threads (streams) that drive a batch: two hide each other's round trips, more add nothing
    std::string err;
    std::function<int64_t()> launches;
};

inline int net_solve(NetObject* o, int batch, const double* inits, double* x_out, uint8_t* solved_out, int32_t* level_iters_out,
                     int32_t* error_out) {
    o->solver->solve_batched(batch, inits, x_out, o->result_of, o->results, o->threads);
    const int nl = o->solver->net().nlevels;
    for (int b = 0; b < batch; ++b) {
        const SolveOut& r = o->results[o->result_of[b]];
        if (solved_out) solved_out[b] = r.solved;
        if (level_iters_out) for (int l = 0; l < nl; ++l) level_iters_out[(size_t)b * nl + l] = r.level_iters[l];
        if (error_out) error_out[b] = r.error;
    }
    return 0;
}
inline int net_sol_count(NetObject* o, int b, int player) {
    const SolveOut* r = o->out(b);
    if (!r || player < 0 || player >= (int)r->sol.size()) return -1;
    const int lid = r->sol[player];
    return lid < 0 ? -1 : (int)o->solver->cache().list(lid).size();
}
inline int net_sol_piece(NetObject* o, int b, int player, int k) {
    const int n = net_sol_count(o, b, player);
    if (k < 0 || k >= n) return -1;
    return o->solver->cache().list(o->out(b)->sol[player])[k];
}
inline int net_piece_rows(NetObject* o, int piece) {
    if (piece < 0 || piece >= o->solver->cache().npolys()) return -1;
    return o->solver->cache().poly(piece).m();
}
inline int net_piece_get(NetObject* o, int piece, double* A, double* l, double* u, uint8_t* rl, uint8_t* ru) {
    if (piece < 0 || piece >= o->solver->cache().npolys()) return -1;
    const Poly& P = o->solver->cache().poly(piece);
    if (A) std::copy(P.A.begin(), P.A.end(), A);
    if (l) std::copy(P.l.begin(), P.l.end(), l);
    if (u) std::copy(P.u.begin(), P.u.end(), u);
    if (rl) std::copy(P.rl.begin(), P.rl.end(), rl);
    if (ru) std::copy(P.ru.begin(), P.ru.end(), ru);
    return 0;
}
inline int net_stats(NetObject* o, int64_t* out) {
    Stats& s = o->solver->cache().stats;
    out[0] = o->launches ? o->launches() : 0;
    out[1] = s.rounds; out[2] = s.requests; out[3] = s.calls; out[4] = s.lps; out[5] = s.pieces; out[6] = s.nodes; out[7] = s.gavis;
    out[8] = s.collect_miss; out[9] = s.combine_miss; out[10] = s.host_ns; out[11] = s.backend_ns; out[12] = s.cohorts; out[13] = s.apply_ns; out[14] = s.lps_empty; out[15] = s.lp_calls;
    return 0;
}

}  // namespace qpnnet
