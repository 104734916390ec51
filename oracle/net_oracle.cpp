// ORACLE (test infrastructure / CPU baseline): the native network state machine of
// quadraticprogramnetworks.jl_b200/csrc/net/ with every numeric request served by the C oracle (qpn_oracle.c)
// on the host -- the same host logic as the product, none of its kernels.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs load the library this file builds
// (oracle/_build/libqpn_net_oracle.so).  Parity status of the numerics: see the header of qpn_oracle.c / DESIGN.md
// (pinned by the reference's simple_bilevel known answers; unpinned against PATH / OSQP themselves).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <unordered_map>

#include "../quadraticprogramnetworks.jl_b200/csrc/net/netdesc.hpp"
#include "../quadraticprogramnetworks.jl_b200/csrc/net/vertex_enum.h"
#include "../quadraticprogramnetworks.jl_b200/csrc/net/cycle_check.h"

extern "C" {
int qpo_gavi_solve(int d1, int d2, int np, const double* M, const double* N, const double* o, const double* l1, const double* u1,
                   const double* A, const double* B, const double* l2, const double* u2, const double* w, double* z0, int presolve,
                   int max_pivots, double* z_out, double* zfull_out, int32_t* status, int32_t* pivots, int8_t* basis);
void qpo_comp_indices(int d1, int d2, int np, const double* M, const double* N, const double* o, const double* l1, const double* u1,
                      const double* A, const double* B, const double* l2, const double* u2, const double* z, const double* w,
                      double tol, int8_t* mask);
int qpo_halfspace_in(int m, int d, const double* A, const double* l, const double* u, const uint8_t* rl, const uint8_t* ru,
                     const double* x, double tol);
int qpo_verify_solution(int nd, int nv, int m, const double* Qd, const double* qd, const double* A, const double* l, const double* u,
                        const int32_t* dec, const double* x, double tol, double* lam_out, int32_t* how, int8_t* active,
                        int32_t* fallback_pivots);
}

namespace {
using namespace qpnnet;

struct OracleStore;

struct PieceCM { int m = 0; std::vector<double> A, l, u; };      // column-major copy for qpo_halfspace_in

struct OracleWorker : Worker {
    OracleStore* store;
    const NetData* net = nullptr;            // set at set_batch (the store learns the net after the solver is built)
    const GeoCache* cache = nullptr;
    int B = 0, nv = 0, nproj = 0;
    std::vector<double> X, Xf, PV;
    std::vector<int> order, next_order;
    std::vector<std::vector<std::vector<double>>> hist;      // [slot][level]: flat list of earlier projections
    std::vector<int32_t> result_of;
    std::vector<Post> posts;
    std::deque<std::vector<uint8_t>> reps;   // representative answers of the current round
    explicit OracleWorker(OracleStore* s) : store(s) {}

    // (test hook: QPN_ORACLE_WIDE=1 makes the host backend take the wide-batch path of the set algebra, which the CUDA
    // backend always takes, so that the CPU suite can check that it changes no result)
    bool wide_batches() const override { static const bool w = getenv("QPN_ORACLE_WIDE") != nullptr; return w; }
    int gavi_solve_one(const GaviData& g, const double* w, const double* z0, double* z) override {
        std::vector<double> z0c(z0, z0 + g.d1 + g.d2);
        static const double none = 0.0;
        int32_t st = 0, pv = 0;
        qpo_gavi_solve(g.d1, g.d2, g.np, g.M.data(), g.N.data(), g.o.data(), g.l1.data(), g.u1.data(), g.A.data(), g.B.data(),
                       g.l2.data(), g.u2.data(), w ? w : &none, z0c.data(), 1, 0, z, nullptr, &st, &pv, nullptr);
        return st;
    }
    void project(int slot) {
        for (int k = 0; k < nproj; ++k) {
            double acc = 0.0;
            for (int j = 0; j < nv; ++j) acc = std::fma(X[(size_t)slot * nv + j], net->proj[(size_t)k * nv + j], acc);
            PV[(size_t)slot * nproj + k] = acc;
        }
    }
    void set_batch(int B_, const double* x_init) override;
    int post(const Post& p) override { posts.push_back(p); return (int)posts.size() - 1; }
    void mark_done(Seg seg, int result) override { for (int k = 0; k < seg.n; ++k) result_of[order[seg.off + k]] = result; }

    // the answers of one member, in the byte layout of Part::rep (netsolver.hpp)
    // the chain of cycle checks a verify request begins with: 0 = no hit, 1 + level of the first hit
    uint8_t cycle_chain(const Post& p, int slot) {
        const double* pv = PV.data() + (size_t)slot * nproj;
        for (int level = p.cyc_level; level < p.cyc_level + p.ncyc; ++level) {
            std::vector<double>& h = hist[slot][level];
            for (size_t q = 0; q + nproj <= h.size(); q += nproj) if (qpn_cycle_hit(pv, h.data() + q, nproj)) return (uint8_t)(1 + level);
            h.insert(h.end(), pv, pv + nproj);
        }
        return 0;
    }
    void answer_verify(const Post& p, int slot, uint8_t* out, int snap) {
        const double* x = X.data() + (size_t)slot * nv;
        const int want = p.want_vertices;
        *out++ = cycle_chain(p, slot);
        for (int r = 0; r < p.nnodes; ++r) {
            const NodeInfo& n = cache->node_info(p.nodes[r]);
            const GaviData& g = n.g;
            const int dz = n.nd + n.m, vb = (n.m + 1) / 2;
            std::vector<double> lam(n.m + 1), z(dz + 1), w(n.par.size() + 1), ax(n.m + 1), qt(n.nd + 1), V((size_t)QPN_VE_MAXV * QPN_VE_MAXA), zv(dz + 1);
            std::vector<int8_t> mv(dz + 1);
            int8_t* mask = (int8_t*)out + 1;
            uint8_t* vcount = out + 1 + dz;
            uint8_t* vmask = out + 2 + dz;
            int32_t how = 0, fp = 0;
            out[0] = (uint8_t)qpo_verify_solution(n.nd, n.nv, n.m, n.Qd.data(), n.qd.data(), n.A.data(), n.l.data(), n.u.data(), n.dec.data(), x,
                                                  1e-4, lam.data(), &how, nullptr, &fp);
            if (out[0]) {
                for (int e = 0; e < n.nd; ++e) z[e] = x[n.dec[e]];
                for (int i = 0; i < n.m; ++i) z[n.nd + i] = lam[i];
                for (size_t c = 0; c < n.par.size(); ++c) w[c] = x[n.par[c]];
                qpo_comp_indices(g.d1, g.d2, g.np, g.M.data(), g.N.data(), g.o.data(), g.l1.data(), g.u1.data(), g.A.data(), g.B.data(),
                                 g.l2.data(), g.u2.data(), z.data(), w.data(), 1e-2, mask);
                if (want > 0) {
                    // expand's get_verts (avi_solutions.jl:252-255): vertices of the multiplier polytope at x, then comp_indices there
                    for (int e = 0; e < n.nd; ++e) {
                        double acc = 0.0;
                        for (int j = 0; j < nv; ++j) acc = std::fma(n.Qd[(size_t)j * n.nd + e], x[j], acc);
                        qt[e] = acc + n.qd[e];
                    }
                    for (int i = 0; i < n.m; ++i) {
                        double acc = 0.0;
                        for (int j = 0; j < nv; ++j) acc = std::fma(n.A[(size_t)j * n.m + i], x[j], acc);
                        ax[i] = acc;
                    }
                    int idxA[QPN_VE_MAXA], na = 0;
                    const int nvx = qpn_multiplier_vertices(n.nd, n.m, nv, n.A.data(), n.dec.data(), n.l.data(), n.u.data(), ax.data(), qt.data(),
                                                            lam.data(), want, idxA, &na, V.data());
                    *vcount = (uint8_t)nvx;
                    for (int q = 0; q < nvx; ++q) {
                        for (int e = 0; e < n.nd; ++e) zv[e] = z[e];
                        for (int i = 0; i < n.m; ++i) zv[n.nd + i] = 0.0;
                        for (int j = 0; j < na; ++j) zv[n.nd + idxA[j]] = V[(size_t)q * QPN_VE_MAXA + j];
                        qpo_comp_indices(g.d1, g.d2, g.np, g.M.data(), g.N.data(), g.o.data(), g.l1.data(), g.u1.data(), g.A.data(), g.B.data(),
                                         g.l2.data(), g.u2.data(), zv.data(), w.data(), 1e-2, mv.data());
                        uint8_t* nib = vmask + (size_t)q * vb;
                        for (int i = 0; i < n.m; ++i) nib[i >> 1] |= (uint8_t)((mv[n.nd + i] & 0xf) << ((i & 1) * 4));
                    }
                }
            }
            if (snap) std::memcpy(Xf.data() + (size_t)slot * nv, x, sizeof(double) * nv);
            out += verify_rep_bytes(dz, n.m, want);
        }
    }
    void answer_qep(const Post& p, int slot, uint8_t* out) {
        const LevelGaviInfo& L = cache->gavi_info(p.gavi);
        const GaviData& g = L.g;
        const int dz = g.d1 + g.d2, ndl = (int)L.dec.size();
        std::vector<double> w(g.np + 1), z0(dz + 1), z(dz + 1), xn(nv);
        double* x = X.data() + (size_t)slot * nv;
        for (int j = 0; j < g.np; ++j) w[j] = x[L.par[j]];
        for (int j = 0; j < dz; ++j) z0[j] = j < ndl ? x[L.dec[j]] : 0.0;
        int32_t status = 0, pivots = 0;
        uint8_t moved = 0;
        qpo_gavi_solve(g.d1, g.d2, g.np, g.M.data(), g.N.data(), g.o.data(), g.l1.data(), g.u1.data(), g.A.data(), g.B.data(), g.l2.data(),
                       g.u2.data(), w.data(), z0.data(), 1, 0, z.data(), nullptr, &status, &pivots, nullptr);
        if (status == 1) {
            std::memcpy(xn.data(), x, sizeof(double) * nv);
            for (int j = 0; j < ndl; ++j) xn[L.dec[j]] = z[j];
            double dn = 0.0;
            for (int j = 0; j < nv; ++j) { const double e = xn[j] - x[j]; dn = std::fma(e, e, dn); }
            moved = !(std::sqrt(dn) < 1e-4);
            if (moved) {
                std::memcpy(x, xn.data(), sizeof(double) * nv);
                project(slot);
            }
        }
        if (p.snap) std::memcpy(Xf.data() + (size_t)slot * nv, x, sizeof(double) * nv);
        std::memcpy(out, &status, 4);
        out[4] = moved;
        // the verify request that rides along (netsolver.hpp): answered where the solve succeeded and moved
        if (p.nnodes > 0 && status == 1 && moved) answer_verify(p, slot, out + 8, p.vsnap);
    }
    void answer_member(const Post& p, int slot, uint8_t* out);

    void finish_round(std::vector<Part>& parts) override {
        parts.clear();
        reps.clear();
        next_order.clear();
        std::vector<uint8_t> rows;
        for (size_t ci = 0; ci < posts.size(); ++ci) {
            const Post& p = posts[ci];
            size_t rb = 0;
            if (p.kind == POST_QEP) {
                rb = 8;
                if (p.nnodes > 0) {
                    rb += 1;
                    for (int r = 0; r < p.nnodes; ++r) { const NodeInfo& n = cache->node_info(p.nodes[r]); rb += verify_rep_bytes(n.nd + n.m, n.m, p.want_vertices); }
                }
            }
            else if (p.kind == POST_MEMBER) { for (int k = 0; k < p.nlists; ++k) rb += p.piece_lists[k]->size(); }
            else {
                rb = 1;
                for (int r = 0; r < p.nnodes; ++r) { const NodeInfo& n = cache->node_info(p.nodes[r]); rb += verify_rep_bytes(n.nd + n.m, n.m, p.want_vertices); }
            }
            const size_t stride = rb ? rb : 1;
            rows.assign(stride * p.seg.n, 0);
            for (int k = 0; k < p.seg.n; ++k) {
                const int slot = order[p.seg.off + k];
                uint8_t* out = rows.data() + stride * k;
                if (p.kind == POST_VERIFY) answer_verify(p, slot, out, p.snap);
                else if (p.kind == POST_QEP) answer_qep(p, slot, out);
                else answer_member(p, slot, out);
            }
            // partition by the answers, parts in order of first appearance, members in their old order
            std::unordered_map<std::string, int> part_of;
            std::vector<std::vector<int>> members;
            for (int k = 0; k < p.seg.n; ++k) {
                std::string key((const char*)rows.data() + stride * k, stride);
                auto it = part_of.find(key);
                if (it == part_of.end()) { it = part_of.emplace(std::move(key), (int)members.size()).first; members.emplace_back(); }
                members[it->second].push_back(k);
            }
            for (auto& mem : members) {
                Part pt;
                pt.cohort = (int)ci;
                pt.seg = Seg{(int)next_order.size(), (int)mem.size()};
                reps.emplace_back(rows.begin() + stride * mem[0], rows.begin() + stride * (mem[0] + 1));
                pt.rep = reps.back().data();
                for (int k : mem) next_order.push_back(order[p.seg.off + k]);
                parts.push_back(pt);
            }
        }
        order.swap(next_order);
        posts.clear();
    }
    void download(double* x_out, int32_t* result_of_slot, const uint8_t* solved_of_result, int nresults) override {
        for (int b = 0; b < B; ++b) {
            const int r = result_of[b];
            result_of_slot[b] = r;
            const bool solved = r >= 0 && r < nresults && solved_of_result[r];
            std::memcpy(x_out + (size_t)b * nv, (solved ? X.data() : Xf.data()) + (size_t)b * nv, sizeof(double) * nv);
        }
    }
};

struct OracleStore : Store {
    const NetData* net = nullptr;
    const GeoCache* cache = nullptr;
    std::deque<PieceCM> pieces;
    std::shared_mutex mu;
    Worker* make_worker() override { return new OracleWorker(this); }
    void new_node(int, const NodeInfo&, Worker*) override {}
    void new_gavi(int, const LevelGaviInfo&, Worker*) override {}
    void new_piece(int id, const Poly& P, Worker*) override {
        PieceCM c;
        c.m = P.m();
        c.A.assign((size_t)P.m() * P.d, 0.0);
        for (int i = 0; i < P.m(); ++i) for (int j = 0; j < P.d; ++j) c.A[(size_t)j * P.m() + i] = P.row(i)[j];
        c.l = P.l; c.u = P.u;
        std::unique_lock<std::shared_mutex> lk(mu);
        if ((int)pieces.size() <= id) pieces.resize(id + 1);
        pieces[id] = std::move(c);
    }
};

void OracleWorker::set_batch(int B_, const double* x_init) {
    net = store->net;
    cache = store->cache;
    B = B_; nv = net->nv;
    nproj = net->check_for_cycling ? net->num_projections : 0;
    X.assign(x_init, x_init + (size_t)B * nv);
    Xf = X;
    PV.assign((size_t)B * (nproj > 0 ? nproj : 1), 0.0);
    for (int b = 0; b < B; ++b) project(b);
    order.resize(B);
    for (int b = 0; b < B; ++b) order[b] = b;
    hist.assign(B, std::vector<std::vector<double>>(net->nlevels));
    result_of.assign(B, -1);
    posts.clear();
}
void OracleWorker::answer_member(const Post& p, int slot, uint8_t* out) {
    const double* x = X.data() + (size_t)slot * nv;
    for (int q = 0; q < p.nlists; ++q)
        for (int id : *p.piece_lists[q]) {
            const PieceCM* c;
            { std::shared_lock<std::shared_mutex> lk(store->mu); c = &store->pieces[id]; }
            *out++ = c->m == 0 ? 1 : (uint8_t)qpo_halfspace_in(c->m, nv, c->A.data(), c->l.data(), c->u.data(), nullptr, nullptr, x, 1e-6);
        }
}
}  // namespace

using qpnnet::NetObject;

extern "C" {
int qpo_net_create(const qpn_net_desc* desc, void** out) {
    if (!desc || !out) return -1;
    std::vector<qpnnet::Poly> polys;
    qpnnet::NetData nd = qpnnet::net_from_desc(desc, polys);
    auto store = std::make_unique<OracleStore>();
    OracleStore* sp = store.get();
    NetObject* o = new NetObject();
    o->solver.reset(new qpnnet::NetSolver(std::move(nd), std::move(polys), std::move(store)));
    sp->net = &o->solver->net();
    sp->cache = &o->solver->cache();
    *out = o;
    return 0;
}
int qpo_net_destroy(void* net) { delete (NetObject*)net; return 0; }
int qpo_net_set_option(void* net, const char* name, int64_t v) {
    if (!net || !name) return -1;
    if (!std::strcmp(name, "threads")) { ((NetObject*)net)->threads = (int)v; return 0; }
    return -1;
}
int qpo_net_solve_batched(void* net, int batch, const double* inits, double* x_out, uint8_t* solved_out, int32_t* level_iters_out,
                          int32_t* error_out) {
    return qpnnet::net_solve((NetObject*)net, batch, inits, x_out, solved_out, level_iters_out, error_out);
}
int qpo_net_sol_count(void* net, int b, int player) { return qpnnet::net_sol_count((NetObject*)net, b, player); }
int qpo_net_sol_piece(void* net, int b, int player, int k) { return qpnnet::net_sol_piece((NetObject*)net, b, player, k); }
int qpo_net_piece_rows(void* net, int piece) { return qpnnet::net_piece_rows((NetObject*)net, piece); }
int qpo_net_piece_get(void* net, int piece, double* A, double* l, double* u, uint8_t* rl, uint8_t* ru) {
    return qpnnet::net_piece_get((NetObject*)net, piece, A, l, u, rl, ru);
}
int qpo_net_stats(void* net, int64_t* out) { return qpnnet::net_stats((NetObject*)net, out); }
}
