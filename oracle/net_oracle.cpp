// ORACLE (test infrastructure / CPU baseline): the native network state machine of
// quadraticprogramnetworks.jl_b200/csrc/net/ with every numeric request served by the C oracle (qpn_oracle.c)
// on the host -- the same host logic as the product, none of its kernels.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs load the library this file builds
// (oracle/_build/libqpn_net_oracle.so).  Parity status of the numerics: see the header of qpn_oracle.c / DESIGN.md
// (pinned by the reference's simple_bilevel known answers; unpinned against PATH / OSQP themselves).
#include <cmath>
#include <cstring>

#include "../quadraticprogramnetworks.jl_b200/csrc/net/netdesc.hpp"
#include "../quadraticprogramnetworks.jl_b200/csrc/net/vertex_enum.h"

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

struct OracleWorker : Worker {
    OracleStore* store;
    const NetData* net = nullptr;            // set at set_batch (the store learns the net after the solver is built)
    int B = 0, nv = 0;
    std::vector<double> X, Xf;
    explicit OracleWorker(OracleStore* s) : store(s) {}

    int gavi_solve_one(const GaviData& g, const double* w, const double* z0, double* z) override {
        std::vector<double> z0c(z0, z0 + g.d1 + g.d2);
        static const double none = 0.0;
        int32_t st = 0, pv = 0;
        qpo_gavi_solve(g.d1, g.d2, g.np, g.M.data(), g.N.data(), g.o.data(), g.l1.data(), g.u1.data(), g.A.data(), g.B.data(),
                       g.l2.data(), g.u2.data(), w ? w : &none, z0c.data(), 1, 0, z, nullptr, &st, &pv, nullptr);
        return st;
    }
    void set_batch(int B_, const double* x_init) override;
    // result storage of the current round (the batches point into these)
    std::deque<std::vector<uint8_t>> u8;
    std::deque<std::vector<int8_t>> i8;
    std::deque<std::vector<int32_t>> i32;
    std::deque<std::vector<double>> f64;
    bool fresh_round = true;
    void begin_round() { if (fresh_round) { u8.clear(); i8.clear(); i32.clear(); f64.clear(); fresh_round = false; } }

    void run_verify(int, const NodeInfo& n, VerifyBatch** bs, int nb, bool snap) override {
        begin_round();
        std::vector<double> lam(n.m + 1), z(n.nd + n.m + 1), w(n.par.size() + 1);
        const GaviData& g = n.g;
        const int dz = n.nd + n.m;
        for (int q = 0; q < nb; ++q) {
            VerifyBatch& b = *bs[q];
            u8.emplace_back(b.n, 0);
            i8.emplace_back((size_t)b.n * dz, 0);
            uint8_t* sol = u8.back().data();
            int8_t* mask = i8.back().data();
            const int want = b.want_vertices > QPN_VE_MAXV ? QPN_VE_MAXV : b.want_vertices;
            const int vb = (n.m + 1) / 2, vstride = want * vb;
            uint8_t *vcount = nullptr, *vmask = nullptr;
            if (want > 0) {
                u8.emplace_back(b.n, 0); vcount = u8.back().data();
                u8.emplace_back((size_t)b.n * vstride + 1, 0); vmask = u8.back().data();
            }
            std::vector<double> ax(n.m + 1), qt(n.nd + 1), V((size_t)QPN_VE_MAXV * QPN_VE_MAXA), zv(dz + 1);
            std::vector<int8_t> mv(dz + 1);
            for (int k = 0; k < b.n; ++k) {
                const double* x = X.data() + (size_t)b.slots[k] * nv;
                int32_t how = 0, fp = 0;
                sol[k] = (uint8_t)qpo_verify_solution(n.nd, n.nv, n.m, n.Qd.data(), n.qd.data(), n.A.data(), n.l.data(), n.u.data(),
                                                      n.dec.data(), x, 1e-4, lam.data(), &how, nullptr, &fp);
                if (sol[k]) {
                    for (int e = 0; e < n.nd; ++e) z[e] = x[n.dec[e]];
                    for (int i = 0; i < n.m; ++i) z[n.nd + i] = lam[i];
                    for (size_t c = 0; c < n.par.size(); ++c) w[c] = x[n.par[c]];
                    qpo_comp_indices(g.d1, g.d2, g.np, g.M.data(), g.N.data(), g.o.data(), g.l1.data(), g.u1.data(), g.A.data(),
                                     g.B.data(), g.l2.data(), g.u2.data(), z.data(), w.data(), 1e-2, mask + (size_t)k * dz);
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
                        const int nvx = qpn_multiplier_vertices(n.nd, n.m, nv, n.A.data(), n.dec.data(), n.l.data(), n.u.data(), ax.data(),
                                                                qt.data(), lam.data(), want, idxA, &na, V.data());
                        vcount[k] = (uint8_t)nvx;
                        for (int q = 0; q < nvx; ++q) {
                            for (int e = 0; e < n.nd; ++e) zv[e] = z[e];
                            for (int i = 0; i < n.m; ++i) zv[n.nd + i] = 0.0;
                            for (int j = 0; j < na; ++j) zv[n.nd + idxA[j]] = V[(size_t)q * QPN_VE_MAXA + j];
                            qpo_comp_indices(g.d1, g.d2, g.np, g.M.data(), g.N.data(), g.o.data(), g.l1.data(), g.u1.data(), g.A.data(),
                                             g.B.data(), g.l2.data(), g.u2.data(), zv.data(), w.data(), 1e-2, mv.data());
                            uint8_t* nib = vmask + (size_t)k * vstride + (size_t)q * vb;
                            for (int i = 0; i < n.m; ++i) nib[i >> 1] |= (uint8_t)((mv[n.nd + i] & 0xf) << ((i & 1) * 4));
                        }
                    }
                }
                if (snap) std::memcpy(Xf.data() + (size_t)b.slots[k] * nv, x, sizeof(double) * nv);
            }
            b.sol = sol; b.mask = mask; b.dz = dz;
            b.vcount = vcount; b.vmask = vmask; b.vstride = vstride; b.vbytes = vb;
        }
    }
    void run_qep(int, const LevelGaviInfo& L, QepBatch** bs, int nb, bool snap) override {
        begin_round();
        const GaviData& g = L.g;
        const int dz = g.d1 + g.d2, ndl = (int)L.dec.size();
        std::vector<double> w(g.np + 1), z0(dz + 1), z(dz + 1), xn(nv);
        const int nproj = net->check_for_cycling ? net->num_projections : 0;
        for (int q = 0; q < nb; ++q) {
            QepBatch& b = *bs[q];
            i32.emplace_back(b.n, 0); int32_t* status = i32.back().data();
            i32.emplace_back(b.n, 0); int32_t* pivots = i32.back().data();
            u8.emplace_back(b.n, 0); uint8_t* moved = u8.back().data();
            f64.emplace_back((size_t)b.n * (nproj > 0 ? nproj : 1), 0.0); double* pv = f64.back().data();
            for (int k = 0; k < b.n; ++k) {
                double* x = X.data() + (size_t)b.slots[k] * nv;
                for (int j = 0; j < g.np; ++j) w[j] = x[L.par[j]];
                for (int j = 0; j < dz; ++j) z0[j] = j < ndl ? x[L.dec[j]] : 0.0;
                qpo_gavi_solve(g.d1, g.d2, g.np, g.M.data(), g.N.data(), g.o.data(), g.l1.data(), g.u1.data(), g.A.data(), g.B.data(),
                               g.l2.data(), g.u2.data(), w.data(), z0.data(), 1, 0, z.data(), nullptr, &status[k], &pivots[k], nullptr);
                if (status[k] == 1) {
                    std::memcpy(xn.data(), x, sizeof(double) * nv);
                    for (int j = 0; j < ndl; ++j) xn[L.dec[j]] = z[j];
                    double dn = 0.0;
                    for (int j = 0; j < nv; ++j) { const double e = xn[j] - x[j]; dn = std::fma(e, e, dn); }
                    moved[k] = !(std::sqrt(dn) < 1e-4);
                    if (moved[k]) {
                        std::memcpy(x, xn.data(), sizeof(double) * nv);
                        for (int p = 0; p < nproj; ++p) {
                            double acc = 0.0;
                            for (int j = 0; j < nv; ++j) acc = std::fma(x[j], net->proj[(size_t)p * nv + j], acc);
                            pv[(size_t)k * nproj + p] = acc;
                        }
                    }
                }
                if (snap) std::memcpy(Xf.data() + (size_t)b.slots[k] * nv, x, sizeof(double) * nv);
            }
            b.status = status; b.pivots = pivots; b.moved = moved; b.pv = pv;
        }
    }
    void run_member(MemberBatch** bs, int nb) override;
    void finish() override { fresh_round = true; }
    void projections(double* pv_out) override {
        const int nproj = net->check_for_cycling ? net->num_projections : 0;
        for (int b = 0; b < B; ++b)
            for (int k = 0; k < nproj; ++k) {
                double acc = 0.0;
                for (int j = 0; j < nv; ++j) acc = std::fma(X[(size_t)b * nv + j], net->proj[(size_t)k * nv + j], acc);
                pv_out[(size_t)b * nproj + k] = acc;
            }
    }
    void download(double* x_out, const uint8_t* solved) override {
        for (int b = 0; b < B; ++b)
            std::memcpy(x_out + (size_t)b * nv, (solved[b] ? X.data() : Xf.data()) + (size_t)b * nv, sizeof(double) * nv);
    }
};

struct PieceCM { int m = 0; std::vector<double> A, l, u; };      // column-major copy for qpo_halfspace_in

struct OracleStore : Store {
    const NetData* net = nullptr;
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
    B = B_; nv = net->nv;
    X.assign(x_init, x_init + (size_t)B * nv);
    Xf = X;
}
void OracleWorker::run_member(MemberBatch** bs, int nb) {
    begin_round();
    for (int q = 0; q < nb; ++q) {
        MemberBatch& b = *bs[q];
        const size_t np = b.pieces->size();
        std::vector<const PieceCM*> pcs(np);
        { std::shared_lock<std::shared_mutex> lk(store->mu); for (size_t p = 0; p < np; ++p) pcs[p] = &store->pieces[(*b.pieces)[p]]; }
        u8.emplace_back((size_t)b.n * np, 0);
        uint8_t* in = u8.back().data();
        for (int k = 0; k < b.n; ++k) {
            const double* x = X.data() + (size_t)b.slots[k] * nv;
            for (size_t p = 0; p < np; ++p) {
                const PieceCM* c = pcs[p];
                in[(size_t)k * np + p] =
                    c->m == 0 ? 1 : (uint8_t)qpo_halfspace_in(c->m, nv, c->A.data(), c->l.data(), c->u.data(), nullptr, nullptr, x, 1e-6);
            }
        }
        b.in = in;
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
