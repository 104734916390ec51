// Native batched host state machine of solve(qpn, inits::Matrix) for networks with children.
//
// What it replaces, in the reference: the recursion of solve_base! (/root/reference/src/algorithm.jl:1-127),
// process_qp and combine (src/qp_processing.jl:151-291), the IntersectionRoot walk (src/intersection.jl:55-151) and
// collect(LocalGAVISolutions) with all_Ks / local_piece / expand / project_and_permute
// (src/avi_solutions.jl:200-215,241-321,400-496,79-90).  Every numeric step -- verify_solution, comp_indices,
// solve_qep, membership, every LP of the set algebra -- is a request to a numeric backend (the CUDA engine in
// libqpn_cuda; the C oracle in the test / baseline build under oracle/).
//
// Shape.  The control state of solve_base! -- which level, which iteration, which solution graphs
// came up from the children, which child pieces are assigned -- takes few distinct values across a batch (8,192
// perturbed robust_avoid instances walk 105 distinct control paths).  The machine therefore advances COHORTS:
// sets of instances that share their whole control state.  A cohort runs the reference's recursion as an explicit,
// copyable state machine until it needs numbers and then posts ONE request per round over its members (verify these
// nodes -- with the cycle checks that precede them --, these membership tests, this solve_qep -- with the verify request
// that follows a successful solve).  The BACKEND answers every member, partitions the cohort by the answers and hands
// back one representative's answers per part (on the GPU the partition is a segmented sort by a 64-bit signature of
// the answers and the host never sees per-instance data; x, the cycle-check histories and the instance order stay
// resident there); every part goes on with a copy of the state.  Everything geometric is a pure function of exact
// problem data (node, child pieces, complementarity recipe K), never of the instance, and so are the machine's own
// transitions (what a level asks to have verified; what follows the answers): all of it is memoised in a cache
// shared by all cohorts, worker threads and batches of the net.
#pragma once
#include <atomic>
#include <cstdint>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "poly.hpp"

namespace qpnnet {

// ---- the network as plain arrays (what setup(:name) produces; programs.jl:79-116) ----------------------------
struct NetData {
    int nv = 0, nplayers = 0, nlevels = 0;
    std::vector<std::vector<double>> Q;          // per player: nv x nv, row-major
    std::vector<std::vector<double>> q;          // per player: nv
    std::vector<std::vector<int>> base;          // per player: poly ids of its own constraints (constraint_indices order)
    std::vector<std::vector<int>> children;      // network_edges (sorted)
    std::vector<std::vector<int>> dec;           // decision_inds (sorted; programs.jl:340-346)
    std::vector<std::vector<int>> levels;        // players per level (sorted), levels[0] = level 1
    std::vector<int> level_of;                   // 0-based level per player
    // options (QPNetOptions, programs.jl:61-77)
    int max_iters = 150, num_projections = 4, exploration_vertices = 0, gen_solution_map = 0, check_for_cycling = 1;
    std::vector<char> remove_subsets_at;         // per level (levels_to_remove_subsets)
    std::vector<double> proj;                    // num_projections x nv, row-major
};

// ---- resident objects ---------------------------------------------------------------------------------------
// One node's view for verify_solution (qp_processing.jl:57-66) plus its single-node GAVI (avi.jl:447-475).
struct NodeInfo {
    int pid = 0, nd = 0, nv = 0, m = 0;
    std::vector<int> polys;                      // base constraint polys followed by the chosen child pieces
    std::vector<double> Qd, qd, A, l, u;         // column-major: Qd nd x nv, A m x nv
    std::vector<int32_t> dec, par;
    GaviData g;                                  // z = [x_dec; lam], w = x_par
};

// The joint KKT system of a level's players for one assignment of child pieces (avi.jl:305-377,382-404).
struct LevelGaviInfo {
    int level = 0;
    GaviData g;
    std::vector<int32_t> dec, par;
};

// ---- requests a cohort posts -------------------------------------------------------------------------------------
// A cohort is a set of instances that share their whole control state (level stack, solution graphs, iteration
// counters): whatever one of them asks, all of them ask.  The backend keeps the instance slots of a worker in an ORDER
// array in which every live cohort is a contiguous segment; a request names a resident object and the cohort's
// segment.  The backend answers every member, partitions the segment by the members' answers (on the GPU: a 64-bit
// signature per member, a segmented sort, a boundary scan -- the host never sees per-instance data) and hands back,
// per part, its new segment and the answers of ONE representative member.
struct Seg { int off = 0, n = 0; };

enum PostKind { POST_VERIFY = 1, POST_MEMBER = 2, POST_QEP = 3 };

struct Post {
    int kind = POST_VERIFY;
    Seg seg;
    // POST_VERIFY: verify_solution at each instance's x, then comp_indices of the node GAVI at (x, lam), for each of the
    // cohort's nodes; want_vertices > 0: expand's get_verts too (avi_solutions.jl:252-255): up to that many new vertices of
    // the multiplier polytope and the comp_indices masks of the m multiplier rows at each, two rows per byte.
    // Before that, the cycle checks the loop passes leading here begin with (algorithm.jl:14-30): levels
    // cyc_level .. cyc_level + ncyc - 1 in turn, each against the iterate history of (instance, level), a miss appended,
    // the first hit ends the chain (a new pass of level L opens fresh passes of every level below it before anything is
    // verified, and x does not change in between, so the checks ride with the verify request instead of costing a
    // round each)
    int cyc_level = 0, ncyc = 0;
    const int* nodes = nullptr;
    int nnodes = 0, want_vertices = 0, snap = 0;
    // POST_MEMBER: x in closure(piece) for the pieces of each list (intersection.jl:74,82)
    const std::vector<int>* const* piece_lists = nullptr;
    int nlists = 0;
    // POST_QEP: solve_qep for a level GAVI; on success (and a move of >= 1e-4) x[dec] is replaced.  What follows a
    // successful solve is known in advance -- the next loop pass of the level opens passes of every level below it and ends
    // up verifying the bottom level's nodes at the new x -- so that verify request (nodes / nnodes / want_vertices, its
    // cycle checks cyc_level / ncyc, vsnap in place of snap) rides along when nnodes > 0 and is answered for the members
    // whose solve succeeded and moved
    int gavi = 0, vsnap = 0;
};

// Answers of a part's representative, as bytes:
//   POST_VERIFY: [cycle: 0 = no hit, 1 + level of the first hit]; per node r: [sol] [mask: dz_r] [vcount]
//                [vmask: want * ceil(m_r / 2)]   (mask / vertices valid when sol)
//   POST_MEMBER: per list, per piece: [in]
//   POST_QEP   : [status: int32] [moved] [3 pad], then (nnodes > 0) the POST_VERIFY answers of the request that rides along
struct Part {
    int cohort = 0;                              // index of the post within the round
    Seg seg;                                     // the part's segment in the NEW order
    const uint8_t* rep = nullptr;                // valid until the worker's next round
};
static inline int verify_rep_bytes(int dz, int m, int want) { return 2 + dz + want * ((m + 1) / 2); }

// ---- numeric backend ---------------------------------------------------------------------------------------------
struct Worker : LPBackend {
    // instance slots 0..B-1 of this worker: x = x_fail = init (nv x B column-major); order = identity (one segment
    // {0, B}); cycle-check histories empty; the projections of every x taken
    virtual void set_batch(int B, const double* x_init) = 0;
    // enqueue a cohort's request for this round; returns its index within the round
    virtual int post(const Post& p) = 0;
    // the members of a finished cohort get `result` (an index into the caller's table of outcomes)
    virtual void mark_done(Seg seg, int result) = 0;
    // run the round: parts ordered by cohort, then by position in the new order
    virtual void finish_round(std::vector<Part>& parts) = 0;
    // result_of_slot (B): what mark_done recorded; x_out (nv x B) = x where solved_of_result[result], the reference's
    // x_fail otherwise (algorithm.jl:116,125)
    virtual void download(double* x_out, int32_t* result_of_slot, const uint8_t* solved_of_result, int nresults) = 0;
    virtual int64_t launches() const { return 0; }
};
struct Store {              // shared by the workers of one net: resident copies of nodes / GAVIs / pieces
    virtual ~Store() {}
    virtual Worker* make_worker() = 0;
    // called once per object, under the cache's creation lock, with the worker that met the object first
    virtual void new_node(int id, const NodeInfo& info, Worker* w) = 0;
    virtual void new_gavi(int id, const LevelGaviInfo& info, Worker* w) = 0;
    virtual void new_piece(int id, const Poly& P, Worker* w) = 0;
};

// Append-only array whose elements never move and can be READ without a lock while one writer appends (the writers
// are serialised by the cache's mutex; an id reaches a reader only through one of the cache's maps, under that mutex,
// after the element was constructed).  The state machine reads polys / lists / nodes several times per cohort per round.
template <class T> class StableVec {
    static constexpr size_t CH = 1024, MAXCH = 1 << 15;
    std::unique_ptr<std::atomic<T*>[]> chunks_;
    std::atomic<size_t> size_{0};

  public:
    StableVec() : chunks_(new std::atomic<T*>[MAXCH]()) {}
    StableVec(const StableVec&) = delete;
    StableVec& operator=(const StableVec&) = delete;
    ~StableVec() {
        const size_t n = size_.load();
        for (size_t i = 0; i < n; ++i) chunks_[i / CH].load()[i % CH].~T();
        for (size_t c = 0; c * CH < n; ++c) ::operator delete((void*)chunks_[c].load());
    }
    size_t size() const { return size_.load(std::memory_order_acquire); }
    const T& operator[](size_t i) const { return chunks_[i / CH].load(std::memory_order_acquire)[i % CH]; }
    void push_back(T v) {
        const size_t i = size_.load(std::memory_order_relaxed);
        if (i % CH == 0) chunks_[i / CH].store((T*)::operator new(sizeof(T) * CH), std::memory_order_release);
        new (&chunks_[i / CH].load(std::memory_order_relaxed)[i % CH]) T(std::move(v));
        size_.store(i + 1, std::memory_order_release);
    }
};

// ---- the cache of everything instance-independent -------------------------------------------------------------
struct VecHash {
    size_t operator()(const std::vector<int>& v) const {
        uint64_t h = 0x9E3779B97F4A7C15ull ^ v.size();
        for (int x : v) { h ^= (uint64_t)(uint32_t)x + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); }
        return (size_t)h;
    }
};

struct Stats {
    std::atomic<long> lps{0}, rounds{0}, requests{0}, calls{0}, pieces{0}, nodes{0}, gavis{0}, collect_miss{0}, combine_miss{0}, cohorts{0};
    std::atomic<long> host_ns{0}, backend_ns{0}, apply_ns{0};
    std::atomic<long> lps_empty{0}, lps_subset{0}, lps_project{0}, lp_calls{0};   // where the LPs of the set algebra come from   // summed over worker threads: instance logic / numeric backend (incl. waits)
};

// ---- memoised transitions of the cohort machine ------------------------------------------------------------------
// What a level asks to have verified is a pure function of (level, solution graphs below it); what follows the answers is
// a pure function of (that plan, the answers).  Both are memoised like the geometry, so a cohort's step in a round is two
// hash lookups instead of a walk through the (already memoised) set algebra.
struct VerifyPlan {                              // process_qp, first phase (qp_processing.jl:151-188)
    int level = 0, error = 0, want = 0, rep_bytes = 0;
    struct PV { int pid; std::vector<std::vector<int>> combos; int first_req; };
    std::vector<PV> pvs;                         // players and their child-piece combinations
    std::vector<int> req_nodes;                  // one node per (player, combination)
    std::vector<int> S;                          // the solution graphs it was built for (per player: list id, -1 = none)
};
struct VerifyOutcome {                           // what the answers to a plan lead to
    enum { FAIL = 0, QEP = 1, DONE = 2, MEMBER = 3 };
    int kind = FAIL, error = 0, plan = 0;
    int gavi = -1;                               // QEP: the level GAVI with the offending child pieces (algorithm.jl:68-101)
    int S_new = -1;                              // DONE: list id of the solution graphs with this level's added (algorithm.jl:84,104-116)
    struct Comb { int pid; std::vector<int> union_lists, red, flat; };
    std::vector<Comb> combs;                     // MEMBER: players whose leaves wait for membership bits (combine, qp_processing.jl:243-291)
    std::vector<const std::vector<int>*> comb_lists;
    std::vector<int> S_out;                      // MEMBER: the graphs that are already final
};

class GeoCache {
  public:
    GeoCache(const NetData& net, Store* store) : net_(net), store_(store) {}
    const NetData& net() const { return net_; }
    Stats stats;

    // polyhedra interned by exact content
    int intern_poly(Poly&& P, Worker* w);
    const Poly& poly(int id) const { return polys_[id]; }
    int set_id(int id) const { return set_ids_[id]; }
    int npolys() const { return (int)polys_.size(); }
    // lists of polyhedra (solution graphs) interned by their ids
    int intern_list(const std::vector<int>& ids);
    const std::vector<int>& list(int id) const { return lists_[id]; }

    int node(int pid, const std::vector<int>& pieces, Worker* w);                 // (player, child pieces) -> node id
    const NodeInfo& node_info(int id) const { return nodes_[id]; }
    int level_gavi(int level, const std::vector<int>& assignment, Worker* w);     // (level, piece per child) -> gavi id
    const LevelGaviInfo& gavi_info(int id) const { return gavis_[id]; }

    // geometry, memoised (the LPs run on worker w)
    bool empty(int poly, double tol, Worker* w);                                  // exemplar(P)[0]
    bool subset(int p1, int p2, Worker* w);
    int remove_subsets(int list, Worker* w);
    int intersect2(int a, int b, Worker* w);
    int intersect_all(const std::vector<int>& ids, Worker* w);
    const std::vector<int>& complement_of(int poly, Worker* w);
    // (node, recipe K) -> projected piece id, or -1 when the local piece is empty (avi_solutions.jl:241-261)
    int expand(int node, const std::string& K, Worker* w);
    // the lifted local piece of (node, K) over (z, w) (avi_solutions.jl:400-496)
    int local_piece_id(int node, const std::string& K, Worker* w);
    // collect(LocalGAVISolutions) without vertex exploration: (node, mask) -> list of piece ids
    // `extra`: the masks at the explored vertices, concatenated (each nd + m entries)
    int collect(int node, const std::vector<int8_t>& mask, const std::vector<int8_t>& extra, Worker* w, bool* bad_mask);
    // all non-empty leaves of the intersection tree for one membership pattern (intersection.jl:55-151)
    int leaves(const std::vector<int>& union_lists, const std::vector<int>& red_lengths, const std::vector<uint8_t>& in_bits,
               Worker* w);

    // memoised transitions
    int verify_plan(int level, const std::vector<int>& S, Worker* w);
    const VerifyPlan& plan(int id) const { return plans_[id]; }
    int verify_outcome(int plan, const uint8_t* answers, Worker* w);          // answers: the rows behind the cycle byte
    const VerifyOutcome& outcome(int id) const { return outcomes_[id]; }
    int member_outcome(int outcome, const uint8_t* bits, Worker* w);          // -> list id of the new solution graphs

  private:
    template <class Map, class Key, class F> auto memo(Map& map, const Key& key, F&& compute) -> typename Map::mapped_type;
    int finish_graphs(const VerifyPlan& P, const std::vector<int>& S_out, Worker* w);
    StableVec<VerifyPlan> plans_;
    std::unordered_map<std::vector<int>, int, VecHash> plan_ids_;
    StableVec<VerifyOutcome> outcomes_;
    std::unordered_map<std::string, int> outcome_ids_, member_ids_;
    const NetData& net_;
    Store* store_;
    mutable std::shared_mutex mu_;
    std::mutex create_mu_;                       // serialises the creation of resident objects
    StableVec<Poly> polys_;
    StableVec<int> set_ids_;
    std::unordered_map<std::string, int> poly_by_exact_, set_by_key_;
    StableVec<std::vector<int>> lists_;
    std::unordered_map<std::vector<int>, int, VecHash> list_ids_;
    StableVec<NodeInfo> nodes_;
    std::unordered_map<std::vector<int>, int, VecHash> node_ids_;
    StableVec<LevelGaviInfo> gavis_;
    std::unordered_map<std::vector<int>, int, VecHash> gavi_ids_;
    std::unordered_map<uint64_t, char> empty_, subset_;
    std::unordered_map<int, int> remove_subsets_;
    std::unordered_map<uint64_t, int> intersect2_;
    std::unordered_map<int, std::vector<int>> complement_;
    std::unordered_map<std::string, int> expand_, local_piece_, collect_;
    std::unordered_map<std::string, int> leaves_;
};

// ---- the solver --------------------------------------------------------------------------------------------------
struct SolveOut {
    uint8_t solved = 0;
    std::vector<int> level_iters;                // loop passes of solve_base! per level, summed over its calls
    std::vector<int> sol;                        // per player: list id of its solution graph, -1 = none
    int error = 0;                               // 0 none; see ERR_* in netsolver.cpp
    int pivots = 0;
};

class NetSolver {
  public:
    // polys: the constraint polys net.base refers to (by position); they are interned first
    NetSolver(NetData net, std::vector<Poly> polys, std::unique_ptr<Store> store);
    ~NetSolver();
    // inits: nv x B column-major.  x_out: nv x B (x_opt, or x_fail when not solved).  The outcome of instance b is
    // results[result_of[b]]: instances that ended in the same cohort share one entry (a batch has a few thousand).
    void solve_batched(int B, const double* inits, double* x_out, std::vector<int>& result_of, std::vector<SolveOut>& results, int threads);
    GeoCache& cache() { return *cache_; }
    const NetData& net() const { return net_; }
    std::string last_error;

  private:
    void run_shard(int tid, int lo, int hi, const double* inits, double* x_out, int* result_of, std::vector<SolveOut>& results);
    NetData net_;
    std::unique_ptr<Store> store_;
    std::unique_ptr<GeoCache> cache_;
    std::vector<std::unique_ptr<Worker>> workers_;
};

}  // namespace qpnnet
