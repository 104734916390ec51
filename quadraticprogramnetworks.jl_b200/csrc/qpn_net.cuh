// Kernels of the native network state machine (csrc/net/): the numeric requests of a batch of instances whose x
// stays resident on the GPU (X: nv x slots, column-major), and the partition of the batch into cohorts.
//   net_cycle_kernel  : the cycle check of solve_base! (algorithm.jl:14-30) against the (instance, level) history
//   net_verify_kernel : verify_solution (qp_processing.jl:57-149) at x, then comp_indices (avi_solutions.jl:587-612)
//                       of the node's own GAVI at (x, lam) -- what process_qp needs before it builds a solution graph --
//                       and the vertices of the multiplier polytope for expand (avi_solutions.jl:252-255)
//   net_qep_kernel    : solve_qep (avi.jl:382-444) for a level GAVI with plans, the 1e-4 disagreement test and the
//                       cycle-check projections of the new iterate (algorithm.jl:14-30,95-99); x updated in place
//   net_member_kernel : x in closure(piece) (intersection.jl:74,82 through sets.jl:820-825)
//   net_round_*       : the members of every cohort sorted by what they were told, cut into parts, one representative's
//                       answers per part copied out -- the only per-round data the host reads
// The arithmetic of each numeric kernel is the arithmetic of the single-purpose kernels (same device functions, same
// summation orders), so results are bit-equal to the C oracle's.
#pragma once
#include "qpn_level.cuh"
#include "vertex_enum_warp.cuh"
#include "net/cycle_check.h"

namespace qpn {

// Resident tables of the store (device memory): what a request's group refers to.
struct NodeTabEntry {
    NodeDesc node;
    GaviDesc g;             // the node's own GAVI (process_solution_graph, avi.jl:447-475)
    const int32_t* par;
    int dz, pad;
};
struct GaviTabEntry {
    GaviDesc g;
    GaviPlans plans;
    const int32_t *dec, *par;
    int nd_level, n;
};
struct PieceTabEntry {
    const double *A, *l, *u;    // rows row-major over nv
    int m, pad;
};

// ---- a round of the state machine --------------------------------------------------------------------------------
// The instance slots of a worker live in an ORDER array in which every live cohort is a contiguous segment.  A round's
// posts number their members 0 .. total-1 (`dst`: cohort after cohort); every kernel of the round adds what it told a
// member to that member's 64-bit signature keys[dst] (a sum of mixed (position, byte) terms over the member's answer
// bytes: commutative, so the threads of a launch may add in any order and the sum is deterministic).  net_round_* then
// sort each cohort's members by signature (cub::DeviceSegmentedSort, stable), cut the sorted segments into parts and
// copy ONE representative's answers per part; the sorted slots are the next round's order array.
enum { RK_VERIFY = 1, RK_MEMBER = 2, RK_QEP = 3 };
struct CohortDev {
    int kind, src_off, n, dst_off;
    int first, count;                  // verify: its groups in the VGroup table; member: its MGroup, pieces per member
    int rep_bytes, cyc_level;          // verify: the cycle checks it begins with: levels cyc_level .. cyc_level + ncyc - 1,
    int ncyc, log_off, cyc_pos, pad1;  //         (log_off: unused), the offset of their byte in the answer row.
};                                     //          A solve_qep cohort may carry the verify request
                                       //         that follows a successful solve (first / count / cyc_* as for verify; cyc_pos = 8)
struct VGroup {                        // one (cohort, node) of a verify launch: `count` pairs from pair `start` of the launch
    int node, start, count, snap;      // (pair0: index of its first pair among all verify pairs of the round)
    int src_off, dst_off, want, rep_off;   // rep_off: byte offset of this node's answers in the member's answer row
    unsigned mask_off, vm_off;         // byte offsets of the group's rows in the mask / vertex-mask buffers
    int dz, vbytes, pair0, gate;       // gate: answered only where the cohort's solve_qep of this round succeeded and moved
};
struct QGroup { int gavi, start, count, snap, src_off, dst_off; };
struct MGroup { int start, count, src_off, dst_off, np, list_off; unsigned out_off; int pad; };   // pairs piece-major: idx = p * count + k
struct DoneDev { int src_off, n, result, pad; };
struct PartDev { int pos, cohort, old, data_off; };

__host__ __device__ __forceinline__ unsigned long long qpn_mix64(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
// the signature term of answer byte `byte` at offset `pos` of the member's answer row (the layout of Part::rep)
__host__ __device__ __forceinline__ unsigned long long qpn_sig_term(unsigned pos, unsigned byte) {
    return qpn_mix64((((unsigned long long)pos + 1ull) << 8) | (unsigned long long)(byte & 0xffu));
}

__device__ __forceinline__ int find_group_start(const int* starts, int ngroups, int b) {
    int lo = 0, hi = ngroups - 1;          // last group whose start <= b
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (starts[mid] <= b) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// adds the warp's terms to *key (every lane of the warp must call it)
__device__ __forceinline__ void sig_add_warp(unsigned long long* key, unsigned long long term) {
    for (int o = 16; o > 0; o >>= 1) term += __shfl_xor_sync(0xffffffffu, term, o);
    if ((threadIdx.x & 31) == 0 && term != 0ull) atomicAdd(key, term);
}

// One thread per member of the round: its slot, an empty signature, its own index as the sort payload.
__global__ void net_round_fill_kernel(const CohortDev* __restrict__ cohorts, const int* __restrict__ cstarts, int ncohorts, int total,
                                      const int32_t* __restrict__ order, int32_t* __restrict__ slot_of, unsigned long long* __restrict__ keys,
                                      int32_t* __restrict__ vals) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= total) return;
    const CohortDev c = cohorts[find_group_start(cstarts, ncohorts, d)];
    slot_of[d] = order[c.src_off + (d - c.dst_off)];
    keys[d] = 0ull;
    vals[d] = d;
}

// The members of finished cohorts get the index of their outcome.
__global__ void net_done_kernel(const DoneDev* __restrict__ done, const int32_t* __restrict__ order, int32_t* __restrict__ result_of) {
    const DoneDev d = done[blockIdx.x];
    for (int k = threadIdx.x; k < d.n; k += blockDim.x) result_of[order[d.src_off + k]] = d.result;
}

// The cycle checks a verify request begins with (algorithm.jl:14-30), one thread per member: levels cyc_level ..
// cyc_level + ncyc - 1 in turn, the first hit ends the chain.  History of (slot, level): a list of CHUNKS of
// QPN_HIST_CHUNK entries (nproj projections each, contiguous, so the loads of a chunk are in flight together and a history
// of k entries costs k / 8 dependent hops instead of k), newest chunk first; `cnt` entries in all.  A miss appends; a
// new chunk comes from the bump counter *alloc (the host keeps the log large enough for one chunk per check of the
// launch).  hit_out: 0 = no hit, 1 + level of the first hit -- the cycle byte of the member's answer row.
#define QPN_HIST_CHUNK 8
__global__ void net_cycle_kernel(const CohortDev* __restrict__ cohorts, const int* __restrict__ cidx, const int* __restrict__ cycstarts,
                                 int ncyc_cohorts, int total, int nlevels, int nproj, const int32_t* __restrict__ slot_of,
                                 const double* __restrict__ PV, int32_t* __restrict__ head, int32_t* __restrict__ count,
                                 double* __restrict__ ent_pv, int32_t* __restrict__ ent_next, int* __restrict__ alloc,
                                 uint8_t* __restrict__ hit_out, unsigned long long* __restrict__ keys,
                                 const int32_t* __restrict__ qep_status, const uint8_t* __restrict__ qep_moved) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int q = find_group_start(cycstarts, ncyc_cohorts, t);
    const CohortDev c = cohorts[cidx[q]];
    const int k = t - cycstarts[q], d = c.dst_off + k;
    if (c.kind == RK_QEP && !(qep_status[d] == 1 && qep_moved[d])) { hit_out[d] = 0; return; }   // no new pass begins
    const int slot = slot_of[d];
    const double* pv = PV + (size_t)slot * nproj;
    int code = 0;
    for (int j = 0; j < c.ncyc && !code; ++j) {
        const size_t idx = (size_t)slot * nlevels + c.cyc_level + j;
        const int cnt = count[idx];
        int hd = head[idx], hit = 0;
        int in_chunk = cnt ? ((cnt - 1) % QPN_HIST_CHUNK) + 1 : 0;       // entries of the newest chunk
        // (the reference scans its cache from the oldest entry; "any earlier iterate" does not depend on the order)
        for (int ch = hd; ch >= 0 && !hit; ch = ent_next[ch]) {
            const double* base = ent_pv + (size_t)ch * QPN_HIST_CHUNK * nproj;
#pragma unroll
            for (int e = 0; e < QPN_HIST_CHUNK; ++e)
                if (e < in_chunk) hit |= qpn_cycle_hit(pv, base + (size_t)e * nproj, nproj) ? 1 : 0;
            in_chunk = QPN_HIST_CHUNK;
        }
        if (hit) { code = 1 + c.cyc_level + j; break; }
        const int pos = cnt % QPN_HIST_CHUNK;
        if (pos == 0) {
            const int nc = atomicAdd(alloc, 1);
            ent_next[nc] = hd;
            head[idx] = nc;
            hd = nc;
        }
        double* dst = ent_pv + ((size_t)hd * QPN_HIST_CHUNK + pos) * nproj;
        for (int i = 0; i < nproj; ++i) dst[i] = pv[i];
        count[idx] = cnt + 1;
    }
    hit_out[d] = (uint8_t)code;
    if (code) atomicAdd(keys + d, qpn_sig_term((unsigned)c.cyc_pos, (unsigned)code));     // (other kernels add to the same signature)
}

// grid = the verify requests of a round that explore vertices, or those that do not ((cohort, node) groups back to back:
// two launches, so that the requests of the first level do not carry the vertex scratch), block = roundup32(max over the
// groups of max(m, nd, 1)).  Dynamic smem: the largest group's verify_solution_kernel layout (Tab(m, m+1) + VerifySmem + x(nv) +
// qt(nd) + ax(m)) plus the vertex scratch (ve_scratch_bytes).
// MAXT = 64: the nodes of the examples (at most 64 rows): registers capped so that 24 one-warp CTAs fit an SM; MAXT = 1024:
// any node.
template <int MAXT>
__global__ void __launch_bounds__(MAXT, MAXT <= 64 ? 12 : 1)
net_verify_kernel(const NodeTabEntry* __restrict__ table, const VGroup* __restrict__ groups,
                                  const int* __restrict__ gstarts, int ngroups, const int32_t* __restrict__ order,
                                  const double* __restrict__ X, double* __restrict__ Xf_all, double tol,
                                  uint8_t* __restrict__ solution_out, int8_t* __restrict__ mask_base,
                                  uint8_t* __restrict__ vcount_out, uint8_t* __restrict__ vmask_base,
                                  unsigned long long* __restrict__ keys, const int32_t* __restrict__ qep_status,
                                  const uint8_t* __restrict__ qep_moved) {
    const int b = blockIdx.x, i = threadIdx.x;
    const VGroup grp = groups[find_group_start(gstarts, ngroups, b)];
    if (grp.gate) {                                       // rides with a solve_qep: only where it succeeded and moved
        const int d = grp.dst_off + (b - grp.start);
        if (!(qep_status[d] == 1 && qep_moved[d])) {
            if (i == 0) { solution_out[grp.pair0 + (b - grp.start)] = 0; if (grp.want > 0) vcount_out[grp.pair0 + (b - grp.start)] = 0; }
            return;
        }
    }
    const NodeTabEntry& ent = table[grp.node];
    const NodeDesc node = ent.node;
    const GaviDesc g = ent.g;
    const int32_t* par = ent.par;
    const int m = node.m, nd = node.nd, want_v = grp.want;
    double* Xf = grp.snap ? Xf_all : nullptr;
    const int kloc = b - grp.start;
    int8_t* my_mask = mask_base + grp.mask_off + (size_t)kloc * (nd + m);
    uint8_t* my_vmask = vmask_base + grp.vm_off + (size_t)kloc * want_v * ((m + 1) >> 1);
    const int slot = order[grp.src_off + kloc];
    unsigned long long* key = keys + grp.dst_off + kloc;
    const unsigned rp = (unsigned)grp.rep_off;
    unsigned long long sig = 0ull;
    const int tn = m > 0 ? m : 1;
    Tab tab;
    tab_carve(tab, tn, tn + 1, 0);
    VerifySmem vs;
    const int p = (int)tab_smem_bytes(tn, tn + 1);
    verify_carve(vs, nd, m, p);
    double* xs = reinterpret_cast<double*>(qpn_smem + p + verify_smem_bytes(nd, m));
    double* qt = xs + node.nv;
    double* ax = qt + nd;
    for (int j = i; j < node.nv; j += blockDim.x) {
        const double v = X[(size_t)slot * node.nv + j];
        xs[j] = v;
        if (Xf) Xf[(size_t)slot * node.nv + j] = v;       // the level-1 iterate as the reference would report it on failure
    }
    QPN_SYNC();
    node_products(node, xs, qt, ax);
    QPN_SYNC();
    int how = 0, piv = 0;
    const int sol = verify_solution_smem<false>(tab, vs, node, qt, ax, tol, &how, &piv);
    QPN_SYNC();
    const int dz = nd + m;
    if (sol) {
        // comp_indices at z = [x_dec; lam], w = x_par (comp_indices_kernel's sums, term for term)
        const double* lam = vs.lam_out();
        for (int r = i; r < dz; r += blockDim.x) {
            double acc = 0.0, acc2 = 0.0;
            int8_t mk;
            if (r < g.d1) {
                for (int j = 0; j < nd; ++j) acc = fma(g.M[(size_t)j * g.d1 + r], xs[node.dec[j]], acc);
                for (int j = 0; j < m; ++j) acc = fma(g.M[(size_t)(nd + j) * g.d1 + r], lam[j], acc);
                for (int j = 0; j < g.np; ++j) acc2 = fma(g.N[(size_t)j * g.d1 + r], xs[par[j]], acc2);
                const double rr = (acc + acc2) + g.o[r];
                mk = comp_mask(g.l1[r], g.u1[r], rr, xs[node.dec[r]], 1e-2);
            } else {
                const int k = r - g.d1;
                for (int j = 0; j < nd; ++j) acc = fma(g.A[(size_t)j * g.d2 + k], xs[node.dec[j]], acc);
                for (int j = 0; j < m; ++j) acc = fma(g.A[(size_t)(nd + j) * g.d2 + k], lam[j], acc);
                for (int j = 0; j < g.np; ++j) acc2 = fma(g.B[(size_t)j * g.d2 + k], xs[par[j]], acc2);
                const double s = acc + acc2;
                mk = comp_mask(g.l2[k], g.u2[k], lam[k], s, 1e-2);
            }
            my_mask[r] = mk;
            sig += qpn_sig_term(rp + 1u + (unsigned)r, (unsigned)(uint8_t)mk);
        }
    }
    const int pair = grp.pair0 + kloc;
    if (i == 0) { solution_out[pair] = (uint8_t)sol; sig += qpn_sig_term(rp, (unsigned)sol); }
    if (want_v > 0) {
        // expand's get_verts (avi_solutions.jl:252-255): the vertices of the node's multiplier polytope at x, enumerated by
        // the first warp (vertex_enum_warp.cuh), then comp_indices at every vertex by all threads -- only the masks of the m
        // multiplier rows can differ from the point's own, two rows per output byte
        const VeSmem ve = ve_carve(ax + m, nd);           // behind the kernel's vectors
        double* Vs = ve.V;
        int* hdr = ve.hdr;                                // [0] vertices, [1] active rows; their indices in ve.idxA
        if (i < 32) {
            int nvx = 0, a = 0;
            if (sol) nvx = multiplier_vertices_warp(ve, nd, m, node.A, node.dec, node.l, node.u, ax, qt, vs.lam_out(), want_v, &a);
            if (i == 0) {
                hdr[0] = nvx; hdr[1] = a;
                vcount_out[pair] = (uint8_t)nvx;
                if (sol) sig += qpn_sig_term(rp + 1u + (unsigned)dz, (unsigned)nvx);
            }
        }
        QPN_SYNC();
        const int nvx = hdr[0], a = hdr[1], vbytes = (m + 1) >> 1;
        double* lv = vs.zs();
        for (int q = 0; q < nvx; ++q) {
            for (int r = i; r < m; r += blockDim.x) lv[r] = 0.0;
            QPN_SYNC();
            for (int j = i; j < a; j += blockDim.x) lv[ve.idxA[j]] = Vs[q * QPN_VE_MAXA + j];
            QPN_SYNC();
            for (int t = i; t < vbytes; t += blockDim.x) {
                int packed = 0;
                for (int h = 0; h < 2; ++h) {
                    const int k = 2 * t + h;
                    if (k >= m) break;
                    double acc = 0.0, acc2 = 0.0;
                    for (int j = 0; j < nd; ++j) acc = fma(g.A[(size_t)j * g.d2 + k], xs[node.dec[j]], acc);
                    for (int j = 0; j < m; ++j) acc = fma(g.A[(size_t)(nd + j) * g.d2 + k], lv[j], acc);
                    for (int j = 0; j < g.np; ++j) acc2 = fma(g.B[(size_t)j * g.d2 + k], xs[par[j]], acc2);
                    packed |= (comp_mask(g.l2[k], g.u2[k], lv[k], acc + acc2, 1e-2) & 0xf) << (4 * h);
                }
                my_vmask[(size_t)q * vbytes + t] = (uint8_t)packed;
                sig += qpn_sig_term(rp + 2u + (unsigned)dz + (unsigned)(q * vbytes + t), (unsigned)packed);
            }
            QPN_SYNC();
        }
    }
    sig_add_warp(key, sig);
}

// grid = the solve_qep requests of a round whose level GAVIs fall into this thread bucket (groups back to back),
// block = the bucket's largest roundup32(lifted n).  Dynamic smem: the largest group's gavi_solve_kernel layout + x(nv) +
// xn(nv) + pv(nproj).
template <int MAXT>
__global__ void __launch_bounds__(MAXT, 896 / MAXT)
net_qep_kernel(const GaviTabEntry* __restrict__ table, const QGroup* __restrict__ groups, const int* __restrict__ gstarts,
               int ngroups, int nv, int nproj, const double* __restrict__ proj, const int32_t* __restrict__ order,
               double* __restrict__ X, double* __restrict__ Xf_all, int32_t* __restrict__ status_out,
               uint8_t* __restrict__ moved_out, double* __restrict__ PV, unsigned long long* __restrict__ keys) {
    const int b = blockIdx.x, i = threadIdx.x;
    const QGroup grp = groups[find_group_start(gstarts, ngroups, b)];
    const GaviTabEntry& ent = table[grp.gavi];
    const GaviDesc g = ent.g;
    const GaviPlans& plans = ent.plans;
    const int32_t* dec = ent.dec;
    const int32_t* par = ent.par;
    const int nd_level = ent.nd_level;
    double* Xf = grp.snap ? Xf_all : nullptr;
    const int kloc = b - grp.start;
    const int slot = order[grp.src_off + kloc], d = grp.dst_off + kloc;
    const int dz = g.d1 + g.d2, n = g.d1 + 2 * g.d2;
    const int max_pivots = 50 * n + 100;
    GaviSmem s;
    tab_carve_ex(s.t, n, (size_t)plans.t_doubles, plans.ldr_max, 0);
    const int p = gavi_carve_extra(s, g, (int)tab_smem_bytes_ex(n, (size_t)plans.t_doubles, plans.ldr_max));
    double* xs = reinterpret_cast<double*>(qpn_smem + p);
    double* xn = xs + nv;
    double* x = X + (size_t)slot * nv;
    for (int j = i; j < nv; j += blockDim.x) xs[j] = x[j];
    QPN_SYNC();
    for (int j = i; j < g.np; j += blockDim.x) s.w()[j] = xs[par[j]];
    for (int j = i; j < dz; j += blockDim.x) s.z0()[j] = j < nd_level ? xs[dec[j]] : 0.0;
    QPN_SYNC();
    int piv = 0;
    const int st = gavi_solve_smem(s, g, plans.has ? &plans.A : nullptr, plans.has ? &plans.B : nullptr, 1, max_pivots, &piv);
    QPN_SYNC();
    int moved = 0;
    if (st == ST_SUCCESS) {
        for (int j = i; j < nv; j += blockDim.x) xn[j] = xs[j];
        QPN_SYNC();
        for (int j = i; j < nd_level; j += blockDim.x) xn[dec[j]] = s.zs()[j];
        QPN_SYNC();
        double dn = 0.0;
        for (int j = 0; j < nv; ++j) { const double e = xn[j] - xs[j]; dn = fma(e, e, dn); }
        moved = !(sqrt(dn) < 1e-4);                       // algorithm.jl:96-97
        if (moved) {
            for (int j = i; j < nv; j += blockDim.x) x[j] = xn[j];
            for (int k = i; k < nproj; k += blockDim.x) {     // what the next cycle check of this instance compares
                double acc = 0.0;
                for (int j = 0; j < nv; ++j) acc = fma(xn[j], proj[(size_t)k * nv + j], acc);
                PV[(size_t)slot * nproj + k] = acc;
            }
        }
    }
    if (Xf) {
        const double* src = moved ? xn : xs;
        for (int j = i; j < nv; j += blockDim.x) Xf[(size_t)slot * nv + j] = src[j];
    }
    if (i == 0) {
        status_out[d] = st; moved_out[d] = (uint8_t)moved;
        // answer row: [status: int32 little endian] [moved]
        keys[d] += qpn_sig_term(0u, (unsigned)st & 0xffu) + qpn_sig_term(1u, ((unsigned)st >> 8) & 0xffu) + qpn_sig_term(4u, (unsigned)moved);
    }
}

// pv[b][k] = sum_j x_b[j] proj[k][j] (sequential fma, as the level kernel's cycle check): one thread per (slot, k).
__global__ void net_proj_kernel(int B, int nv, int nproj, const double* __restrict__ X, const double* __restrict__ proj,
                                double* __restrict__ pv) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)B * nproj) return;
    const int b = (int)(gid / nproj), k = (int)(gid - (long long)b * nproj);
    const double* x = X + (size_t)b * nv;
    double acc = 0.0;
    for (int j = 0; j < nv; ++j) acc = fma(x[j], proj[(size_t)k * nv + j], acc);
    pv[gid] = acc;
}

// x_out[b] = solved(result_of[b]) ? X[b] : Xf[b]  (what the reference returns as x_opt / x_fail)
__global__ void net_select_kernel(int B, int nv, const double* __restrict__ X, const double* __restrict__ Xf,
                                  const int32_t* __restrict__ result_of, const uint8_t* __restrict__ solved_of_result, int nresults,
                                  double* __restrict__ x_out) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)B * nv) return;
    const int r = result_of[gid / nv];
    x_out[gid] = (r >= 0 && r < nresults && solved_of_result[r]) ? X[gid] : Xf[gid];
}
// a[i] = i (or `fill` when fill >= -1... see callers): start of a batch
__global__ void net_iota_kernel(int n, int32_t* __restrict__ a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = i;
}
__global__ void net_fill_i32_kernel(long long n, int32_t* __restrict__ a, int fill) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = fill;
}

// Membership: one thread per (member, piece) pair of a cohort, consecutive threads = consecutive members of the same
// piece (the piece's rows are read at one address by the whole warp); each dot product sequential in the coordinate
// index.  A pair whose piece has no rows is inside.
__global__ void net_member_kernel(int npairs, const MGroup* __restrict__ groups, const int* __restrict__ gstarts, int ngroups,
                                  const int32_t* __restrict__ piece_ids, const PieceTabEntry* __restrict__ pieces, int nv,
                                  const int32_t* __restrict__ order, const double* __restrict__ X, double tol, uint8_t* __restrict__ in_out,
                                  unsigned long long* __restrict__ keys) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= npairs) return;
    const MGroup grp = groups[find_group_start(gstarts, ngroups, t)];
    const int loc = t - grp.start, p = loc / grp.count, k = loc - p * grp.count;
    const PieceTabEntry pc = pieces[piece_ids[grp.list_off + p]];
    const double* x = X + (size_t)order[grp.src_off + k] * nv;
    int ok = 1;
    for (int row = 0; row < pc.m; ++row) {
        const double* a = pc.A + (size_t)row * nv;
        double ax = 0.0;
        for (int j = 0; j < nv; ++j) ax = fma(a[j], x[j], ax);
        if (!((pc.l[row] - tol <= ax) && (ax - tol <= pc.u[row]))) ok = 0;
    }
    in_out[grp.out_off + (size_t)k * grp.np + p] = (uint8_t)ok;
    atomicAdd(keys + grp.dst_off + k, qpn_sig_term((unsigned)p, (unsigned)ok));
}

// After the segmented sort: position `pos` of the new order holds the member that had index vals[pos] in this round.
// A position whose signature differs from its predecessor's (or that opens a cohort) starts a part: it takes the next
// part record and room for one answer row.  hdr[0] = parts, hdr[1] = answer bytes.
__global__ void net_round_boundary_kernel(const CohortDev* __restrict__ cohorts, const int* __restrict__ cstarts, int ncohorts, int total,
                                          const unsigned long long* __restrict__ keys_sorted, const int32_t* __restrict__ vals_sorted,
                                          const int32_t* __restrict__ slot_of, int32_t* __restrict__ order_next, PartDev* __restrict__ parts,
                                          int* __restrict__ hdr) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= total) return;
    const int ci = find_group_start(cstarts, ncohorts, pos);
    const CohortDev c = cohorts[ci];
    const int old = vals_sorted[pos];
    order_next[pos] = slot_of[old];
    if (pos == c.dst_off || keys_sorted[pos] != keys_sorted[pos - 1]) {
        const int q = atomicAdd(hdr, 1);
        const int off = atomicAdd(hdr + 1, (c.rep_bytes + 7) & ~7);
        parts[q] = PartDev{pos, ci, old, off};
    }
}

// One CTA per part (grid-stride over the parts the boundary kernel counted): the representative's answers, in the layout
// of Part::rep (csrc/net/netsolver.hpp).
__global__ void net_round_gather_kernel(const CohortDev* __restrict__ cohorts, const VGroup* __restrict__ vgroups, const MGroup* __restrict__ mgroups,
                                        const PartDev* __restrict__ parts, const int* __restrict__ hdr, const uint8_t* __restrict__ hit,
                                        const uint8_t* __restrict__ sol, const int8_t* __restrict__ mask_base, const uint8_t* __restrict__ vcount,
                                        const uint8_t* __restrict__ vmask_base, const uint8_t* __restrict__ in_bits,
                                        const int32_t* __restrict__ status, const uint8_t* __restrict__ moved, uint8_t* __restrict__ rep) {
    const int nparts = hdr[0];
    for (int q = blockIdx.x; q < nparts; q += gridDim.x) {
        const PartDev pt = parts[q];
        const CohortDev c = cohorts[pt.cohort];
        uint8_t* out = rep + pt.data_off;
        const int k = pt.old - c.dst_off;
        if (c.kind == RK_QEP) {
            if (threadIdx.x == 0) {
                const int st = status[pt.old];
                out[0] = (uint8_t)(st & 0xff); out[1] = (uint8_t)((st >> 8) & 0xff); out[2] = (uint8_t)((st >> 16) & 0xff); out[3] = (uint8_t)((st >> 24) & 0xff);
                out[4] = moved[pt.old]; out[5] = out[6] = out[7] = 0;
            }
        }
        if (c.kind == RK_MEMBER) {
            const MGroup g = mgroups[c.first];
            for (int p = threadIdx.x; p < g.np; p += blockDim.x) out[p] = in_bits[g.out_off + (size_t)k * g.np + p];
        } else if (c.kind == RK_VERIFY || c.count > 0) {     // a verify request, alone or behind a solve_qep
            if (threadIdx.x == 0) out[c.cyc_pos] = c.ncyc > 0 ? hit[pt.old] : 0;
            for (int r = 0; r < c.count; ++r) {
                const VGroup g = vgroups[c.first + r];
                uint8_t* o = out + g.rep_off;
                const int pair = g.pair0 + k;
                const int s = sol[pair];
                if (threadIdx.x == 0) { o[0] = (uint8_t)s; o[1 + g.dz] = (s && g.want > 0) ? vcount[pair] : 0; }
                if (s) {
                    const int8_t* mk = mask_base + g.mask_off + (size_t)k * g.dz;
                    for (int j = threadIdx.x; j < g.dz; j += blockDim.x) o[1 + j] = (uint8_t)mk[j];
                    if (g.want > 0) {
                        const int nb = (int)vcount[pair] * g.vbytes;
                        const uint8_t* vm = vmask_base + g.vm_off + (size_t)k * g.want * g.vbytes;
                        for (int j = threadIdx.x; j < nb; j += blockDim.x) o[2 + g.dz + j] = vm[j];
                    }
                }
            }
        }
    }
}

}  // namespace qpn
