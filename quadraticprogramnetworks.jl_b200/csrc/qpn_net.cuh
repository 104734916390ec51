// Kernels of the native network state machine (csrc/net/): the numeric requests of a batch of instances whose x
// stays resident on the GPU (X: nv x slots, column-major).  Every launch works on a LIST of instance slots -- the
// instances that asked for the same thing this round -- against a resident node / level GAVI / piece.
//   net_verify_kernel : verify_solution (qp_processing.jl:57-149) at x, then comp_indices (avi_solutions.jl:587-612)
//                       of the node's own GAVI at (x, lam) -- what process_qp needs before it builds a solution graph
//   net_qep_kernel    : solve_qep (avi.jl:382-444) for a level GAVI with plans, the 1e-4 disagreement test and the
//                       cycle-check projections of the new iterate (algorithm.jl:14-30,95-99); x updated in place
//   net_member_kernel : x in closure(piece) (intersection.jl:74,82 through sets.jl:820-825), one warp per (instance, piece)
// The arithmetic of each is the arithmetic of the single-purpose kernels (same device functions, same summation
// orders), so results are bit-equal to the C oracle's.
#pragma once
#include "qpn_level.cuh"
#include "net/vertex_enum.h"

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
// One group of a launch: `count` consecutive requests (from `start`) against the same resident object.
struct VGroup { int node, start, count, snap; unsigned mask_off, vm_off; int want, pad; };
struct QGroup { int gavi, start, count, snap; };

__device__ __forceinline__ int find_group_start(const int* starts, int ngroups, int b) {
    int lo = 0, hi = ngroups - 1;          // last group whose start <= b
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (starts[mid] <= b) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// grid = all verify requests of a round (every node's group back to back), block = roundup32(max over the groups of
// max(m, nd, 1)).  Dynamic smem: the largest group's verify_solution_kernel layout (Tab(m, m+1) + VerifySmem + x(nv) +
// qt(nd) + ax(m)) plus the vertex scratch.
__global__ void net_verify_kernel(const NodeTabEntry* __restrict__ table, const VGroup* __restrict__ groups,
                                  const int* __restrict__ gstarts, int ngroups, const int32_t* __restrict__ inst,
                                  const double* __restrict__ X, double* __restrict__ Xf_all, double tol,
                                  uint8_t* __restrict__ solution_out, int8_t* __restrict__ mask_base,
                                  uint8_t* __restrict__ vcount_out, uint8_t* __restrict__ vmask_base) {
    const int b = blockIdx.x, i = threadIdx.x;
    const VGroup grp = groups[find_group_start(gstarts, ngroups, b)];
    const NodeTabEntry& ent = table[grp.node];
    const NodeDesc node = ent.node;
    const GaviDesc g = ent.g;
    const int32_t* par = ent.par;
    const int m = node.m, nd = node.nd, want_v = grp.want;
    double* Xf = grp.snap ? Xf_all : nullptr;
    const int kloc = b - grp.start;
    int8_t* my_mask = mask_base + grp.mask_off + (size_t)kloc * (nd + m);
    uint8_t* my_vmask = vmask_base + grp.vm_off + (size_t)kloc * want_v * ((m + 1) >> 1);
    const int slot = inst[b];
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
        }
    }
    if (i == 0) solution_out[b] = (uint8_t)sol;
    if (want_v > 0) {
        // expand's get_verts (avi_solutions.jl:252-255): the vertices of the node's multiplier polytope at x, enumerated by
        // one thread (net/vertex_enum.h), then comp_indices at every vertex by all of them -- only the masks of the m
        // multiplier rows can differ from the point's own, two rows per output byte
        double* Vs = ax + m;                              // QPN_VE_MAXV x QPN_VE_MAXA doubles behind the kernel's vectors
        int* hdr = reinterpret_cast<int*>(Vs + QPN_VE_MAXV * QPN_VE_MAXA);     // [0] vertices, [1] active rows, [2..] their indices
        if (i == 0) {
            int nvx = 0, a = 0;
            if (sol) nvx = qpn_multiplier_vertices(nd, m, node.nv, node.A, node.dec, node.l, node.u, ax, qt, vs.lam_out(), want_v, hdr + 2, &a, Vs);
            hdr[0] = nvx; hdr[1] = a;
            vcount_out[b] = (uint8_t)nvx;
        }
        QPN_SYNC();
        const int nvx = hdr[0], a = hdr[1], vbytes = (m + 1) >> 1;
        double* lv = vs.zs();
        for (int q = 0; q < nvx; ++q) {
            for (int r = i; r < m; r += blockDim.x) lv[r] = 0.0;
            QPN_SYNC();
            for (int j = i; j < a; j += blockDim.x) lv[hdr[2 + j]] = Vs[q * QPN_VE_MAXA + j];
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
            }
            QPN_SYNC();
        }
    }
}

// grid = the solve_qep requests of a round whose level GAVIs fall into this thread bucket (groups back to back),
// block = the bucket's largest roundup32(lifted n).  Dynamic smem: the largest group's gavi_solve_kernel layout + x(nv) +
// xn(nv) + pv(nproj).
template <int MAXT>
__global__ void __launch_bounds__(MAXT, 896 / MAXT)
net_qep_kernel(const GaviTabEntry* __restrict__ table, const QGroup* __restrict__ groups, const int* __restrict__ gstarts,
               int ngroups, int nv, int nproj, const double* __restrict__ proj, const int32_t* __restrict__ inst,
               double* __restrict__ X, double* __restrict__ Xf_all, int32_t* __restrict__ status_out,
               int32_t* __restrict__ pivots_out, uint8_t* __restrict__ moved_out, double* __restrict__ pv_out) {
    const int b = blockIdx.x, i = threadIdx.x;
    const QGroup grp = groups[find_group_start(gstarts, ngroups, b)];
    const GaviTabEntry& ent = table[grp.gavi];
    const GaviDesc g = ent.g;
    const GaviPlans& plans = ent.plans;
    const int32_t* dec = ent.dec;
    const int32_t* par = ent.par;
    const int nd_level = ent.nd_level;
    double* Xf = grp.snap ? Xf_all : nullptr;
    const int slot = inst[b];
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
            for (int k = i; k < nproj; k += blockDim.x) {
                double acc = 0.0;
                for (int j = 0; j < nv; ++j) acc = fma(xn[j], proj[(size_t)k * nv + j], acc);
                pv_out[(size_t)b * nproj + k] = acc;
            }
        }
    }
    if (Xf) {
        const double* src = moved ? xn : xs;
        for (int j = i; j < nv; j += blockDim.x) Xf[(size_t)slot * nv + j] = src[j];
    }
    if (i == 0) { status_out[b] = st; pivots_out[b] = piv; moved_out[b] = (uint8_t)moved; }
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

// x_out[b] = solved[b] ? X[b] : Xf[b]  (what the reference returns as x_opt / x_fail)
__global__ void net_select_kernel(int B, int nv, const double* __restrict__ X, const double* __restrict__ Xf,
                                  const uint8_t* __restrict__ solved, double* __restrict__ x_out) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)B * nv) return;
    x_out[gid] = solved[gid / nv] ? X[gid] : Xf[gid];
}

// One (instance, piece) pair of a membership round.
struct MemberPair { int32_t inst, piece; };

// One warp per pair; lanes take rows, each dot product sequential in the coordinate index.
__global__ void net_member_kernel(int npairs, const MemberPair* __restrict__ pairs, const PieceTabEntry* __restrict__ pieces, int nv,
                                  const double* __restrict__ X, double tol, uint8_t* __restrict__ in_out) {
    const int warp = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (warp >= npairs) return;
    const MemberPair pr = pairs[warp];
    const PieceTabEntry pc = pieces[pr.piece];
    const double* x = X + (size_t)pr.inst * nv;
    int ok = 1;
    for (int row = lane; row < pc.m; row += 32) {
        const double* a = pc.A + (size_t)row * nv;
        double ax = 0.0;
        for (int j = 0; j < nv; ++j) ax = fma(a[j], x[j], ax);
        if (!((pc.l[row] - tol <= ax) && (ax - tol <= pc.u[row]))) ok = 0;
    }
    ok = __all_sync(0xffffffffu, ok);
    if (lane == 0) in_out[warp] = (uint8_t)ok;
}

}  // namespace qpn
