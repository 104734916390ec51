// verify_solution and the fused per-level equilibrium loop for sizes beyond the shared-memory
// workspaces (qpn_level.cuh): the least-squares matrices and the tableau live in the CTA's
// global-memory slot (avi_pivot_big.cuh), every vector in shared memory.  Same arithmetic, same
// order of operations as the shared-memory kernels and the CPU oracle.
//
// Replaces /root/reference/src/qp_processing.jl:57-149 (verify_solution) and the level loop of
// /root/reference/src/algorithm.jl:13-118 at the sizes of BASELINE.json configs 4-5 (nd up to 256,
// m up to 512 rows per node; lifted level AVI up to n = 1,536).
#pragma once
#include "qpn_big.cuh"

namespace qpn {

// Vectors of one verify_solution call (shared memory) + the two nd x k matrices (global, stored
// with the column index fastest: element (row i, column j) at [i * ldk + j], so that the
// one-thread-per-column passes of the Householder QR are coalesced).
struct VerifyBig {
    int base, nd, m;
    double *Ab, *Ab0;     // global, nd x m each
    __device__ __forceinline__ double* dbl(int off) const { return reinterpret_cast<double*>(qpn_smem + base) + off; }
    __device__ __forceinline__ double* b() const { return dbl(0); }                    // nd
    __device__ __forceinline__ double* lam() const { return dbl(nd); }                 // m
    __device__ __forceinline__ double* v() const { return dbl(nd + m); }               // nd + m
    __device__ __forceinline__ double* lam_out() const { return dbl(2 * nd + 2 * m); } // m
    __device__ __forceinline__ double* qs() const { return dbl(2 * nd + 3 * m); }      // m
    __device__ __forceinline__ double* zs() const { return dbl(2 * nd + 4 * m); }      // m
    __device__ __forceinline__ double* zb() const { return dbl(2 * nd + 5 * m); }      // m
    __device__ __forceinline__ int* idx() const { return reinterpret_cast<int*>(dbl(2 * nd + 6 * m)); }   // m
    __device__ __forceinline__ int* perm() const { return idx() + m; }                 // m
    __device__ __forceinline__ int8_t* kind() const {                                  // m
        return reinterpret_cast<int8_t*>(reinterpret_cast<unsigned char*>(idx()) + ((2 * m * 4 + 15) / 16) * 16);
    }
};
__host__ __device__ __forceinline__ size_t verify_big_bytes(int nd, int m) {
    return 8 * (2 * (size_t)nd + 6 * (size_t)m) + ((2 * (size_t)m * 4 + 15) / 16) * 16 + (((size_t)m + 15) / 16) * 16;
}

// Householder QR least squares with column pivoting; the same sequence of operations as
// lstsq_basic_block (qpn_level.cuh) / oracle lstsq_basic.  A(i, j) = Ab[i * ldk + j].
__device__ __noinline__ void lstsq_basic_big(const Tab& red, int nd, int k, int ldk, double* Ab, double* b, double* lam, int* perm, double* v) {
    const int j = threadIdx.x;
    const int steps = nd < k ? nd : k;
    for (int c = j; c < k; c += blockDim.x) perm[c] = c;
    QPN_SYNC();
    int rank = 0;
    for (int c = 0; c < steps; ++c) {
        double best = -1.0; int jb = -1;
        for (int jj = j; jj < k; jj += blockDim.x) {
            if (jj < c) continue;
            double s = 0.0;
            for (int i = c; i < nd; ++i) { const double a = Ab[(size_t)i * ldk + jj]; s = fma(a, a, s); }
            if (jb < 0 || s > best) { best = s; jb = jj; }
        }
        block_argmax_idx(red, jb >= 0, best, jb);
        const double nrm = sqrt(best);
        if (nrm <= 1e-10) break;
        QPN_SYNC();
        if (jb != c) {
            for (int i = j; i < nd; i += blockDim.x) {
                const double tmp = Ab[(size_t)i * ldk + c];
                Ab[(size_t)i * ldk + c] = Ab[(size_t)i * ldk + jb];
                Ab[(size_t)i * ldk + jb] = tmp;
            }
            if (j == 0) { const int tp = perm[c]; perm[c] = perm[jb]; perm[jb] = tp; }
        }
        QPN_SYNC();
        const double alpha = Ab[(size_t)c * ldk + c] > 0.0 ? -nrm : nrm;
        for (int i = c + j; i < nd; i += blockDim.x) v[i] = Ab[(size_t)i * ldk + c] - (i == c ? alpha : 0.0);
        QPN_SYNC();
        double vn = 0.0;
        for (int i = c; i < nd; ++i) vn = fma(v[i], v[i], vn);
        if (vn > 0.0) {
            for (int jj = j; jj < k + 1; jj += blockDim.x) {
                if (jj < c) continue;
                if (jj < k) {
                    double s = 0.0;
                    for (int i = c; i < nd; ++i) s = fma(v[i], Ab[(size_t)i * ldk + jj], s);
                    s = (2.0 * s) / vn;
                    for (int i = c; i < nd; ++i) Ab[(size_t)i * ldk + jj] = fma(-s, v[i], Ab[(size_t)i * ldk + jj]);
                } else {                                    // the rhs rides as column k
                    double s = 0.0;
                    for (int i = c; i < nd; ++i) s = fma(v[i], b[i], s);
                    s = (2.0 * s) / vn;
                    for (int i = c; i < nd; ++i) b[i] = fma(-s, v[i], b[i]);
                }
            }
        }
        rank++;
        QPN_SYNC();
    }
    QPN_SYNC();
    if (j == 0) {
        for (int t = 0; t < k; ++t) lam[t] = 0.0;
        for (int i = rank - 1; i >= 0; --i) {
            double acc = b[i];
            for (int t = i + 1; t < rank; ++t) acc = fma(-Ab[(size_t)i * ldk + t], v[t], acc);
            v[i] = acc / Ab[(size_t)i * ldk + i];
        }
        for (int i = 0; i < rank; ++i) lam[perm[i]] = v[i];
    }
    QPN_SYNC();
}

// verify_solution for one node; qt / ax complete (barrier) on entry; lam in vs.lam_out().
// t: the CTA's big tableau (fallback AVI of size m and reduction scratch); vs.Ab / vs.Ab0 point into
// the same global slot and are dead by the time the fallback builds its tableau there.
__device__ __noinline__ int verify_solution_big(BigTab& t, VerifyBig& vs, const NodeDesc& nd_, const double* qt, const double* ax,
                                                double tol, int* how, int* pivots) {
    const int nd = nd_.nd, m = nd_.m, i = threadIdx.x;
    int* red_i = t.v.red_i();
    int infeasible = 0;
    for (int r = i; r < m; r += blockDim.x) {
        const double acc = ax[r];
        vs.lam_out()[r] = 0.0;
        const double lo = nd_.l[r], up = nd_.u[r];
        if (!((lo - 1e-3 <= acc) && (acc - 1e-3 <= up))) infeasible = 1;
        const bool pos = acc < lo + 1e-2, neg = acc > up - 1e-2;
        vs.kind()[r] = (pos && neg) ? 3 : pos ? 1 : neg ? 2 : 0;
    }
    infeasible = QPN_SYNC_OR(infeasible);
    if (infeasible) { *how = 0; return 0; }
    double nq = 0.0;
    for (int r = 0; r < nd; ++r) nq = fma(qt[r], qt[r], nq);
    if (m == 0) { *how = 1; return sqrt(nq) <= tol ? 1 : 0; }
    if (i == 0) {
        int k = 0, np_ = 0, nn = 0;
        for (int r = 0; r < m; ++r) if (vs.kind()[r] == 1) { vs.idx()[k++] = r; np_++; }
        for (int r = 0; r < m; ++r) if (vs.kind()[r] == 2) { vs.idx()[k++] = r; nn++; }
        for (int r = 0; r < m; ++r) if (vs.kind()[r] == 3) { vs.idx()[k++] = r; }
        red_i[32] = k; red_i[33] = np_; red_i[34] = nn;
    }
    QPN_SYNC();
    const int k = red_i[32], np_ = red_i[33], nn = red_i[34];
    if (k == 0) {
        QPN_SYNC();
        if (sqrt(nq) <= tol) { *how = 2; return 1; }
        *how = 4;
        return 0;
    }
    const int ldk = k;
    for (int e = i; e < nd * k; e += blockDim.x) {
        const int r = e / k, tcol = e - r * k;
        const double sgn = (tcol >= np_ && tcol < np_ + nn) ? -1.0 : 1.0;
        const double val = sgn * nd_.A[(size_t)nd_.dec[r] * m + vs.idx()[tcol]];
        vs.Ab[e] = val; vs.Ab0[e] = val;
    }
    for (int r = i; r < nd; r += blockDim.x) vs.b()[r] = qt[r];
    QPN_SYNC();
    lstsq_basic_big(t.v, nd, k, ldk, vs.Ab, vs.b(), vs.lam(), vs.perm(), vs.v());
    // acceptance (qp_processing.jl:119): residual entries in parallel, their squares summed in order
    for (int r = i; r < nd; r += blockDim.x) {
        double acc = 0.0;
        for (int tt = 0; tt < k; ++tt) acc = fma(vs.Ab0[(size_t)r * ldk + tt], vs.lam()[tt], acc);
        vs.v()[r] = acc - qt[r];
    }
    QPN_SYNC();
    if (i == 0) {
        int ok = 1;
        for (int tt = 0; tt < np_ + nn; ++tt) if (!(vs.lam()[tt] > -tol)) ok = 0;
        double res = 0.0;
        for (int r = 0; r < nd; ++r) { const double e = vs.v()[r]; res = fma(e, e, res); }
        if (!(sqrt(res) <= tol)) ok = 0;
        red_i[35] = ok;
    }
    QPN_SYNC();
    if (red_i[35]) {
        for (int tt = i; tt < k; tt += blockDim.x)
            vs.lam_out()[vs.idx()[tt]] = (tt >= np_ && tt < np_ + nn) ? -vs.lam()[tt] : vs.lam()[tt];
        QPN_SYNC();
        *how = 2;
        return 1;
    }
    // fallback (qp_processing.jl:129-146): (Ad Ad') lam - Ad qt  comp.  lb <= lam <= ub
    QPN_SYNC();
    for (int r = i; r < m; r += blockDim.x) {
        double acc = 0.0;
        for (int tt = 0; tt < nd; ++tt) acc = fma(nd_.A[(size_t)nd_.dec[tt] * m + r], qt[tt], acc);
        vs.qs()[r] = -acc;
        vs.zs()[r] = 0.0;
        const int8_t kd = vs.kind()[r];
        t.l()[r] = (kd == 2 || kd == 3) ? -QPN_INF : 0.0;
        t.u()[r] = (kd == 1 || kd == 3) ? QPN_INF : 0.0;
    }
    QPN_SYNC();
    auto build = [&](BigTab& tt) {
        const int ldr = tt.ldr;
        for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
            const int r = e / m, c = e - r * m;
            double acc = 0.0;
            for (int q = 0; q < nd; ++q) acc = fma(nd_.A[(size_t)nd_.dec[q] * m + r], nd_.A[(size_t)nd_.dec[q] * m + c], acc);
            tt.Tg[(size_t)r * ldr + c] = -acc;
        }
        QPN_SYNC();
    };
    const int st = solve_avi_big(t, m, nullptr, build, nullptr, vs.qs(), vs.zs(), vs.zb(), 50 * m + 100, nullptr, pivots);
    QPN_SYNC();
    if (st != ST_SUCCESS) { *how = 5; return 0; }
    for (int tt = i; tt < nd; tt += blockDim.x) {
        double acc = 0.0;
        for (int r = 0; r < m; ++r) acc = fma(nd_.A[(size_t)nd_.dec[tt] * m + r], vs.zs()[r], acc);
        vs.v()[tt] = acc - qt[tt];
    }
    for (int r = i; r < m; r += blockDim.x) vs.lam_out()[r] = vs.zs()[r];
    QPN_SYNC();
    if (i == 0) {
        double res2 = 0.0;
        for (int tt = 0; tt < nd; ++tt) { const double e = vs.v()[tt]; res2 = fma(e, e, res2); }
        red_i[35] = sqrt(res2) <= 1e-4 ? 1 : 0;
    }
    QPN_SYNC();
    const int ok2 = red_i[35];
    *how = ok2 ? 3 : 4;
    return ok2;
}

// Doubles of the slot a verify_solution call needs: max(2 nd m, m (m+1) tableau).
__host__ __device__ __forceinline__ size_t verify_big_slot_doubles(int nd, int m) {
    const size_t ab = 2 * (size_t)nd * (m > 0 ? m : 1), tb = big_slot_doubles(m > 0 ? m : 1);
    return ab > tb ? ab : tb;
}

// grid = persistent CTAs, block = QPN_BIG_THREADS.  Dynamic smem: big_smem_bytes(max(m,1)) + verify_big_bytes + 8 (nv + nd + m).
__global__ void __launch_bounds__(QPN_BIG_THREADS, 1)
verify_solution_big_kernel(const __grid_constant__ NodeDesc node, int batch, const double* __restrict__ x, double tol,
                           uint8_t* __restrict__ solution_out, double* __restrict__ lam_out, int32_t* __restrict__ how_out,
                           int8_t* __restrict__ active_out, double* __restrict__ work, size_t slot_doubles, int smem_used) {
    const int m = node.m, tn = m > 0 ? m : 1;
    BigTab t;
    double* slot = work + (size_t)blockIdx.x * slot_doubles;
    int off = big_carve(t, tn, slot, 0);
    big_stage_carve(t, smem_used);
    VerifyBig vs;
    vs.base = off; vs.nd = node.nd; vs.m = m; vs.Ab = slot; vs.Ab0 = slot + (size_t)node.nd * tn;
    off += (int)verify_big_bytes(node.nd, m);
    double* xs = reinterpret_cast<double*>(qpn_smem + off);
    double* qt = xs + node.nv;
    double* ax = qt + node.nd;
    for (int b = blockIdx.x; b < batch; b += gridDim.x) {
        for (int j = threadIdx.x; j < node.nv; j += blockDim.x) xs[j] = x[(size_t)b * node.nv + j];
        QPN_SYNC();
        node_products(node, xs, qt, ax);
        QPN_SYNC();
        int how = 0, piv = 0;
        const int sol = verify_solution_big(t, vs, node, qt, ax, tol, &how, &piv);
        QPN_SYNC();
        for (int r = threadIdx.x; r < m; r += blockDim.x) {
            if (lam_out) lam_out[(size_t)b * m + r] = vs.lam_out()[r];
            if (active_out) active_out[(size_t)b * m + r] = how == 0 ? 0 : vs.kind()[r];
        }
        if (threadIdx.x == 0) { solution_out[b] = (uint8_t)sol; if (how_out) how_out[b] = how; }
        QPN_SYNC();
    }
}

// ---- fused per-level equilibrium loop, big form (algorithm.jl:13-118) ----------------------------
// Shared memory: big vectors for n_level, the GAVI extras (the verify vectors alias them: a verify never
// overlaps a solve), then xs, pv, xn, qt_all, ax_all.
__host__ __device__ __forceinline__ size_t level_big_smem_bytes(const LevelDesc& lv) {
    const int n = lv.g.d1 + 2 * lv.g.d2;
    size_t extra = gavi_extra_bytes(lv.g.d1, lv.g.d2, lv.g.np), vb = verify_big_bytes(lv.max_nd, lv.max_m);
    if (vb > extra) extra = vb;
    return big_smem_bytes(n > lv.max_m ? n : lv.max_m) + extra +
           8 * (2 * (size_t)lv.nv + (size_t)(lv.nproj > 0 ? lv.nproj : 1) + (size_t)lv.nd_total + (size_t)lv.lam_total);
}
__host__ __device__ __forceinline__ size_t level_big_slot_doubles(const LevelDesc& lv) {
    const int n = lv.g.d1 + 2 * lv.g.d2;
    const size_t a = big_slot_doubles(n), b = verify_big_slot_doubles(lv.max_nd, lv.max_m);
    return a > b ? a : b;
}

// Compact form of the slot for a level whose two solves start from plans: the tableau area only has to hold the rows
// an instance sweeps (PlanDesc::nact of them) and the small no-plan tableau of verify_solution's fallback.  0 when
// the level has no such plans.  When that slot fits shared memory next to the vectors, the level kernel keeps it
// there (several CTAs of a few warps per SM instead of one CTA of 1,024 threads around a slot in global memory).
__host__ __device__ __forceinline__ size_t level_big_compact_tcap(const LevelDesc& lv) {
    if (!lv.has_plans || lv.planA.nact >= lv.planA.n || lv.planB.n <= 0) return 0;
    size_t a = (size_t)lv.planA.nact * row_stride(lv.planA.ncol0), b = (size_t)lv.planB.nact * row_stride(lv.planB.ncol0);
    size_t c = (size_t)lv.max_m * row_stride(lv.max_m + 1);
    size_t tcap = a > b ? a : b;
    if (c > tcap) tcap = c;
    return (tcap + 1) & ~(size_t)1;
}
__host__ __device__ __forceinline__ size_t level_big_compact_slot_doubles(const LevelDesc& lv, size_t tcap) {
    const int n = lv.g.d1 + 2 * lv.g.d2;
    const int nmax = n > lv.max_m ? n : lv.max_m;
    const size_t a = big_slot_doubles_ex(nmax, tcap), b = 2 * (size_t)lv.max_nd * (lv.max_m > 0 ? lv.max_m : 1);
    return ((a > b ? a : b) + 1) & ~(size_t)1;
}

// work == nullptr: the slot (slot_doubles, tableau area tcap) lives in this CTA's dynamic shared memory, right after
// the smem_used bytes of the layout below.
__global__ void __launch_bounds__(QPN_BIG_THREADS, 1)
level_equilibrium_big_kernel(const __grid_constant__ LevelDesc lv, int batch, const double* __restrict__ x_init,
                             double* __restrict__ x_out, uint8_t* __restrict__ solved_out, int32_t* __restrict__ iters_out,
                             int32_t* __restrict__ pivots_out, double* __restrict__ lam_out, double* __restrict__ hist,
                             int32_t* __restrict__ hist_count, int hist_cap, int presolve, int hist_fresh, double* __restrict__ work,
                             size_t slot_doubles, int smem_used, size_t tcap) {
    const int i = threadIdx.x, nv = lv.nv;
    const int n_level = lv.g.d1 + 2 * lv.g.d2;
    const int nmax = n_level > lv.max_m ? n_level : lv.max_m;
    BigTab t;
    const int slot_off = (smem_used + 15) & ~15;
    double* slot = work ? work + (size_t)blockIdx.x * slot_doubles : reinterpret_cast<double*>(qpn_smem + slot_off);
    int off = big_carve(t, nmax, slot, 0, tcap);
    big_stage_carve(t, work ? smem_used : slot_off + (int)(8 * slot_doubles));
    GaviSmem gs;
    gavi_carve_extra(gs, lv.g, off);
    VerifyBig vs;
    vs.base = off; vs.nd = lv.max_nd; vs.m = lv.max_m; vs.Ab = slot; vs.Ab0 = slot + (size_t)lv.max_nd * lv.max_m;
    {
        size_t extra = gavi_extra_bytes(lv.g.d1, lv.g.d2, lv.g.np), vb = verify_big_bytes(lv.max_nd, lv.max_m);
        off += (int)(vb > extra ? vb : extra);
    }
    double* xs = reinterpret_cast<double*>(qpn_smem + off);
    double* pv = xs + nv;
    double* xn = pv + (lv.nproj > 0 ? lv.nproj : 1);
    double* qt_all = xn + nv;
    double* ax_all = qt_all + lv.nd_total;
    const int max_piv = 50 * n_level + 100;
    const int rows_all = lv.nd_total + lv.lam_total;
    for (int b = blockIdx.x; b < batch; b += gridDim.x) {
        for (int j = i; j < nv; j += blockDim.x) xs[j] = x_init[(size_t)b * nv + j];
        QPN_SYNC();
        int solved = 0, piv = 0, iters = 0;
        double* myhist = hist ? hist + (size_t)b * hist_cap * lv.nproj : nullptr;
        int nhist = (hist && hist_count && !hist_fresh) ? hist_count[b] : 0;
        for (int it = 1; it <= lv.max_iters; ++it) {
            iters = it;
            if (lv.nproj > 0 && myhist) {
                for (int k = i; k < lv.nproj; k += blockDim.x) {
                    double acc = 0.0;
                    for (int j = 0; j < nv; ++j) acc = fma(xs[j], lv.proj[(size_t)k * nv + j], acc);
                    pv[k] = acc;
                }
                QPN_SYNC();
                int cyc = 0;
                for (int h = i; h < nhist; h += blockDim.x) {
                    const double* ph = myhist + (size_t)h * lv.nproj;
                    double dd = 0.0, na = 0.0, nb2 = 0.0;
                    for (int k = 0; k < lv.nproj; ++k) {
                        const double e = pv[k] - ph[k];
                        dd = fma(e, e, dd); na = fma(pv[k], pv[k], na); nb2 = fma(ph[k], ph[k], nb2);
                    }
                    if (sqrt(dd) <= 1.4901161193847656e-8 * fmax(sqrt(na), sqrt(nb2))) cyc = 1;
                }
                cyc = QPN_SYNC_OR(cyc);
                if (cyc) break;
                if (nhist < hist_cap) {
                    for (int k = i; k < lv.nproj; k += blockDim.x) myhist[(size_t)nhist * lv.nproj + k] = pv[k];
                    nhist++;
                }
                QPN_SYNC();
            }
            // process_qp for every player (algorithm.jl:47-49)
            for (int r = i; r < rows_all; r += blockDim.x) {
                int pl = 0;
                if (r < lv.nd_total) {
                    while (pl + 1 < lv.nplayers && r >= lv.nd_off[pl + 1]) ++pl;
                    const NodeDesc& node = lv.players[pl];
                    const int rr = r - lv.nd_off[pl];
                    double acc = 0.0;
                    for (int j = 0; j < nv; ++j) acc = fma(node.Qd[(size_t)j * node.nd + rr], xs[j], acc);
                    qt_all[r] = acc + node.qd[rr];
                } else {
                    const int q = r - lv.nd_total;
                    while (pl + 1 < lv.nplayers && q >= lv.m_off[pl + 1]) ++pl;
                    const NodeDesc& node = lv.players[pl];
                    const int rr = q - lv.m_off[pl];
                    double acc = 0.0;
                    for (int j = 0; j < nv; ++j) acc = fma(node.A[(size_t)j * node.m + rr], xs[j], acc);
                    ax_all[q] = acc;
                }
            }
            QPN_SYNC();
            int all_sol = 1;
            for (int pl = 0; pl < lv.nplayers; ++pl) {
                const NodeDesc& node = lv.players[pl];
                int how = 0;
                vs.nd = node.nd; vs.m = node.m; vs.Ab0 = slot + (size_t)node.nd * (node.m > 0 ? node.m : 1);
                const int sol = verify_solution_big(t, vs, node, qt_all + lv.nd_off[pl], ax_all + lv.m_off[pl], 1e-4, &how, &piv);
                QPN_SYNC();
                if (lam_out)
                    for (int r = i; r < node.m; r += blockDim.x)
                        lam_out[(size_t)b * lv.lam_total + lv.m_off[pl] + r] = sol ? vs.lam_out()[r] : 0.0;
                if (!sol) all_sol = 0;
                QPN_SYNC();
            }
            if (all_sol) { solved = 1; break; }
            // solve_qep (avi.jl:382-444)
            for (int j = i; j < lv.g.np; j += blockDim.x) gs.w()[j] = xs[lv.par[j]];
            for (int j = i; j < lv.g.d1 + lv.g.d2; j += blockDim.x) gs.z0()[j] = j < lv.nd_level ? xs[lv.dec[j]] : 0.0;
            QPN_SYNC();
            const int st = gavi_solve_big(t, gs, lv.g, lv.has_plans ? &lv.planA : nullptr, lv.has_plans ? &lv.planB : nullptr, presolve,
                                          max_piv, &piv);
            QPN_SYNC();
            if (st != ST_SUCCESS) break;
            for (int j = i; j < nv; j += blockDim.x) xn[j] = xs[j];
            QPN_SYNC();
            for (int j = i; j < lv.nd_level; j += blockDim.x) xn[lv.dec[j]] = gs.zs()[j];
            QPN_SYNC();
            double dn = 0.0;
            for (int j = 0; j < nv; ++j) { const double e = xn[j] - xs[j]; dn = fma(e, e, dn); }
            if (sqrt(dn) < 1e-4) break;
            QPN_SYNC();
            for (int j = i; j < nv; j += blockDim.x) xs[j] = xn[j];
            QPN_SYNC();
        }
        for (int j = i; j < nv; j += blockDim.x) x_out[(size_t)b * nv + j] = xs[j];
        if (i == 0) {
            solved_out[b] = (uint8_t)solved; iters_out[b] = iters; pivots_out[b] = piv;
            if (hist && hist_count) hist_count[b] = nhist;
        }
        QPN_SYNC();
    }
}

}  // namespace qpn
