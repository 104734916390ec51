// solve_avi / solve_gavi for sizes beyond the shared-memory tableau (avi_pivot_big.cuh): persistent
// CTAs, one global-memory tableau slot per CTA, instances taken round-robin.
#pragma once
#include "avi_pivot_big.cuh"
#include "qpn_level.cuh"

namespace qpn {

#ifndef QPN_BIG_THREADS_N
#define QPN_BIG_THREADS_N 1024
#endif
constexpr int QPN_BIG_THREADS = QPN_BIG_THREADS_N;

// ---- builders: Tg[i][0:n] = -M[i][:]; each ends with a barrier -----------------------------------
__device__ __forceinline__ void big_zero(BigTab& t, int n) {
    const int ldr = t.ldr, lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = w; i < n; i += nw) {
        double* row = t.Tg + (size_t)i * ldr;
        for (int j = lane; j < ldr; j += 32) row[j] = 0.0;
    }
    QPN_SYNC();
}

__device__ __noinline__ void big_build_matrix(BigTab& t, const MatDesc& M, int b) {
    const int n = t.n, ldr = t.ldr;
    if (M.dense) {
        const double* src = M.dense + (M.shared ? 0 : (size_t)b * n * n);
        // warps take rows, lanes take columns: strided reads of M (L2-resident when shared), coalesced writes
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
        for (int i = w; i < n; i += nw) {
            double* row = t.Tg + (size_t)i * ldr;
            for (int j = lane; j < n; j += 32) row[j] = -src[(size_t)j * n + i];
        }
    } else {
        big_zero(t, n);
        const double* nz = M.nzval + (M.shared ? 0 : (size_t)b * M.nnz);
        for (int j = 0; j < n; ++j) {
            const int k0 = M.colptr[j] - M.base, k1 = M.colptr[j + 1] - M.base;
            for (int k = k0 + threadIdx.x; k < k1; k += blockDim.x)
                t.Tg[(size_t)(M.rowval[k] - M.base) * ldr + j] = -nz[k];
        }
    }
    QPN_SYNC();
}

// -M of the lifted AVI of `convert` (avi.jl:113-128): M = [M 0; A 0 -I; 0 I 0].
__device__ __noinline__ void big_build_lifted(BigTab& t, const GaviDesc& g) {
    const int ldr = t.ldr, d1 = g.d1, d2 = g.d2, dz = d1 + d2, n = d1 + 2 * d2;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int r = w; r < n; r += nw) {
        double* row = t.Tg + (size_t)r * ldr;
        if (r < d1) {
            for (int j = lane; j < n; j += 32) row[j] = j < dz ? -g.M[(size_t)j * d1 + r] : 0.0;
        } else if (r < dz) {
            for (int j = lane; j < n; j += 32) row[j] = j < dz ? -g.A[(size_t)j * d2 + (r - d1)] : ((j - dz == r - d1) ? 1.0 : 0.0);
        } else {
            for (int j = lane; j < n; j += 32) row[j] = (j - d1 == r - dz) ? -1.0 : 0.0;
        }
    }
    QPN_SYNC();
}

// -M of the presolve AVI (avi.jl:79-99): M = [I -A' 0; A 0 -I; 0 I 0] over [z(k); lambda(d2); s(d2)].
__device__ __noinline__ void big_build_presolve(BigTab& t, const GaviDesc& g, const int* cols, int k) {
    const int ldr = t.ldr, d2 = g.d2, pn = k + 2 * d2;
    big_zero(t, pn);
    for (int e = threadIdx.x; e < k * d2; e += blockDim.x) {
        const int a = e / d2, r = e - a * d2;
        const double v = g.A[(size_t)cols[a] * d2 + r];
        t.Tg[(size_t)a * ldr + (k + r)] = v;
        t.Tg[(size_t)(k + r) * ldr + a] = -v;
    }
    for (int e = threadIdx.x; e < k; e += blockDim.x) t.Tg[(size_t)e * ldr + e] = -1.0;
    for (int e = threadIdx.x; e < d2; e += blockDim.x) {
        t.Tg[(size_t)(k + e) * ldr + (k + d2 + e)] = 1.0;
        t.Tg[(size_t)(k + d2 + e) * ldr + (k + e)] = -1.0;
    }
    QPN_SYNC();
}

// ---- start from a plan (same state as big_start + phase 0 + recompute_tcol + compact_dead) --------
__device__ __noinline__ void big_start_plan(BigTab& t, const PlanDesc& P, const double* q, const double* z0, double* zb) {
    const int n = P.n;
    big_shape(t, n, P.ncol0);
    const int ldr = t.ldr;
    for (int i = threadIdx.x; i < n; i += blockDim.x) zb[i] = fmin(fmax(z0[i], t.l()[i]), t.u()[i]);
    const int nact = P.nact;                              // rows nact .. n-1 are frozen: never copied, read from the plan at the end
    {
        const double2* src = reinterpret_cast<const double2*>(P.T0);
        double2* dst = reinterpret_cast<double2*>(t.Tg);
        const size_t cnt = (size_t)nact * ldr / 2;
        for (size_t e = threadIdx.x; e < cnt; e += blockDim.x) dst[e] = src[e];
    }
    for (int v = threadIdx.x; v <= 2 * n; v += blockDim.x) { t.rowof()[v] = -1; t.colof()[v] = -1; }
    QPN_SYNC();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double zi = z0[i];
        t.rr()[i] = ((csr_row_dot(P, i, zb) + q[i]) + zi) - zb[i];
        t.zst()[i] = (zi <= t.l()[i]) ? AT_L : (zi >= t.u()[i]) ? AT_U : FLOATING;
    }
    for (int j = threadIdx.x; j < P.ncol0; j += blockDim.x) {
        const int v = P.colvar0[j];
        t.colvar()[j] = v; t.colof()[v] = j; t.nbval()[j] = v < n ? zb[v] : 0.0;
    }
    QPN_SYNC();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (i < nact) {
            double acc = 0.0;
            for (int k = 0; k < n; ++k) {
                const double pik = P.PT[(size_t)k * n + i];
                if (pik != 0.0) acc = fma(pik, t.rr()[k], acc);
            }
            t.Tg[(size_t)i * ldr + P.tcol0] = acc;
        }
        const int rv = P.rowvar0[i];
        t.rowvar()[i] = rv; t.rowof()[rv] = i;
        t.beta()[i] = rv < n ? zb[rv] : zb[rv - n] - z0[rv - n];      // (a plan may export its rows in another order: go by the variable)
        if (rv < n) t.zst()[rv] = BASIC;
    }
    t.ncol = P.ncol0; t.pivots = P.npiv0; t.cc = -1; t.npend = 0;
    t.nact = nact; t.tcol0 = P.tcol0; t.T0f = nact < n ? P.T0 : nullptr; t.PTf = P.PT;
    QPN_SYNC();
}

// ---- one AVI solve on the big tableau -------------------------------------------------------------
// P != null: start from the plan and check with its CSR rows.  Else `build(t)` fills Tg with -M
// (called again for the final check) unless Md -- the dense column-major matrix -- is given, in
// which case products read it directly.  t.l() / t.u(): bounds; qs: q; zs: start on entry, z on exit.
template <class Build>
__device__ __forceinline__ int solve_avi_big(BigTab& t, int n, const PlanDesc* P, Build build, const double* Md, const double* qs,
                                             double* zs, double* zb, int max_pivots, int8_t* code, int* pivots_acc) {
    if (P) big_start_plan(t, *P, qs, zs, zb);
    else { big_shape(t, n, n + 1); build(t); big_start(t, Md, qs, zs); }
    int st = avi_pivot_run_big(t, max_pivots, P != nullptr, zs, code);
    *pivots_acc += t.pivots;
    int bad = 0;
    if (P) {
        for (int i = threadIdx.x; i < n; i += blockDim.x)
            bad += check_avi_index(csr_row_dot(*P, i, zs) + qs[i], zs[i], t.l()[i], t.u()[i], 1e-6);
    } else {
        if (!Md) build(t);
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            double acc = 0.0;
            if (Md) {
                for (int j = 0; j < n; ++j) { const double mij = Md[(size_t)j * n + i]; if (mij != 0.0) acc = fma(mij, zs[j], acc); }
            } else {
                const double* row = t.Tg + (size_t)i * t.ldr;
                for (int j = 0; j < n; ++j) { const double mij = -row[j]; if (mij != 0.0) acc = fma(mij, zs[j], acc); }
            }
            bad += check_avi_index(acc + qs[i], zs[i], t.l()[i], t.u()[i], 1e-6);
        }
    }
    bad = QPN_SYNC_OR(bad);
    if (st == ST_SUCCESS && bad) st = ST_FAILURE;
    return st;
}

// ---- solve_avi (avi.jl:63-77), big form -------------------------------------------------------------
// grid = resident CTAs (<= slots), block = QPN_BIG_THREADS.  Dynamic smem: big_smem_bytes(n) + 3n doubles + n bytes.
__global__ void __launch_bounds__(QPN_BIG_THREADS, 1)
avi_solve_big_kernel(int n, int batch, const __grid_constant__ MatDesc M, const __grid_constant__ PlanDesc P, int has_plan,
                     const double* __restrict__ q, const double* __restrict__ l, const double* __restrict__ u, int lu_shared,
                     const double* __restrict__ z0, int max_pivots, double* __restrict__ z_out, int32_t* __restrict__ status_out,
                     int32_t* __restrict__ pivots_out, int8_t* __restrict__ basis_out, double* __restrict__ work, size_t slot_doubles,
                     int smem_used) {
    BigTab t;
    const int off = big_carve(t, n, work + (size_t)blockIdx.x * slot_doubles, 0);
    big_stage_carve(t, smem_used);
    double* qs = reinterpret_cast<double*>(qpn_smem + off);
    double* zs = qs + n;
    double* zb = zs + n;
    int8_t* code = reinterpret_cast<int8_t*>(zb + n);
    for (int b = blockIdx.x; b < batch; b += gridDim.x) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            qs[i] = q[(size_t)b * n + i];
            zs[i] = z0[(size_t)b * n + i];
            t.l()[i] = l[(lu_shared ? 0 : (size_t)b * n) + i];
            t.u()[i] = u[(lu_shared ? 0 : (size_t)b * n) + i];
        }
        QPN_SYNC();
        int piv = 0;
        const double* Md = M.dense ? M.dense + (M.shared ? 0 : (size_t)b * n * n) : nullptr;
        const int st = solve_avi_big(t, n, has_plan ? &P : nullptr, [&](BigTab& tt) { big_build_matrix(tt, M, b); }, Md, qs, zs, zb,
                                     max_pivots, code, &piv);
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            z_out[(size_t)b * n + i] = zs[i];
            if (basis_out) basis_out[(size_t)b * n + i] = code[i];
        }
        if (threadIdx.x == 0) { status_out[b] = st; pivots_out[b] = piv; }
        QPN_SYNC();
    }
}

// ---- solve_gavi (avi.jl:101-111), big form ----------------------------------------------------------
__device__ __forceinline__ int gavi_solve_big(BigTab& t, GaviSmem& s, const GaviDesc& g, const PlanDesc* planA, const PlanDesc* planB,
                                              int presolve, int max_pivots, int* pivots) {
    const int d1 = g.d1, d2 = g.d2, dz = d1 + d2, n = d1 + 2 * d2;
    gavi_slack(s, g, true);
    if (presolve && d2 > 0) {
        int infeasible = 0;
        for (int r = threadIdx.x; r < d2; r += blockDim.x)
            if (!(g.l2[r] <= s.s0()[r] && s.s0()[r] <= g.u2[r])) infeasible = 1;
        infeasible = QPN_SYNC_OR(infeasible);
        if (infeasible) {
            const int* cols;
            int k;
            if (planB) { cols = planB->cols; k = planB->ncols; }
            else { find_cols(g, s.cols()); cols = s.cols(); k = s.cols()[dz]; }
            const int pn = k + 2 * d2;
            for (int i = threadIdx.x; i < pn; i += blockDim.x) {
                if (i < k) { s.qs()[i] = -s.z0()[cols[i]]; s.zs()[i] = s.z0()[cols[i]]; t.l()[i] = -QPN_INF; t.u()[i] = QPN_INF; }
                else if (i < k + d2) {
                    const int r = i - k;
                    double full = 0.0, part = 0.0;
                    for (int j = 0; j < dz; ++j) full = fma(g.A[(size_t)j * d2 + r], s.z0()[j], full);
                    for (int a = 0; a < k; ++a) part = fma(g.A[(size_t)cols[a] * d2 + r], s.z0()[cols[a]], part);
                    s.qs()[i] = (full - part) + s.c()[r];
                    s.zs()[i] = 0.0; t.l()[i] = -QPN_INF; t.u()[i] = QPN_INF;
                } else {
                    const int r = i - k - d2;
                    s.qs()[i] = 0.0; s.zs()[i] = s.s0()[r]; t.l()[i] = g.l2[r]; t.u()[i] = g.u2[r];
                }
            }
            QPN_SYNC();
            const int pst = solve_avi_big(t, pn, planB, [&](BigTab& tt) { big_build_presolve(tt, g, cols, k); }, nullptr, s.qs(), s.zs(),
                                          s.zb(), 50 * pn + 100, nullptr, pivots);
            QPN_SYNC();
            if (pst == ST_SUCCESS)
                for (int i = threadIdx.x; i < k; i += blockDim.x) s.z0()[cols[i]] = s.zs()[i];
            QPN_SYNC();
            gavi_slack(s, g, false);
        }
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (i < d1) {
            double acc = 0.0;
            for (int j = 0; j < g.np; ++j) acc = fma(g.N[(size_t)j * d1 + i], s.w()[j], acc);
            s.qs()[i] = acc + g.o[i];
            t.l()[i] = g.l1[i]; t.u()[i] = g.u1[i];
        } else if (i < dz) {
            s.qs()[i] = s.c()[i - d1]; t.l()[i] = -QPN_INF; t.u()[i] = QPN_INF;
        } else {
            s.qs()[i] = 0.0; t.l()[i] = g.l2[i - dz]; t.u()[i] = g.u2[i - dz];
        }
        s.zs()[i] = i < dz ? s.z0()[i] : s.s0()[i - dz];
    }
    QPN_SYNC();
    return solve_avi_big(t, n, planA, [&](BigTab& tt) { big_build_lifted(tt, g); }, nullptr, s.qs(), s.zs(), s.zb(), max_pivots,
                         s.code(), pivots);
}

// Dynamic smem: big_smem_bytes(n) + gavi_extra_bytes.
__global__ void __launch_bounds__(QPN_BIG_THREADS, 1)
gavi_solve_big_kernel(const __grid_constant__ GaviDesc g, const __grid_constant__ GaviPlans plans, int batch,
                      const double* __restrict__ w, const double* __restrict__ z0, int presolve, int max_pivots,
                      double* __restrict__ z_out, double* __restrict__ zfull_out, int32_t* __restrict__ status_out,
                      int32_t* __restrict__ pivots_out, int8_t* __restrict__ basis_out, double* __restrict__ work, size_t slot_doubles,
                      int smem_used) {
    const int dz = g.d1 + g.d2, n = g.d1 + 2 * g.d2;
    BigTab t;
    const int off = big_carve(t, n, work + (size_t)blockIdx.x * slot_doubles, 0);
    big_stage_carve(t, smem_used);
    GaviSmem s;
    gavi_carve_extra(s, g, off);
    for (int b = blockIdx.x; b < batch; b += gridDim.x) {
        for (int j = threadIdx.x; j < g.np; j += blockDim.x) s.w()[j] = w[(size_t)b * g.np + j];
        for (int j = threadIdx.x; j < dz; j += blockDim.x) s.z0()[j] = z0[(size_t)b * dz + j];
        QPN_SYNC();
        int piv = 0;
        const int st = gavi_solve_big(t, s, g, plans.has ? &plans.A : nullptr, plans.has ? &plans.B : nullptr, presolve, max_pivots, &piv);
        QPN_SYNC();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            if (i < dz) z_out[(size_t)b * dz + i] = s.zs()[i];
            if (zfull_out) zfull_out[(size_t)b * n + i] = s.zs()[i];
            if (basis_out) basis_out[(size_t)b * n + i] = s.code()[i];
        }
        if (threadIdx.x == 0) { status_out[b] = st; pivots_out[b] = piv; }
        QPN_SYNC();
    }
}

// ---- plan construction on the big tableau: one CTA, once per shared matrix -----------------------
// kind 0: lifted AVI of the GAVI; kind 1: its presolve AVI; kind 2: the plain AVI (M, l, u).
// Same exports as plan_finish (qpn_level.cuh).  Dynamic smem: big_smem_bytes(n) + gavi extras (kinds 0, 1)
// or 2n doubles (kind 2), + n ints.
__global__ void __launch_bounds__(QPN_BIG_THREADS, 1)
plan_build_big_kernel(const __grid_constant__ GaviDesc g, int kind, int n_avi, const __grid_constant__ MatDesc M,
                      const double* __restrict__ l, const double* __restrict__ u, double* __restrict__ T0, double* __restrict__ PT,
                      int* __restrict__ rowvar0, int* __restrict__ colvar0, int* __restrict__ csr_ptr, int* __restrict__ csr_col,
                      double* __restrict__ csr_val, int* __restrict__ cols_out, int* __restrict__ hdr, double* __restrict__ work,
                      int smem_used) {
    const int d1 = g.d1, d2 = g.d2, dz = d1 + d2;
    const int nmax = kind == 2 ? n_avi : d1 + 2 * d2;
    BigTab t;
    int off = big_carve(t, nmax, work, 0);
    big_stage_carve(t, smem_used);
    GaviSmem s;
    double *qs, *zs;
    int* cnt;
    int n, k = 0;
    if (kind == 2) {
        qs = reinterpret_cast<double*>(qpn_smem + off); zs = qs + nmax; cnt = reinterpret_cast<int*>(zs + nmax);
        n = n_avi;
        for (int i = threadIdx.x; i < n; i += blockDim.x) { t.l()[i] = l[i]; t.u()[i] = u[i]; }
    } else {
        off = gavi_carve_extra(s, g, off);
        qs = s.qs(); zs = s.zs(); cnt = reinterpret_cast<int*>(qpn_smem + off);
        if (kind == 1) {
            find_cols(g, s.cols());
            k = s.cols()[dz];
            n = k + 2 * d2;
            for (int j = threadIdx.x; j < k; j += blockDim.x) cols_out[j] = s.cols()[j];
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const bool fr = i < k + d2;
                t.l()[i] = fr ? -QPN_INF : g.l2[i - k - d2];
                t.u()[i] = fr ? QPN_INF : g.u2[i - k - d2];
            }
        } else {
            n = d1 + 2 * d2;
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                t.l()[i] = i < d1 ? g.l1[i] : i < dz ? -QPN_INF : g.l2[i - dz];
                t.u()[i] = i < d1 ? g.u1[i] : i < dz ? QPN_INF : g.u2[i - dz];
            }
        }
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) { qs[i] = 0.0; zs[i] = 0.0; }
    QPN_SYNC();
    big_shape(t, n, n + 1);
    if (kind == 2) big_build_matrix(t, M, 0);
    else if (kind == 1) big_build_presolve(t, g, s.cols(), k);
    else big_build_lifted(t, g);
    // the original matrix in CSR (rows ascending in the column index)
    const int ldr = t.ldr;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double* row = t.Tg + (size_t)i * ldr;
        int c = 0;
        for (int j = 0; j < n; ++j) c += (row[j] != 0.0);
        cnt[i] = c;
    }
    QPN_SYNC();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int r = 0; r < n; ++r) { csr_ptr[r] = acc; acc += cnt[r]; }
        csr_ptr[n] = acc;
    }
    QPN_SYNC();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double* row = t.Tg + (size_t)i * ldr;
        int o = csr_ptr[i];
        for (int j = 0; j < n; ++j) if (row[j] != 0.0) { csr_col[o] = j; csr_val[o] = -row[j]; ++o; }
    }
    QPN_SYNC();
    big_start(t, nullptr, qs, zs);
    // phase 0 exactly as crash() runs it
    for (int v = 0; v < n; ++v) {
        if (!is_free_var(t, v)) continue;
        const int c = t.colof()[v];
        const int rho = best_free_row(t, c);
        if (rho >= 0) { pivot(t, rho, c, false); set_zst(t, v, BASIC); }
    }
    big_flush(t);
    // Export order of the rows (used for B^-1 right below and for T0 / rowvar0 further down): the rows an instance
    // sweeps first, the rows of free basics (frozen from here on, avi_pivot.cuh: freeze) last, each group in its
    // original order -- every tie rule only compares swept rows, so their relative order is all that matters.  An
    // instance then copies rows 0 .. nact-1 and reads the frozen ones from the plan, once, at the end.
    int* perm = cnt;                                       // the CSR counts are done with
    if (threadIdx.x == 0) {
        int na = 0;
        for (int i = 0; i < n; ++i) if (t.rowvar()[i] >= n) perm[i] = na++;
        hdr[4] = na;
        for (int i = 0; i < n; ++i) if (t.rowvar()[i] < n) perm[i] = na++;
    }
    QPN_SYNC();
    // B^-1 from the slack columns (see recompute_tcol)
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double* row = t.Tg + (size_t)i * ldr;
        const int pi = perm[i];
        for (int kk = 0; kk < n; ++kk) {
            const int ck = t.colof()[n + kk];
            PT[(size_t)kk * n + pi] = ck >= 0 ? -row[ck] : (t.rowof()[n + kk] == i ? -1.0 : 0.0);
        }
    }
    const int npiv0 = t.pivots;
    QPN_SYNC();
    compact_dead(t);
    const int ncol0 = t.ncol, ldr0 = row_stride(ncol0);
    {
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
        for (int i = w; i < n; i += nw) {
            const double* row = t.Tg + (size_t)i * ldr;
            const size_t o = (size_t)perm[i] * ldr0;
            for (int j = lane; j < ldr0; j += 32) T0[o + j] = j < ncol0 ? row[j] : 0.0;
        }
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) rowvar0[perm[i]] = t.rowvar()[i];
    for (int j = threadIdx.x; j < ncol0; j += blockDim.x) colvar0[j] = t.colvar()[j];
    if (threadIdx.x == 0) { hdr[0] = ncol0; hdr[1] = npiv0; hdr[2] = t.colof()[2 * n]; hdr[3] = k; }
}

}  // namespace qpn
