// solve_gavi, verify_solution and the fused per-level equilibrium loop on sm_100a.
// One CTA per instance; everything an instance needs lives in shared memory.
#pragma once
#include "qpn_kernels.cuh"

namespace qpn {

// ---- plans: the instance-independent part of a solve, computed once per shared matrix ----------
// Phase 0 of the crash (free variables against rows of free variables) depends only on the
// matrix and on which variables are free.  plan_build_kernel runs it once and exports the
// exchanged tableau (live columns only), B^-1 (to rebuild the homotopy column from an instance's
// r), the basis maps and the original matrix in CSR (for r and for the final check).
struct PlanDesc {
    int n, ncol0, npiv0, tcol0;
    int nact;                // rows 0 .. nact-1 are swept; rows nact .. n-1 hold free basics (frozen, avi_pivot.cuh) and were
                             // exported last (stable order).  nact == n: rows in their original order.
    const double* T0;        // n x row_stride(ncol0), row-major
    const double* PT;        // n x n, PT[k*n + i] = (B^-1)[i][k]
    const int* rowvar0;      // n
    const int* colvar0;      // ncol0
    const int* csr_ptr;      // n + 1
    const int* csr_col;
    const double* csr_val;
    const int* cols;         // presolve plans: columns of A that are not structurally zero
    int ncols;
};

__device__ __forceinline__ double csr_row_dot(const PlanDesc& P, int i, const double* v) {
    double acc = 0.0;
    for (int k = P.csr_ptr[i]; k < P.csr_ptr[i + 1]; ++k) acc = fma(P.csr_val[k], v[P.csr_col[k]], acc);
    return acc;
}

// ---- TMA: the plan's exchanged tableau into the instance's shared-memory tableau ----------------------------------------
// The swept rows of T0 lie in global memory exactly as the instance keeps them in shared memory (same row stride, rows
// 0 .. nact-1 back to back), so the start of a solve is ONE bulk asynchronous copy (cp.async.bulk, completion counted in
// bytes on an mbarrier) issued by one thread; the other threads build the residual, the basis maps and B^-1 r meanwhile
// and wait on the barrier only where the tableau is first touched.  Used when the tile is large enough to pay for the
// barrier (QPN_TMA_MIN_BYTES: robust_avoid's levels 5.7 - 40 KB, the synthetic chain's 34 KB; four_player's 640 B keeps
// the 128-bit copy loop).
#ifndef QPN_TMA_MIN_BYTES
#define QPN_TMA_MIN_BYTES 2048
#endif
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile("{\n .reg .pred p;\n QPN_WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra QPN_DONE_%=;\n bra QPN_WAIT_%=;\n QPN_DONE_%=:\n}"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_inval(unsigned long long* bar) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Start of a solve from a plan: same state as tab_start + phase 0 + recompute_tcol + compact_dead.
// t.l() / t.u() hold the bounds; zb: n doubles of scratch.  Ends with a barrier.
__device__ __noinline__ void tab_start_plan_core(Tab t, const PlanDesc& P, const double* q, const double* z0, double* zb) {
    const int n = P.n, i = threadIdx.x;
    const int ldr = t.ldr;
    if (i < n) zb[i] = fmin(fmax(z0[i], t.l()[i]), t.u()[i]);
    const int nact = P.nact;                                      // rows nact .. n-1 are frozen: never copied (read from the plan at the end)
    const unsigned tbytes = (unsigned)(nact * ldr) * 8u;          // ldr is even: a multiple of 16
    unsigned long long* tbar = reinterpret_cast<unsigned long long*>(t.red_d() + 35);
    const bool aligned = (reinterpret_cast<uintptr_t>(P.T0) & 15) == 0;   // (T() is 16-byte aligned)
    const bool tma = aligned && tbytes >= QPN_TMA_MIN_BYTES;
    if (tma) {
        if (i == 0) {
            // every earlier access to the tableau buffer by this CTA lies before a barrier this thread has passed; the fence
            // orders them before the copy engine's writes
            mbar_init(tbar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(tbar, tbytes);
            tma_load_1d(t.T(), P.T0, tbytes, tbar);
        }
    } else if (aligned) {                                         // 128-bit copies
        const double2* src = reinterpret_cast<const double2*>(P.T0);
        double2* dst = reinterpret_cast<double2*>(t.T());
        #pragma unroll 1
        for (int e = i; e < (nact * ldr) >> 1; e += blockDim.x) dst[e] = src[e];
    } else {
        #pragma unroll 1
        for (int e = i; e < nact * ldr; e += blockDim.x) t.T()[e] = P.T0[e];
    }
    #pragma unroll 1
    for (int v = i; v <= 2 * n; v += blockDim.x) { t.rowof()[v] = -1; t.colof()[v] = -1; }
    if (i == 0) { *t.frozen_src() = nact < n ? P.T0 : nullptr; t.frozen_hdr()[0] = nact; t.frozen_hdr()[1] = P.tcol0; }
    QPN_SYNC();
    double zi = 0.0;
    if (i < n) {
        zi = z0[i];
        t.rr()[i] = ((csr_row_dot(P, i, zb) + q[i]) + zi) - zb[i];
        t.zst()[i] = (zi <= t.l()[i]) ? AT_L : (zi >= t.u()[i]) ? AT_U : FLOATING;
    }
    #pragma unroll 1
    for (int j = i; j < P.ncol0; j += blockDim.x) {
        const int v = P.colvar0[j];
        t.colvar()[j] = v; t.colof()[v] = j; t.nbval()[j] = v < n ? zb[v] : 0.0;
    }
    QPN_SYNC();
    double birv_i = 0.0;
    if (i < n) {
        {
            // The specification skips zero entries of B^-1; with a finite r (finite q, z0 clamped to finite bounds or finite
            // itself) fma(0, r_k, acc) == acc bit for bit (acc starts at +0.0 and can never become -0.0), so the branch-free
            // chain gives the same value and lets the loads run ahead of the dependent fma chain.
            double acc = 0.0;
            const double* pt = P.PT + i;
            const double* rr = t.rr();
            int k = 0;
            #pragma unroll 1
            for (; k + 4 <= n; k += 4) {
                const double p0 = pt[(size_t)k * n], p1 = pt[(size_t)(k + 1) * n], p2 = pt[(size_t)(k + 2) * n], p3 = pt[(size_t)(k + 3) * n];
                acc = fma(p0, rr[k], acc); acc = fma(p1, rr[k + 1], acc); acc = fma(p2, rr[k + 2], acc); acc = fma(p3, rr[k + 3], acc);
            }
            #pragma unroll 4
            for (; k < n; ++k) acc = fma(pt[(size_t)k * n], rr[k], acc);
            birv_i = acc;
            if (i >= nact) t.birv()[i] = acc;             // a frozen row is not in T(): frozen_values picks its entry up here
        }
        const int rv = P.rowvar0[i];
        t.rowvar()[i] = rv; t.rowof()[rv] = i;
        t.beta()[i] = rv < n ? zb[rv] : zb[rv - n] - z0[rv - n];      // (a plan may export its rows in another order: go by the variable)
        if (rv < n) t.zst()[rv] = BASIC;            // after the barrier that followed the default marks
    }
    if (tma) mbar_wait(tbar, 0);                      // the tableau has landed (phase 0 of a barrier initialised in this call)
    if (i < nact) t.T()[(size_t)i * ldr + P.tcol0] = birv_i;      // the homotopy column replaces the plan's placeholder
    QPN_SYNC();
    if (tma && i == 0) mbar_inval(tbar);              // the next solve of this CTA initialises it again
}
__device__ __forceinline__ void tab_start_plan(Tab& t, const PlanDesc& P, const double* q, const double* z0, double* zb) {
    tab_shape(t, P.n, P.ncol0);
    tab_start_plan_core(t, P, q, z0, zb);
    t.ncol = P.ncol0;
    t.pivots = P.npiv0;
    t.own_frozen = 0;
}

// Solve (M z + q) comp. l <= z <= u in the shared-memory tableau.  `build(t)` must fill
// T[i][0:n] with -M[i][:] and end with a barrier; it is called twice (start and final check).
// t.l() / t.u() hold the bounds, qs the vector q, zs the start on entry and z on exit.
// code (optional, smem, n entries) receives the basis codes.  Same procedure as the oracle's
// single-instance solve.
template <class Build>
__device__ __forceinline__ int solve_avi_smem(Tab& t, int n, Build build, const double* qs, double* zs,
                                     int max_pivots, int8_t* code_out, int* pivots_acc) {
    tab_shape(t, n, n + 1);
    build(t);
    tab_start(t, qs, zs);
    const PivotResult pr = avi_pivot_run(t, max_pivots, false);
    const double zi = pr.zi; const int8_t code = (int8_t)pr.code;
    int st = pr.st;
    *pivots_acc += pr.pivots;
    const int i = threadIdx.x;
    QPN_SYNC();
    if (i < n) { zs[i] = zi; if (code_out) code_out[i] = code; }
    build(t);
    int bad = 0;
    if (i < n) bad = check_avi_index(residual_row(t, zs, qs[i], i), zi, t.l()[i], t.u()[i], 1e-6);
    bad = QPN_SYNC_OR(bad);
    if (st == ST_SUCCESS && bad) st = ST_FAILURE;
    return st;
}

// The same solve started from a plan; the final check uses the plan's CSR rows.
__device__ __forceinline__ int solve_avi_plan(Tab& t, const PlanDesc& P, const double* qs, double* zs, double* zb,
                                     int max_pivots, int8_t* code_out, int* pivots_acc) {
    const int n = P.n, i = threadIdx.x;
    tab_start_plan(t, P, qs, zs, zb);
    const PivotResult pr = avi_pivot_run(t, max_pivots, true);
    const double zi = pr.zi; const int8_t code = (int8_t)pr.code;
    int st = pr.st;
    *pivots_acc += pr.pivots;
    QPN_SYNC();
    if (i < n) { zs[i] = zi; if (code_out) code_out[i] = code; }
    QPN_SYNC();
    int bad = 0;
    if (i < n) bad = check_avi_index(csr_row_dot(P, i, zs) + qs[i], zi, t.l()[i], t.u()[i], 1e-6);
    bad = QPN_SYNC_OR(bad);
    if (st == ST_SUCCESS && bad) st = ST_FAILURE;
    return st;
}

// ---- shared-memory plan of a GAVI solve ------------------------------------------------------
struct GaviSmem {
    Tab t;
    int base;             // byte offset of the extra arrays in qpn_smem
    int n, np, dz, d2;    // sizes they were carved for
    __device__ __forceinline__ double* dbl(int off) const { return reinterpret_cast<double*>(qpn_smem + base) + off; }
    __device__ __forceinline__ double* qs() const { return dbl(0); }              // n (n = d1 + 2 d2)
    __device__ __forceinline__ double* zs() const { return dbl(n); }              // n
    __device__ __forceinline__ double* zb() const { return dbl(2 * n); }          // n
    __device__ __forceinline__ double* w() const { return dbl(3 * n); }           // np
    __device__ __forceinline__ double* z0() const { return dbl(3 * n + np); }     // d1 + d2
    __device__ __forceinline__ double* c() const { return dbl(3 * n + np + dz); } // d2
    __device__ __forceinline__ double* s0() const { return dbl(3 * n + np + dz + d2); }   // d2
    __device__ __forceinline__ int* cols() const { return reinterpret_cast<int*>(dbl(3 * n + np + dz + 2 * d2)); }   // dz (+1: count)
    __device__ __forceinline__ int8_t* code() const {                              // n
        return reinterpret_cast<int8_t*>(reinterpret_cast<unsigned char*>(cols()) + (((dz + 2) * 4 + 15) / 16) * 16);
    }
};

__host__ __device__ __forceinline__ size_t gavi_extra_bytes(int d1, int d2, int np) {
    const size_t n = (size_t)d1 + 2 * d2, dz = (size_t)d1 + d2;
    size_t dbl = 3 * n + np + dz + 2 * (size_t)d2;
    size_t ints = dz + 2;
    return dbl * 8 + ((ints * 4 + 15) / 16) * 16 + ((n + 15) / 16) * 16;
}
// Workspace shape of a GAVI solve without plans (tableau n x (n+1)).
__host__ __device__ __forceinline__ size_t gavi_smem_bytes(int d1, int d2, int np) {
    const int n = d1 + 2 * d2;
    return tab_smem_bytes(n, n + 1) + gavi_extra_bytes(d1, d2, np);
}

// Returns the byte offset just past the extra arrays.
__device__ __forceinline__ int gavi_carve_extra(GaviSmem& s, const GaviDesc& g, int base_off) {
    s.base = base_off; s.n = g.d1 + 2 * g.d2; s.np = g.np; s.dz = g.d1 + g.d2; s.d2 = g.d2;
    return base_off + (int)gavi_extra_bytes(g.d1, g.d2, g.np);
}
__device__ __forceinline__ int gavi_carve(GaviSmem& s, const GaviDesc& g, int base_off) {
    const int n = g.d1 + 2 * g.d2;
    tab_carve(s.t, n, n + 1, base_off);
    return gavi_carve_extra(s, g, base_off + (int)tab_smem_bytes(n, n + 1));
}

// s0 = A z0 + B w (c = B w kept).  Ends with a barrier.
__device__ __forceinline__ void gavi_slack(GaviSmem& s, const GaviDesc& g, bool recompute_c) {
    const int i = threadIdx.x, dz = g.d1 + g.d2;
    #pragma unroll 1
    for (int r = i; r < g.d2; r += blockDim.x) {
        if (recompute_c) {
            double acc = 0.0;
            #pragma unroll 4
            for (int j = 0; j < g.np; ++j) acc = fma(g.B[(size_t)j * g.d2 + r], s.w()[j], acc);
            s.c()[r] = acc;
        }
        double acc = 0.0;
        #pragma unroll 4
        for (int j = 0; j < dz; ++j) acc = fma(g.A[(size_t)j * g.d2 + r], s.z0()[j], acc);
        s.s0()[r] = acc + s.c()[r];
    }
    QPN_SYNC();
}

// -M of the presolve AVI (avi.jl:79-99 as a lifted KKT system over the non-zero columns `cols`
// of A):  M = [I -A' 0; A 0 -I; 0 I 0] over [z(k); lambda(d2); s(d2)].  Ends with a barrier.
__device__ __forceinline__ void build_presolve(Tab& tt, const GaviDesc& g, const int* cols, int k) {
    const int ldr = tt.ldr, d2 = g.d2, pn = k + 2 * d2;
    if (threadIdx.x < pn) {
        double* row = tt.T() + (size_t)threadIdx.x * ldr;
        #pragma unroll 1
        for (int j = 0; j < pn; ++j) row[j] = 0.0;
    }
    QPN_SYNC();
    #pragma unroll 1
    for (int e = threadIdx.x; e < k * d2; e += blockDim.x) {
        const int a = e / d2, r = e - a * d2;
        const double v = g.A[(size_t)cols[a] * d2 + r];
        tt.T()[(size_t)a * ldr + (k + r)] = v;         // -(-A')   row a, column k+r
        tt.T()[(size_t)(k + r) * ldr + a] = -v;        // -(A)     row k+r, column a
    }
    #pragma unroll 1
    for (int e = threadIdx.x; e < k; e += blockDim.x) tt.T()[(size_t)e * ldr + e] = -1.0;
    #pragma unroll 1
    for (int e = threadIdx.x; e < d2; e += blockDim.x) {
        tt.T()[(size_t)(k + e) * ldr + (k + d2 + e)] = 1.0;      // -(-1)  row k+e, column k+d2+e
        tt.T()[(size_t)(k + d2 + e) * ldr + (k + e)] = -1.0;     // -(+1)  row k+d2+e, column k+e
    }
    QPN_SYNC();
}

// -M of the lifted AVI of `convert` (avi.jl:113-128): M = [M 0; A -I; 0 I 0].  Ends with a barrier.
__device__ __forceinline__ void build_lifted(Tab& tt, const GaviDesc& g) {
    const int ldr = tt.ldr, r = threadIdx.x, d1 = g.d1, d2 = g.d2, dz = d1 + d2, n = d1 + 2 * d2;
    if (r < n) {
        double* row = tt.T() + (size_t)r * ldr;
        if (r < d1) {
            #pragma unroll 4
            for (int j = 0; j < dz; ++j) row[j] = -g.M[(size_t)j * d1 + r];
            #pragma unroll 1
            for (int j = dz; j < n; ++j) row[j] = 0.0;
        } else if (r < dz) {
            #pragma unroll 4
            for (int j = 0; j < dz; ++j) row[j] = -g.A[(size_t)j * d2 + (r - d1)];
            #pragma unroll 1
            for (int j = dz; j < n; ++j) row[j] = (j - dz == r - d1) ? 1.0 : 0.0;
        } else {
            #pragma unroll 1
            for (int j = 0; j < n; ++j) row[j] = (j - d1 == r - dz) ? -1.0 : 0.0;
        }
    }
    QPN_SYNC();
}

// Non-zero columns of A into cols[0..k), k into cols[dz].  Ends with a barrier.
__device__ __forceinline__ void find_cols(const GaviDesc& g, int* cols) {
    const int dz = g.d1 + g.d2, d2 = g.d2, i = threadIdx.x;
    #pragma unroll 1
    for (int j = i; j < dz; j += blockDim.x) {
        bool nz = false;
        #pragma unroll 4
        for (int r = 0; r < d2; ++r) nz |= (g.A[(size_t)j * d2 + r] != 0.0);
        cols[j] = nz ? 1 : 0;
    }
    QPN_SYNC();
    if (i == 0) {
        int k = 0;
        #pragma unroll 1
        for (int j = 0; j < dz; ++j) if (cols[j]) cols[k++] = j;
        cols[dz] = k;
    }
    QPN_SYNC();
}

// solve_gavi (avi.jl:101-111) for the instance whose w and z0 are already in s.w() / s.z0().
// planA (lifted AVI) / planB (presolve AVI) may be null: the instance then runs phase 0 itself.
// On return s.zs() holds the lifted solution [z1; z2; s] and s.code() the basis codes.
__device__ __forceinline__ int gavi_solve_smem(GaviSmem& s, const GaviDesc& g, const PlanDesc* planA, const PlanDesc* planB,
                                      int presolve, int max_pivots, int* pivots) {
    const int d1 = g.d1, d2 = g.d2, dz = d1 + d2, n = d1 + 2 * d2, i = threadIdx.x;
    Tab& t = s.t;
    gavi_slack(s, g, true);
    if (presolve && d2 > 0) {
        int infeasible = 0;
        #pragma unroll 1
        for (int r = i; r < d2; r += blockDim.x)
            if (!(g.l2[r] <= s.s0()[r] && s.s0()[r] <= g.u2[r])) infeasible = 1;
        infeasible = QPN_SYNC_OR(infeasible);
        if (infeasible) {
            // find_closest_feasible! (avi.jl:79-99): min |z - z0|^2 s.t. l2 - Bw <= A z <= u2 - Bw over
            // the columns of A that are not structurally zero, as the lifted KKT AVI.
            const int* cols;
            int k;
            if (planB) { cols = planB->cols; k = planB->ncols; }
            else { find_cols(g, s.cols()); cols = s.cols(); k = s.cols()[dz]; }
            const int pn = k + 2 * d2;
            if (i < pn) {
                if (i < k) { s.qs()[i] = -s.z0()[cols[i]]; s.zs()[i] = s.z0()[cols[i]]; t.l()[i] = -QPN_INF; t.u()[i] = QPN_INF; }
                else if (i < k + d2) {
                    const int r = i - k;
                    double full = 0.0, part = 0.0;
                    #pragma unroll 4
                    for (int j = 0; j < dz; ++j) full = fma(g.A[(size_t)j * d2 + r], s.z0()[j], full);
                    #pragma unroll 4
                    for (int a = 0; a < k; ++a) part = fma(g.A[(size_t)cols[a] * d2 + r], s.z0()[cols[a]], part);
                    s.qs()[i] = (full - part) + s.c()[r];
                    s.zs()[i] = 0.0; t.l()[i] = -QPN_INF; t.u()[i] = QPN_INF;
                } else {
                    const int r = i - k - d2;
                    s.qs()[i] = 0.0; s.zs()[i] = s.s0()[r]; t.l()[i] = g.l2[r]; t.u()[i] = g.u2[r];
                }
            }
            QPN_SYNC();
            int pst;
            if (planB) pst = solve_avi_plan(t, *planB, s.qs(), s.zs(), s.zb(), 50 * pn + 100, nullptr, pivots);
            else pst = solve_avi_smem(t, pn, [&](Tab& tt) { build_presolve(tt, g, cols, k); }, s.qs(), s.zs(), 50 * pn + 100, nullptr, pivots);
            QPN_SYNC();
            if (pst == ST_SUCCESS && i < k) s.z0()[cols[i]] = s.zs()[i];
            QPN_SYNC();
            gavi_slack(s, g, false);
        }
    }
    // convert (avi.jl:113-128): lifted AVI over [z1; z2; s]
    if (i < n) {
        if (i < d1) {
            double acc = 0.0;
            #pragma unroll 4
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
    if (planA) return solve_avi_plan(t, *planA, s.qs(), s.zs(), s.zb(), max_pivots, s.code(), pivots);
    return solve_avi_smem(t, n, [&](Tab& tt) { build_lifted(tt, g); }, s.qs(), s.zs(), max_pivots, s.code(), pivots);
}

// Plans of a GAVI's two AVIs and the workspace shape they allow.
struct GaviPlans {
    int has;
    PlanDesc A, B;          // lifted AVI / presolve AVI
    int t_doubles, ldr_max;
};
__host__ inline void gavi_workspace_shape(const GaviDesc& g, GaviPlans& pl, int extra_rows = 0, int extra_cap = 0) {
    const int n = g.d1 + 2 * g.d2;
    int ldr = 2; size_t td = 2;
    auto take = [&](int rows, int cap) {
        const int l = row_stride(cap);
        if (l > ldr) ldr = l;
        if ((size_t)rows * l > td) td = (size_t)rows * l;
    };
    if (pl.has) { take(pl.A.nact, pl.A.ncol0); take(pl.B.nact, pl.B.ncol0); }      // swept rows only
    else take(n, n + 1);
    if (extra_rows > 0) take(extra_rows, extra_cap);
    pl.t_doubles = (int)td; pl.ldr_max = ldr;
}

template <int MAXT>
__global__ void __launch_bounds__(MAXT, 896 / MAXT) gavi_solve_kernel(const __grid_constant__ GaviDesc g, const __grid_constant__ GaviPlans plans, int batch, const double* __restrict__ w,
                                  const double* __restrict__ z0, int presolve, int max_pivots,
                                  double* __restrict__ z_out, double* __restrict__ zfull_out,
                                  int32_t* __restrict__ status_out, int32_t* __restrict__ pivots_out,
                                  int8_t* __restrict__ basis_out) {
    const int b = blockIdx.x, i = threadIdx.x;
    const int dz = g.d1 + g.d2, n = g.d1 + 2 * g.d2;
    GaviSmem s;
    tab_carve_ex(s.t, n, (size_t)plans.t_doubles, plans.ldr_max, 0);
    gavi_carve_extra(s, g, (int)tab_smem_bytes_ex(n, (size_t)plans.t_doubles, plans.ldr_max));
    for (int j = i; j < g.np; j += blockDim.x) s.w()[j] = w[(size_t)b * g.np + j];
    for (int j = i; j < dz; j += blockDim.x) s.z0()[j] = z0[(size_t)b * dz + j];
    QPN_SYNC();
    int piv = 0;
    const int st = gavi_solve_smem(s, g, plans.has ? &plans.A : nullptr, plans.has ? &plans.B : nullptr, presolve, max_pivots, &piv);
    QPN_SYNC();
    if (i < dz) z_out[(size_t)b * dz + i] = s.zs()[i];
    if (i < n) {
        if (zfull_out) zfull_out[(size_t)b * n + i] = s.zs()[i];
        if (basis_out) basis_out[(size_t)b * n + i] = s.code()[i];
    }
    if (i == 0) { status_out[b] = st; pivots_out[b] = piv; }
}

// Common tail of plan construction: T holds -M (n x n) and t.l / t.u the bounds; qs = zs = 0.
// Exports the CSR rows, runs phase 0 exactly as crash() does, exports B^-1, compacts, exports T0.
__device__ __forceinline__ void plan_finish(Tab& t, int n, int k, const double* qs, const double* zs, int* cnt,
                                            double* __restrict__ T0, double* __restrict__ PT, int* __restrict__ rowvar0,
                                            int* __restrict__ colvar0, int* __restrict__ csr_ptr, int* __restrict__ csr_col,
                                            double* __restrict__ csr_val, int* __restrict__ hdr) {
    const int i = threadIdx.x;
    // the original matrix in CSR (rows ascending in the column index)
    if (i < n) {
        const double* row = t.T() + (size_t)i * t.ldr;
        int c = 0;
        for (int j = 0; j < n; ++j) c += (row[j] != 0.0);
        cnt[i] = c;
    }
    QPN_SYNC();
    if (i == 0) {
        int acc = 0;
        for (int r = 0; r < n; ++r) { csr_ptr[r] = acc; acc += cnt[r]; }
        csr_ptr[n] = acc;
    }
    QPN_SYNC();
    if (i < n) {
        const double* row = t.T() + (size_t)i * t.ldr;
        int o = csr_ptr[i];
        for (int j = 0; j < n; ++j) if (row[j] != 0.0) { csr_col[o] = j; csr_val[o] = -row[j]; ++o; }
    }
    QPN_SYNC();
    tab_start(t, qs, zs);
    // phase 0 exactly as crash() runs it
    for (int v = 0; v < n; ++v) {
        if (!is_free_var(t, v)) continue;
        const int c = t.colof()[v];
        const int rho = best_free_row(t, c);
        if (rho >= 0) { pivot(t, rho, c, false); set_zst(t, v, BASIC); }
    }
    // Export order of the rows: the rows an instance sweeps first, the rows of free basics (frozen from here on,
    // avi_pivot.cuh: freeze) last, each group in its original order -- every tie rule only compares swept rows, so
    // their relative order is all that matters.  An instance then holds rows 0 .. nact-1 only and reads the frozen
    // ones from the plan, once, at the end.
    int* perm = cnt;                                       // the CSR counts are done with
    QPN_SYNC();
    if (i == 0) {
        int na = 0;
        for (int r = 0; r < n; ++r) if (t.rowvar()[r] >= n) perm[r] = na++;
        hdr[4] = na;
        for (int r = 0; r < n; ++r) if (t.rowvar()[r] < n) perm[r] = na++;
    }
    QPN_SYNC();
    // B^-1 from the slack columns (see recompute_tcol)
    if (i < n) {
        const double* row = t.T() + (size_t)i * t.ldr;
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
    if (i < n) {
        const double* row = t.T() + (size_t)i * t.ldr;
        const int pi = perm[i];
        for (int j = 0; j < ldr0; ++j) T0[(size_t)pi * ldr0 + j] = j < ncol0 ? row[j] : 0.0;
        rowvar0[pi] = t.rowvar()[i];
    }
    for (int j = i; j < ncol0; j += blockDim.x) colvar0[j] = t.colvar()[j];
    if (i == 0) { hdr[0] = ncol0; hdr[1] = npiv0; hdr[2] = t.colof()[2 * n]; hdr[3] = k; }
}

// ---- plan construction: one CTA, once per shared matrix ---------------------------------------
// kind 0: lifted AVI of the GAVI; kind 1: its presolve AVI.  Buffers are sized by the host for the
// worst case (T0: n x row_stride(n+1), CSR: n*n entries).  hdr: [ncol0, npiv0, tcol0, ncols].
__global__ void plan_build_kernel(const __grid_constant__ GaviDesc g, int kind, double* __restrict__ T0, double* __restrict__ PT,
                                  int* __restrict__ rowvar0, int* __restrict__ colvar0, int* __restrict__ csr_ptr,
                                  int* __restrict__ csr_col, double* __restrict__ csr_val, int* __restrict__ cols_out,
                                  int* __restrict__ hdr) {
    const int i = threadIdx.x, d1 = g.d1, d2 = g.d2, dz = d1 + d2;
    GaviSmem s;
    gavi_carve(s, g, 0);
    Tab& t = s.t;
    int n, k = 0;
    if (kind == 1) {
        find_cols(g, s.cols());
        k = s.cols()[dz];
        n = k + 2 * d2;
        for (int j = i; j < k; j += blockDim.x) cols_out[j] = s.cols()[j];
        if (i < n) {
            const bool fr = i < k + d2;
            t.l()[i] = fr ? -QPN_INF : g.l2[i - k - d2];
            t.u()[i] = fr ? QPN_INF : g.u2[i - k - d2];
        }
    } else {
        n = d1 + 2 * d2;
        if (i < n) {
            t.l()[i] = i < d1 ? g.l1[i] : i < dz ? -QPN_INF : g.l2[i - dz];
            t.u()[i] = i < d1 ? g.u1[i] : i < dz ? QPN_INF : g.u2[i - dz];
        }
    }
    if (i < n) { s.qs()[i] = 0.0; s.zs()[i] = 0.0; }
    QPN_SYNC();
    tab_shape(t, n, n + 1);
    if (kind == 1) build_presolve(t, g, s.cols(), k); else build_lifted(t, g);
    plan_finish(t, n, k, s.qs(), s.zs(), reinterpret_cast<int*>(s.zb()), T0, PT, rowvar0, colvar0, csr_ptr, csr_col, csr_val, hdr);
}

// ---- verify_solution (qp_processing.jl:57-149) ---------------------------------------------------
struct NodeDesc {
    int nd, nv, m;
    const double *Qd, *qd, *A, *l, *u;
    const int32_t* dec;
};

struct VerifySmem {
    int base, nd, m;      // byte offset in qpn_smem and the sizes it was carved for
    int ab;               // byte offset of the two nd x m least-squares matrices: their own region, or -- in the level kernel --
                          // the tableau buffer, which they time-share (they are dead before the fallback builds its tableau,
                          // and no solve is in flight while a node is verified)
    __device__ __forceinline__ double* dbl(int off) const { return reinterpret_cast<double*>(qpn_smem + base) + off; }
    __device__ __forceinline__ double* Ab() const { return reinterpret_cast<double*>(qpn_smem + ab); }   // nd x m
    __device__ __forceinline__ double* Ab0() const { return Ab() + nd * m; }                  // nd x m
    __device__ __forceinline__ double* b() const { return dbl(0); }                           // nd
    __device__ __forceinline__ double* lam() const { return dbl(nd); }                        // m
    __device__ __forceinline__ double* v() const { return dbl(nd + m); }                      // nd + m
    __device__ __forceinline__ double* lam_out() const { return dbl(2 * nd + 2 * m); }        // m
    __device__ __forceinline__ double* qs() const { return dbl(2 * nd + 3 * m); }             // m (fallback AVI)
    __device__ __forceinline__ double* zs() const { return dbl(2 * nd + 4 * m); }             // m
    __device__ __forceinline__ int* idx() const { return reinterpret_cast<int*>(dbl(2 * nd + 5 * m)); }   // m
    __device__ __forceinline__ int* perm() const { return idx() + m; }                        // m
    __device__ __forceinline__ int8_t* kind() const {                                         // m
        return reinterpret_cast<int8_t*>(reinterpret_cast<unsigned char*>(idx()) + ((2 * m * 4 + 15) / 16) * 16);
    }
};

// Bytes of the vectors (everything but the two matrices) ...
__host__ __device__ __forceinline__ size_t verify_rest_bytes(int nd, int m) {
    size_t dbl = 2 * (size_t)nd + 5 * (size_t)m;
    size_t ints = 2 * (size_t)m;
    return ((dbl * 8 + 15) / 16) * 16 + ((ints * 4 + 15) / 16) * 16 + (((size_t)m + 15) / 16) * 16;
}
// ... and with the matrices in a region of their own right behind them.
__host__ __device__ __forceinline__ size_t verify_smem_bytes(int nd, int m) { return verify_rest_bytes(nd, m) + 16 * (size_t)nd * m; }

// ab_off < 0: the matrices follow the vectors; else they live at byte offset ab_off (time-shared tableau buffer).
__device__ __forceinline__ void verify_carve(VerifySmem& v, int nd, int m, int base_off, int ab_off = -1) {
    v.base = base_off; v.nd = nd; v.m = m;
    v.ab = ab_off >= 0 ? ab_off : base_off + (int)verify_rest_bytes(nd, m);
}

// Householder QR least squares with column pivoting, one thread per column; reductions run
// down a column sequentially so the bits match oracle/qpn_oracle.c:lstsq_basic.
// `red` gives the block reduction scratch (Tab with red_d / red_i).
__device__ __forceinline__ void lstsq_basic_block(const Tab& red, int nd, int k, double* Ab, double* b, double* lam,
                                         int* perm, double* v) {
    const int j = threadIdx.x;
    const int steps = nd < k ? nd : k;
    #pragma unroll 1
    for (int c = j; c < k; c += blockDim.x) perm[c] = c;
    QPN_SYNC();
    int rank = 0;
    #pragma unroll 1
    for (int c = 0; c < steps; ++c) {
        double best = -1.0; int jb = -1;
        // candidates j in [c, k): handled in strides so k may exceed blockDim
        #pragma unroll 1
        for (int jj = j; jj < k; jj += blockDim.x) {
            if (jj < c) continue;
            double s = 0.0;
            #pragma unroll 4
            for (int i = c; i < nd; ++i) s = fma(Ab[(size_t)jj * nd + i], Ab[(size_t)jj * nd + i], s);
            if (jb < 0 || s > best) { best = s; jb = jj; }
        }
        // squared norms are >= 0; with more columns than threads a thread's own best is already
        // the lowest-index maximum of its strided set, so ties across threads keep the lower thread
        block_argmax_idx(red, jb >= 0, best, jb);
        const double nrm = sqrt(best);
        if (nrm <= 1e-10) break;
        QPN_SYNC();
        if (jb != c) {
            #pragma unroll 1
            for (int i = j; i < nd; i += blockDim.x) {
                const double tmp = Ab[(size_t)c * nd + i];
                Ab[(size_t)c * nd + i] = Ab[(size_t)jb * nd + i];
                Ab[(size_t)jb * nd + i] = tmp;
            }
            if (j == 0) { const int tp = perm[c]; perm[c] = perm[jb]; perm[jb] = tp; }
        }
        QPN_SYNC();
        const double alpha = Ab[(size_t)c * nd + c] > 0.0 ? -nrm : nrm;
        #pragma unroll 1
        for (int i = c + j; i < nd; i += blockDim.x) v[i] = Ab[(size_t)c * nd + i] - (i == c ? alpha : 0.0);
        QPN_SYNC();
        double vn = 0.0;
        #pragma unroll 4
        for (int i = c; i < nd; ++i) vn = fma(v[i], v[i], vn);
        if (vn > 0.0) {
            #pragma unroll 1
            for (int jj = j; jj < k + 1; jj += blockDim.x) {
                if (jj < c) continue;
                double* col = (jj < k) ? Ab + (size_t)jj * nd : b;     // the rhs rides as column k
                double s = 0.0;
                #pragma unroll 4
                for (int i = c; i < nd; ++i) s = fma(v[i], col[i], s);
                s = (2.0 * s) / vn;
                #pragma unroll 4
                for (int i = c; i < nd; ++i) col[i] = fma(-s, v[i], col[i]);
            }
        }
        rank++;
        QPN_SYNC();
    }
    QPN_SYNC();
    if (j == 0) {
        #pragma unroll 1
        for (int t = 0; t < k; ++t) lam[t] = 0.0;
        #pragma unroll 1
        for (int i = rank - 1; i >= 0; --i) {
            double acc = b[i];
            #pragma unroll 4
            for (int t = i + 1; t < rank; ++t) acc = fma(-Ab[(size_t)t * nd + i], v[t], acc);
            v[i] = acc / Ab[(size_t)i * nd + i];
        }
        #pragma unroll 1
        for (int i = 0; i < rank; ++i) lam[perm[i]] = v[i];
    }
    QPN_SYNC();
}

// Gradient and constraint values of one node:  qt = Q[dec,:] x + q[dec],  ax = A x.
// Thread r takes row r; each dot product is sequential in the variable index.
__device__ __forceinline__ void node_products(const NodeDesc& nd_, const double* x, double* qt, double* ax) {
    const int nd = nd_.nd, nv = nd_.nv, m = nd_.m;
    #pragma unroll 1
    for (int r = threadIdx.x; r < nd + m; r += blockDim.x) {
        double acc = 0.0;
        if (r < nd) {
            #pragma unroll 4
            for (int j = 0; j < nv; ++j) acc = fma(nd_.Qd[(size_t)j * nd + r], x[j], acc);
            qt[r] = acc + nd_.qd[r];
        } else {
            const int rr = r - nd;
            #pragma unroll 4
            for (int j = 0; j < nv; ++j) acc = fma(nd_.A[(size_t)j * m + rr], x[j], acc);
            ax[rr] = acc;
        }
    }
}

// Returns 1 when the point whose products qt / ax are given is a solution for the node; lam in
// vs.lam_out().  tab: a tableau workspace large enough for an AVI of size m (used by the fallback
// and for the reduction scratch).  qt / ax must be complete (barrier) on entry.
// WIDE (compile time; the level kernel's instantiations for more than 64 threads): the acceptance test reads the signed
// entries of A again, one thread per component, instead of keeping a second copy of the matrix (Ab0) in shared memory
// and checking on one thread -- at nd = m = 64 that copy is 32 KB and the serial check 4,096 dependent steps.  Same
// arithmetic, same order.  The small instantiations keep the original form (their code is what four_player runs).
template <bool WIDE>
__device__ __forceinline__ int verify_solution_smem(Tab& tab, VerifySmem& vs, const NodeDesc& nd_, const double* qt, const double* ax,
                                           double tol, int* how, int* pivots) {
    const int nd = nd_.nd, m = nd_.m, i = threadIdx.x;
    int infeasible = 0;
    #pragma unroll 1
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
    #pragma unroll 1
    for (int r = 0; r < nd; ++r) nq = fma(qt[r], qt[r], nq);
    if (m == 0) { *how = 1; return sqrt(nq) <= tol ? 1 : 0; }
    // no active row at all (the common case at an interior point): skip the serial ordering below
    int any_active = 0;
    #pragma unroll 1
    for (int r = i; r < m; r += blockDim.x) any_active |= (vs.kind()[r] != 0);
    any_active = QPN_SYNC_OR(any_active);
    // order the active rows: lower-active, upper-active, both (qp_processing.jl:105-114)
    if (!any_active) {
        if (i == 0) { tab.red_i()[32] = 0; tab.red_i()[33] = 0; tab.red_i()[34] = 0; }
    } else if (i == 0) {
        int k = 0, np_ = 0, nn = 0;
        #pragma unroll 1
        for (int r = 0; r < m; ++r) if (vs.kind()[r] == 1) { vs.idx()[k++] = r; np_++; }
        #pragma unroll 1
        for (int r = 0; r < m; ++r) if (vs.kind()[r] == 2) { vs.idx()[k++] = r; nn++; }
        #pragma unroll 1
        for (int r = 0; r < m; ++r) if (vs.kind()[r] == 3) { vs.idx()[k++] = r; }
        tab.red_i()[32] = k; tab.red_i()[33] = np_; tab.red_i()[34] = nn;
    }
    QPN_SYNC();
    const int k = tab.red_i()[32], np_ = tab.red_i()[33], nn = tab.red_i()[34];
    if (k == 0) {
        // No active row: lam = 0 and the residual is qt itself.  The least-squares step rejects
        // (else we would have returned above only for m == 0) unless |qt| <= tol; the fallback has
        // every multiplier fixed at 0, so it returns lam = 0 after 0 pivots and rejects as well.
        QPN_SYNC();
        if (sqrt(nq) <= tol) { *how = 2; return 1; }
        *how = 4;
        return 0;
    }
    #pragma unroll 1
    for (int e = i; e < nd * k; e += blockDim.x) {
        const int tcol = e / nd, r = e - tcol * nd;
        const double sgn = (tcol >= np_ && tcol < np_ + nn) ? -1.0 : 1.0;
        const double val = sgn * nd_.A[(size_t)nd_.dec[r] * m + vs.idx()[tcol]];
        vs.Ab()[e] = val;
        if (!WIDE) vs.Ab0()[e] = val;
    }
    #pragma unroll 1
    for (int r = i; r < nd; r += blockDim.x) vs.b()[r] = qt[r];
    QPN_SYNC();
    lstsq_basic_block(tab, nd, k, vs.Ab(), vs.b(), vs.lam(), vs.perm(), vs.v());
    // acceptance (qp_processing.jl:119)
    if (WIDE) {
        #pragma unroll 1
        for (int r = i; r < nd; r += blockDim.x) {
            double acc = 0.0;
            #pragma unroll 1
            for (int t = 0; t < k; ++t) {
                const double a = nd_.A[(size_t)nd_.dec[r] * m + vs.idx()[t]];
                acc = fma((t >= np_ && t < np_ + nn) ? -a : a, vs.lam()[t], acc);
            }
            vs.v()[r] = acc - qt[r];                      // v: the factorisation's scratch, free again
        }
        QPN_SYNC();
    }
    if (i == 0) {
        int ok = 1;
        #pragma unroll 1
        for (int t = 0; t < np_ + nn; ++t) if (!(vs.lam()[t] > -tol)) ok = 0;
        double res = 0.0;
        #pragma unroll 1
        for (int r = 0; r < nd; ++r) {
            double e;
            if (WIDE) e = vs.v()[r];
            else {
                double acc = 0.0;
                #pragma unroll 1
                for (int t = 0; t < k; ++t) acc = fma(vs.Ab0()[(size_t)t * nd + r], vs.lam()[t], acc);
                e = acc - qt[r];
            }
            res = fma(e, e, res);
        }
        if (!(sqrt(res) <= tol)) ok = 0;
        tab.red_i()[35] = ok;
    }
    QPN_SYNC();
    if (tab.red_i()[35]) {
        #pragma unroll 1
        for (int t = i; t < k; t += blockDim.x)
            vs.lam_out()[vs.idx()[t]] = (t >= np_ && t < np_ + nn) ? -vs.lam()[t] : vs.lam()[t];
        QPN_SYNC();
        *how = 2;
        return 1;
    }
    // fallback (qp_processing.jl:129-146): sign-constrained least squares as the box AVI
    //   (Ad Ad') lam - Ad qt  comp.  lb <= lam <= ub
    QPN_SYNC();
    #pragma unroll 1
    for (int r = i; r < m; r += blockDim.x) {
        double acc = 0.0;
        #pragma unroll 4
        for (int t = 0; t < nd; ++t) acc = fma(nd_.A[(size_t)nd_.dec[t] * m + r], qt[t], acc);
        vs.qs()[r] = -acc;
        vs.zs()[r] = 0.0;
        const int8_t kd = vs.kind()[r];
        tab.l()[r] = (kd == 2 || kd == 3) ? -QPN_INF : 0.0;
        tab.u()[r] = (kd == 1 || kd == 3) ? QPN_INF : 0.0;
    }
    QPN_SYNC();
    auto build = [&](Tab& tt) {
        const int ldr = tt.ldr;
        #pragma unroll 1
        for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
            const int c = e / m, r = e - c * m;
            double acc = 0.0;
            #pragma unroll 4
            for (int t = 0; t < nd; ++t) acc = fma(nd_.A[(size_t)nd_.dec[t] * m + r], nd_.A[(size_t)nd_.dec[t] * m + c], acc);
            tt.T()[(size_t)r * ldr + c] = -acc;
        }
        QPN_SYNC();
    };
    const int n_keep = tab.n;
    const int st = solve_avi_smem(tab, m, build, vs.qs(), vs.zs(), 50 * m + 100, nullptr, pivots);
    tab.n = n_keep;
    QPN_SYNC();
    if (st != ST_SUCCESS) { *how = 5; return 0; }
    if (i == 0) {
        double res2 = 0.0;
        #pragma unroll 1
        for (int t = 0; t < nd; ++t) {
            double acc = 0.0;
            #pragma unroll 4
            for (int r = 0; r < m; ++r) acc = fma(nd_.A[(size_t)nd_.dec[t] * m + r], vs.zs()[r], acc);
            const double e = acc - qt[t];
            res2 = fma(e, e, res2);
        }
        tab.red_i()[35] = sqrt(res2) <= 1e-4 ? 1 : 0;
    }
    #pragma unroll 1
    for (int r = i; r < m; r += blockDim.x) vs.lam_out()[r] = vs.zs()[r];
    QPN_SYNC();
    const int ok2 = tab.red_i()[35];
    *how = ok2 ? 3 : 4;
    return ok2;
}

// grid = batch, block = roundup32(max(m, nd, 1)).  Dynamic smem: Tab(m, m+1) + VerifySmem + x(nv) + qt(nd) + ax(m).
__global__ void verify_solution_kernel(const __grid_constant__ NodeDesc node, int batch, const double* __restrict__ x, double tol,
                                       uint8_t* __restrict__ solution_out, double* __restrict__ lam_out,
                                       int32_t* __restrict__ how_out, int8_t* __restrict__ active_out) {
    const int b = blockIdx.x, i = threadIdx.x, m = node.m;
    const int tn = m > 0 ? m : 1;
    Tab tab;
    tab_carve(tab, tn, tn + 1, 0);
    VerifySmem vs;
    const int p = (int)tab_smem_bytes(tn, tn + 1);
    verify_carve(vs, node.nd, m, p);
    double* xs = reinterpret_cast<double*>(qpn_smem + p + verify_smem_bytes(node.nd, m));
    double* qt = xs + node.nv;
    double* ax = qt + node.nd;
    for (int j = i; j < node.nv; j += blockDim.x) xs[j] = x[(size_t)b * node.nv + j];
    QPN_SYNC();
    node_products(node, xs, qt, ax);
    QPN_SYNC();
    int how = 0, piv = 0;
    const int sol = verify_solution_smem<false>(tab, vs, node, qt, ax, tol, &how, &piv);
    QPN_SYNC();
    for (int r = i; r < m; r += blockDim.x) {
        if (lam_out) lam_out[(size_t)b * m + r] = vs.lam_out()[r];
        if (active_out) active_out[(size_t)b * m + r] = how == 0 ? 0 : vs.kind()[r];
    }
    if (i == 0) { solution_out[b] = (uint8_t)sol; if (how_out) how_out[b] = how; }
}

// ---- fused per-level equilibrium loop (algorithm.jl:13-118, level without children) --------------
constexpr int QPN_MAX_PLAYERS = 8;

struct LevelDesc {
    int nv, nplayers;
    NodeDesc players[QPN_MAX_PLAYERS];
    int nd_off[QPN_MAX_PLAYERS], m_off[QPN_MAX_PLAYERS];   // offsets into the stacked qt / ax (= lam) arrays
    GaviDesc g;
    const int32_t* dec;     // nd_level
    const int32_t* par;     // g.np
    int nd_level;
    int max_iters;
    int nproj;
    const double* proj;     // nv x nproj
    int max_nd, max_m;      // over players
    int nd_total;           // sum of nd over players
    int lam_total;          // sum of m over players
    // plans (resident levels only) and the workspace shape they allow
    int has_plans;
    PlanDesc planA, planB;  // lifted AVI / presolve AVI
    int t_doubles, ldr_max; // tableau buffer (doubles) and longest row of any solve in the kernel
};

// Fills t_doubles / ldr_max: the largest (rows x row stride) among the solves the level kernel runs.
__host__ inline void level_workspace_shape(LevelDesc& lv) {
    const int n = lv.g.d1 + 2 * lv.g.d2;
    int ldr = row_stride(lv.max_m + 1);
    size_t td = (size_t)lv.max_m * ldr;                                   // verify_solution's fallback
    {   // ... and its least-squares matrices (VerifySmem::ab): two, or one in the instantiations for more than 64 threads
        int need = n > lv.max_m ? n : lv.max_m;
        if (lv.max_nd > need) need = lv.max_nd;
        const size_t mats = (need > 64 ? 1 : 2) * (size_t)lv.max_nd * lv.max_m;      // need > 64 <=> level_equilibrium_kernel<128 / 256>
        if (mats > td) td = mats;
    }
    auto take = [&](int rows, int cap) {
        const int l = row_stride(cap);
        if (l > ldr) ldr = l;
        if ((size_t)rows * l > td) td = (size_t)rows * l;
    };
    if (lv.has_plans) { take(lv.planA.nact, lv.planA.ncol0); take(lv.planB.nact, lv.planB.ncol0); }      // swept rows only
    else take(n, n + 1);
    lv.t_doubles = (int)td; lv.ldr_max = ldr;
}

__host__ __device__ __forceinline__ size_t level_smem_bytes(const LevelDesc& lv) {
    return tab_smem_bytes_ex(lv.g.d1 + 2 * lv.g.d2, (size_t)lv.t_doubles, lv.ldr_max) + gavi_extra_bytes(lv.g.d1, lv.g.d2, lv.g.np) +
           verify_rest_bytes(lv.max_nd, lv.max_m) +
           8 * (2 * (size_t)lv.nv + (size_t)(lv.nproj > 0 ? lv.nproj : 1) + (size_t)lv.nd_total + (size_t)lv.lam_total);
}

// hist: global scratch, batch x hist_cap x nproj (cycle check history); hist_count: batch.
template <int MAXT>
__global__ void __launch_bounds__(MAXT, 896 / MAXT) level_equilibrium_kernel(const __grid_constant__ LevelDesc lv, int batch, const double* __restrict__ x_init,
                                         double* __restrict__ x_out, uint8_t* __restrict__ solved_out,
                                         int32_t* __restrict__ iters_out, int32_t* __restrict__ pivots_out,
                                         double* __restrict__ lam_out, double* __restrict__ hist,
                                         int32_t* __restrict__ hist_count, int hist_cap, int presolve, int hist_fresh) {
    const int b = blockIdx.x, i = threadIdx.x, nv = lv.nv;
    GaviSmem gs;
    const int n_level = lv.g.d1 + 2 * lv.g.d2;
    tab_carve_ex(gs.t, n_level, (size_t)lv.t_doubles, lv.ldr_max, 0);
    const int p = gavi_carve_extra(gs, lv.g, (int)tab_smem_bytes_ex(n_level, (size_t)lv.t_doubles, lv.ldr_max));
    VerifySmem vs;
    verify_carve(vs, lv.max_nd, lv.max_m, p, 0);           // the matrices time-share the tableau buffer (byte offset 0)
    double* xs = reinterpret_cast<double*>(qpn_smem + p + verify_rest_bytes(lv.max_nd, lv.max_m));
    double* pv = xs + nv;                  // nproj
    double* xn = pv + (lv.nproj > 0 ? lv.nproj : 1);
    double* qt_all = xn + nv;              // nd_total
    double* ax_all = qt_all + lv.nd_total; // lam_total
    for (int j = i; j < nv; j += blockDim.x) xs[j] = x_init[(size_t)b * nv + j];
    QPN_SYNC();
    const int max_piv = 50 * n_level + 100;
    const int rows_all = lv.nd_total + lv.lam_total;
    int solved = 0, piv = 0, iters = 0;
    // The history outlives one call when the caller keeps hist / hist_count (the reference's
    // iterate_cache persists across the calls of one top-level solve, algorithm.jl:20-28).
    double* myhist = hist ? hist + (size_t)b * hist_cap * lv.nproj : nullptr;
    int nhist = (hist && hist_count && !hist_fresh) ? hist_count[b] : 0;      // hist_fresh: this call starts a new top-level solve
    for (int it = 1; it <= lv.max_iters; ++it) {
        iters = it;
        // cycle check (algorithm.jl:14-30): random projections of the iterate against the history
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
                // isapprox with the default rtol = sqrt(eps)
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
        // process_qp for every player at the level (algorithm.jl:47-49): first every player's
        // gradient and constraint values in one pass over the stacked rows ...
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
        // ... then the per-player tests
        int all_sol = 1;
        for (int pl = 0; pl < lv.nplayers; ++pl) {
            const NodeDesc& node = lv.players[pl];
            int how = 0;
            gs.t.n = n_level;
            const int sol = verify_solution_smem<(MAXT > 64)>(gs.t, vs, node, qt_all + lv.nd_off[pl], ax_all + lv.m_off[pl], 1e-4, &how, &piv);
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
        gs.t.n = n_level;
        const int st = gavi_solve_smem(gs, lv.g, lv.has_plans ? &lv.planA : nullptr, lv.has_plans ? &lv.planB : nullptr, presolve, max_piv, &piv);
        QPN_SYNC();
        if (st != ST_SUCCESS) break;
        for (int j = i; j < nv; j += blockDim.x) xn[j] = xs[j];
        QPN_SYNC();
        for (int j = i; j < lv.nd_level; j += blockDim.x) xn[lv.dec[j]] = gs.zs()[j];
        QPN_SYNC();
        double dn = 0.0;
        for (int j = 0; j < nv; ++j) { const double e = xn[j] - xs[j]; dn = fma(e, e, dn); }
        if (sqrt(dn) < 1e-4) break;        // algorithm.jl:96-97: disagreement -> solved = false
        QPN_SYNC();
        for (int j = i; j < nv; j += blockDim.x) xs[j] = xn[j];
        QPN_SYNC();
    }
    for (int j = i; j < nv; j += blockDim.x) x_out[(size_t)b * nv + j] = xs[j];
    if (i == 0) {
        solved_out[b] = (uint8_t)solved; iters_out[b] = iters; pivots_out[b] = piv;
        if (hist && hist_count) hist_count[b] = nhist;
    }
}


// ---- solve_avi with a shared matrix: plan + solve ---------------------------------------------------
// Plan of a plain AVI (matrix and bounds shared by the batch).
__global__ void plan_build_avi_kernel(int n, const __grid_constant__ MatDesc M, const double* __restrict__ l,
                                      const double* __restrict__ u, double* __restrict__ T0, double* __restrict__ PT,
                                      int* __restrict__ rowvar0, int* __restrict__ colvar0, int* __restrict__ csr_ptr,
                                      int* __restrict__ csr_col, double* __restrict__ csr_val, int* __restrict__ hdr) {
    const int i = threadIdx.x;
    Tab t;
    tab_carve(t, n, n + 1, 0);
    double* qs = reinterpret_cast<double*>(qpn_smem + tab_smem_bytes(n, n + 1));
    double* zs = qs + n;
    int* cnt = reinterpret_cast<int*>(zs + n);
    if (i < n) { qs[i] = 0.0; zs[i] = 0.0; t.l()[i] = l[i]; t.u()[i] = u[i]; }
    load_neg_matrix(t, M, 0);
    plan_finish(t, n, 0, qs, zs, cnt, T0, PT, rowvar0, colvar0, csr_ptr, csr_col, csr_val, hdr);
}

// grid = batch, block = roundup32(n).  Dynamic smem: Tab_ex(n, n*rs(ncol0), rs(ncol0)) + q, z, zb (3n).
template <int MAXT>
__global__ void __launch_bounds__(MAXT, 896 / MAXT) avi_solve_plan_kernel(const __grid_constant__ PlanDesc P, int batch,
                                 const double* __restrict__ q, const double* __restrict__ l, const double* __restrict__ u,
                                 const double* __restrict__ z0, int max_pivots, double* __restrict__ z_out,
                                 int32_t* __restrict__ status_out, int32_t* __restrict__ pivots_out, int8_t* __restrict__ basis_out) {
    const int b = blockIdx.x, i = threadIdx.x, n = P.n;
    const int ldr = row_stride(P.ncol0);
    Tab t;
    tab_carve_ex(t, n, (size_t)n * ldr, ldr, 0);
    double* qs = reinterpret_cast<double*>(qpn_smem + tab_smem_bytes_ex(n, (size_t)n * ldr, ldr));
    double* zs = qs + n;
    double* zb = zs + n;
    if (i < n) {
        qs[i] = q[(size_t)b * n + i];
        zs[i] = z0[(size_t)b * n + i];
        t.l()[i] = l[i];
        t.u()[i] = u[i];
    }
    QPN_SYNC();
    int piv = 0;
    int8_t* code = reinterpret_cast<int8_t*>(zb);          // zb is free once the solve has started; reuse it for the codes
    // (solve_avi_plan writes the codes only after avi_pivot_run, when zb is no longer read)
    const int st = solve_avi_plan(t, P, qs, zs, zb, max_pivots, code, &piv);
    QPN_SYNC();
    if (i < n) {
        z_out[(size_t)b * n + i] = zs[i];
        if (basis_out) basis_out[(size_t)b * n + i] = code[i];
    }
    if (i == 0) { status_out[b] = st; pivots_out[b] = piv; }
}

}  // namespace qpn
