// Kernels of libqpn_cuda (sm_100a).  One CTA per instance for everything that pivots;
// one warp per (poly, point) / one thread per index for the coalesced evaluation kernels.
#pragma once
#include "avi_pivot.cuh"

namespace qpn {

// Device view of qpn_matrix.
struct MatDesc {
    const double* dense;
    const int32_t* colptr;
    const int32_t* rowval;
    const double* nzval;
    int nnz, base, shared;
};

// T[i][0:n] = -M[i][:] for instance b (row-major tableau).  Ends with a barrier.
__device__ __forceinline__ void load_neg_matrix(Tab& t, const MatDesc& M, int b) {
    const int n = t.n, ldr = t.ldr, i = threadIdx.x;
    if (M.dense) {
        const double* src = M.dense + (M.shared ? 0 : (size_t)b * n * n);
        if (i < n) {
            double* row = t.T() + (size_t)i * ldr;
            for (int j = 0; j < n; ++j) row[j] = -src[(size_t)j * n + i];      // coalesced across the warp
        }
    } else {
        if (i < n) {
            double* row = t.T() + (size_t)i * ldr;
            for (int j = 0; j < n; ++j) row[j] = 0.0;
        }
        QPN_SYNC();
        const double* nz = M.nzval + (M.shared ? 0 : (size_t)b * M.nnz);
        for (int j = 0; j < n; ++j) {
            const int k0 = M.colptr[j] - M.base, k1 = M.colptr[j + 1] - M.base;
            for (int k = k0 + i; k < k1; k += blockDim.x)
                t.T()[(size_t)(M.rowval[k] - M.base) * ldr + j] = -nz[k];
        }
    }
    QPN_SYNC();
}

// r_i = (M z)_i + q_i with T[i][0:n] = -M[i][:] and z in shared memory; sequential in j.
__device__ __forceinline__ double residual_row(const Tab& t, const double* zs, double qi, int i) {
    const double* row = t.T() + (size_t)i * t.ldr;
    double acc = 0.0;
    for (int j = 0; j < t.n; ++j) {
        const double mij = -row[j];
        if (mij != 0.0) acc = fma(mij, zs[j], acc);
    }
    return acc + qi;
}

// ---- solve_avi (avi.jl:63-77) ------------------------------------------------------------
// grid = batch, block = roundup32(n).  Dynamic smem: Tab(n, n+1) + q(n) + z(n).
template <int MAXT>
__global__ void __launch_bounds__(MAXT, 896 / MAXT) avi_solve_kernel(int n, int batch, const __grid_constant__ MatDesc M, const double* __restrict__ q,
                                 const double* __restrict__ l, const double* __restrict__ u,
                                 int lu_shared, const double* __restrict__ z0, int max_pivots,
                                 double* __restrict__ z_out, int32_t* __restrict__ status_out,
                                 int32_t* __restrict__ pivots_out, int8_t* __restrict__ basis_out) {
    const int b = blockIdx.x, i = threadIdx.x;
    Tab t;
    tab_carve(t, n, n + 1, 0);
    double* qs = reinterpret_cast<double*>(qpn_smem + tab_smem_bytes(n, n + 1));
    double* zs = qs + n;
    if (i < n) {
        qs[i] = q[(size_t)b * n + i];
        zs[i] = z0[(size_t)b * n + i];
        t.l()[i] = l[(lu_shared ? 0 : (size_t)b * n) + i];
        t.u()[i] = u[(lu_shared ? 0 : (size_t)b * n) + i];
    }
    load_neg_matrix(t, M, b);
    tab_start(t, qs, zs);
    const PivotResult pr = avi_pivot_run(t, max_pivots, false);
    const double zi = pr.zi; const int8_t code = (int8_t)pr.code;
    int st = pr.st;
    const int piv = pr.pivots;
    // final check (avi.jl:71-74) against the original matrix
    QPN_SYNC();
    if (i < n) zs[i] = zi;
    load_neg_matrix(t, M, b);
    int bad = 0;
    if (i < n) bad = check_avi_index(residual_row(t, zs, qs[i], i), zi, t.l()[i], t.u()[i], 1e-6);
    bad = QPN_SYNC_OR(bad);
    if (st == ST_SUCCESS && bad) st = ST_FAILURE;
    if (i < n) {
        z_out[(size_t)b * n + i] = zi;
        if (basis_out) basis_out[(size_t)b * n + i] = code;
    }
    if (i == 0) { status_out[b] = st; pivots_out[b] = piv; }
}

// ---- check_avi_solution (avi.jl:148-156): one warp per instance, lanes over rows -----------
__global__ void check_avi_kernel(int n, int batch, MatDesc M, const double* __restrict__ q,
                                 const double* __restrict__ l, const double* __restrict__ u,
                                 int lu_shared, const double* __restrict__ z, double tol,
                                 int32_t* __restrict__ bad_out, double* __restrict__ r_out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= batch) return;
    const int b = warp;
    const double* zb = z + (size_t)b * n;
    int bad = 0;
    for (int i = lane; i < n; i += 32) {
        double acc = 0.0;
        if (M.dense) {
            const double* src = M.dense + (M.shared ? 0 : (size_t)b * n * n);
            for (int j = 0; j < n; ++j) {
                const double mij = src[(size_t)j * n + i];
                if (mij != 0.0) acc = fma(mij, zb[j], acc);
            }
        } else {
            const double* nz = M.nzval + (M.shared ? 0 : (size_t)b * M.nnz);
            for (int j = 0; j < n; ++j) {           // row i of a CSC matrix: scan the columns
                const int k0 = M.colptr[j] - M.base, k1 = M.colptr[j + 1] - M.base;
                for (int k = k0; k < k1; ++k)
                    if (M.rowval[k] - M.base == i) acc = fma(nz[k], zb[j], acc);
            }
        }
        const double r = acc + q[(size_t)b * n + i];
        if (r_out) r_out[(size_t)b * n + i] = r;
        const size_t o = (lu_shared ? 0 : (size_t)b * n) + i;
        bad += check_avi_index(r, zb[i], l[o], u[o], tol);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
    if (lane == 0) bad_out[b] = bad;
}

// Device view of qpn_gavi.
struct GaviDesc {
    int d1, d2, np;
    const double *M, *N, *o, *l1, *u1, *A, *B, *l2, *u2;
};

// ---- comp_indices (avi_solutions.jl:511-612): one thread per (index, instance) -------------
__device__ __forceinline__ bool approx_eq(double a, double b, double atol) { return a == b || fabs(a - b) <= atol; }

__device__ __forceinline__ int8_t comp_mask(double l, double u, double r, double z, double tol) {
    const bool eq = approx_eq(l, u, tol);
    int m = 0;
    if (approx_eq(z, l, tol) && r >= -tol && !eq) m |= 1;
    if (l - tol <= z && z <= u + tol && approx_eq(r, 0.0, tol) && !eq) m |= 2;
    if (approx_eq(z, u, tol) && r <= tol && !eq) m |= 4;
    if (m == 0) m = eq ? 8 : 0;
    return (int8_t)m;
}

__global__ void comp_indices_kernel(GaviDesc g, int batch, const double* __restrict__ z,
                                    const double* __restrict__ w, double tol, int8_t* __restrict__ mask) {
    const int dz = g.d1 + g.d2;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)batch * dz) return;
    const int b = (int)(gid / dz), i = (int)(gid - (long long)b * dz);
    const double* zb = z + (size_t)b * dz;
    const double* wb = w + (size_t)b * g.np;
    if (i < g.d1) {
        double acc = 0.0, acc2 = 0.0;
        for (int j = 0; j < dz; ++j) acc = fma(g.M[(size_t)j * g.d1 + i], zb[j], acc);
        for (int j = 0; j < g.np; ++j) acc2 = fma(g.N[(size_t)j * g.d1 + i], wb[j], acc2);
        const double r = (acc + acc2) + g.o[i];
        mask[(size_t)b * dz + i] = comp_mask(g.l1[i], g.u1[i], r, zb[i], tol);
    } else {
        const int k = i - g.d1;
        double acc = 0.0, acc2 = 0.0;
        for (int j = 0; j < dz; ++j) acc = fma(g.A[(size_t)j * g.d2 + k], zb[j], acc);
        for (int j = 0; j < g.np; ++j) acc2 = fma(g.B[(size_t)j * g.d2 + k], wb[j], acc2);
        const double s = acc + acc2;
        mask[(size_t)b * dz + i] = comp_mask(g.l2[k], g.u2[k], zb[i], s, tol);
    }
}

// ---- Base.in(x, poly) (sets.jl:820-825,850-853): one warp per (poly, point) -----------------
// Lanes take rows of the poly; each row dot product is sequential in the coordinate index.
__global__ void halfspace_in_kernel(int npoly, int d, int mtot, const int32_t* __restrict__ poly_ptr,
                                    const double* __restrict__ A, const double* __restrict__ l,
                                    const double* __restrict__ u, const uint8_t* __restrict__ rl,
                                    const uint8_t* __restrict__ ru, int npts, const double* __restrict__ x,
                                    double tol, uint8_t* __restrict__ in_out) {
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= (long long)npoly * npts) return;
    const int p = (int)(warp % npoly), pt = (int)(warp / npoly);
    const double* xp = x + (size_t)pt * d;
    int ok = 1;
    for (int row = poly_ptr[p] + lane; row < poly_ptr[p + 1]; row += 32) {
        double ax = 0.0;
        for (int j = 0; j < d; ++j) ax = fma(A[(size_t)j * mtot + row], xp[j], ax);
        const bool lo = (rl && rl[row]) ? (l[row] - tol < ax) : (l[row] - tol <= ax);
        const bool up = (ru && ru[row]) ? (ax - tol < u[row]) : (ax - tol <= u[row]);
        if (!(lo && up)) ok = 0;
    }
    ok = __all_sync(0xffffffffu, ok);
    if (lane == 0) in_out[(size_t)pt * npoly + p] = (uint8_t)ok;
}

// ---- the same test for many points: point tiles x all rows -----------------------------------------------
// One CTA takes HS_TP points (staged in shared memory, coordinate-major so that a warp reads one word per
// step: a broadcast) and walks every row of every polyhedron: thread t owns rows t, t + blockDim, ... and keeps
// HS_PP accumulators -- one matrix entry is loaded once per HS_PP points instead of once per point, and each
// a'x is still one sequential fma chain over the coordinate index (same bits as the oracle).  Row verdicts are
// AND-ed per (point, poly) in shared memory.  fp64-FMA / shared-memory bound, not HBM bound:
// 2 m d flop per (poly, point) against 8 d bytes per point.
constexpr int HS_TP = 64;      // points per CTA
constexpr int HS_PP = 16;      // points per register pass
__global__ void __launch_bounds__(512) halfspace_in_tiled_kernel(int npoly, int d, int mtot, const int32_t* __restrict__ poly_ptr,
                                                                 const double* __restrict__ A, const double* __restrict__ l,
                                                                 const double* __restrict__ u, const uint8_t* __restrict__ rl,
                                                                 const uint8_t* __restrict__ ru, int npts, const double* __restrict__ x,
                                                                 double tol, uint8_t* __restrict__ in_out) {
    double* xs = reinterpret_cast<double*>(qpn_smem);                       // d x HS_TP, coordinate-major
    unsigned* okw = reinterpret_cast<unsigned*>(xs + (size_t)d * HS_TP);    // HS_TP x npoly verdict words
    int* rowpoly = reinterpret_cast<int*>(okw + (size_t)HS_TP * npoly);     // mtot: poly of each row
    const int p0 = blockIdx.x * HS_TP;
    const int np_here = min(HS_TP, npts - p0);
    for (int e = threadIdx.x; e < d * HS_TP; e += blockDim.x) {
        const int pt = e / d, j = e - pt * d;                              // coalesced read of x
        xs[(size_t)j * HS_TP + pt] = pt < np_here ? x[(size_t)(p0 + pt) * d + j] : 0.0;
    }
    for (int e = threadIdx.x; e < HS_TP * npoly; e += blockDim.x) okw[e] = 1u;
    for (int p = threadIdx.x; p < npoly; p += blockDim.x)
        for (int r = poly_ptr[p]; r < poly_ptr[p + 1]; ++r) rowpoly[r] = p;
    __syncthreads();
    for (int row = threadIdx.x; row < mtot; row += blockDim.x) {
        const double lo = l[row] - tol, up = u[row];
        const bool sl = rl && rl[row], su = ru && ru[row];
        const int p = rowpoly[row];
        for (int q0 = 0; q0 < np_here; q0 += HS_PP) {
            double acc[HS_PP];
#pragma unroll
            for (int q = 0; q < HS_PP; ++q) acc[q] = 0.0;
            for (int j = 0; j < d; ++j) {
                const double a = A[(size_t)j * mtot + row];
                const double2* xj = reinterpret_cast<const double2*>(xs + (size_t)j * HS_TP + q0);    // 16-byte aligned: HS_TP, q0 even
#pragma unroll
                for (int q = 0; q < HS_PP / 2; ++q) {
                    const double2 xv = xj[q];                                                        // one broadcast word pair per two fma
                    acc[2 * q] = fma(a, xv.x, acc[2 * q]);
                    acc[2 * q + 1] = fma(a, xv.y, acc[2 * q + 1]);
                }
            }
#pragma unroll
            for (int q = 0; q < HS_PP; ++q) {
                const double ax = acc[q];
                const bool okl = sl ? (lo < ax) : (lo <= ax);
                const bool oku = su ? (ax - tol < up) : (ax - tol <= up);
                if (!(okl && oku) && q0 + q < np_here) okw[(size_t)(q0 + q) * npoly + p] = 0u;     // benign race: everyone writes 0
            }
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < np_here * npoly; e += blockDim.x) in_out[(size_t)p0 * npoly + e] = (uint8_t)okw[e];
}

}  // namespace qpn
