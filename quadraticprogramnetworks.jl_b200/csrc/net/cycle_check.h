// The cycle check of solve_base! (/root/reference/src/algorithm.jl:14-30): the projections of the current iterate
// against one earlier entry of the (instance, level) history, isapprox with rtol = sqrt(eps).  Shared by the device
// kernel (net_cycle_kernel) and the host backends so that all of them round alike.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define QPN_CC_HD __host__ __device__
#else
#define QPN_CC_HD
#endif

QPN_CC_HD inline bool qpn_cycle_hit(const double* pv, const double* prev, int n) {
    double dd = 0.0, na = 0.0, nb = 0.0;
    for (int k = 0; k < n; ++k) {
        const double e = pv[k] - prev[k];
        dd = fma(e, e, dd); na = fma(pv[k], pv[k], na); nb = fma(prev[k], prev[k], nb);
    }
    return sqrt(dd) <= 1.4901161193847656e-8 * fmax(sqrt(na), sqrt(nb));
}
