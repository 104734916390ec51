// libqpn_cuda: host side of the C ABI declared in include/qpn_cuda.h.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <deque>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <vector>

#include "../../include/qpn_cuda.h"
#include "qpn_kernels.cuh"
#include "qpn_level.cuh"
#include "qpn_big.cuh"
#include "qpn_level_big.cuh"

using namespace qpn;

struct qpn_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;
    int max_smem_optin = 0;
    int sm_count = 0;
    // grow-only device scratch arena for the host-pointer entry points
    unsigned char* arena = nullptr;
    size_t arena_bytes = 0;
    size_t arena_used = 0;
    // grow-only buffers for the plans of one-off calls (cudaMalloc / cudaFree per call cost milliseconds)
    unsigned char* plan_buf[2] = {nullptr, nullptr};
    size_t plan_buf_bytes[2] = {0, 0};
    // global-memory tableau slots of the big path (avi_pivot_big.cuh), grow-only
    double* big_work = nullptr;
    size_t big_work_doubles = 0;
    int64_t big_launches = 0;   // launches that took the global-memory tableau path
    int force_big = 0;          // option "force_big": route every pivoting solve through the big path (tests)
    int big_ctas_per_sm = 0;    // option "big_ctas_per_sm": 0 = as many as fit
    int big_smem_threads = 0;   // option "big_smem_threads": threads per CTA when the slot is in shared memory (0 = by size)
    int big_slot_in_smem = 1;   // option "big_slot_in_smem": 0 = always keep the tableau slot in global memory (tests)
};

static std::string g_create_error;

#ifdef QPN_TRACE
// Debug build: qpn_trace_enable(h, batch) maps a pinned host buffer of 4 ints per CTA.
static int* g_trace_host = nullptr;
extern "C" int* qpn_trace_enable(qpn_handle* h, int batch) {
    int* dptr = nullptr;
    if (g_trace_host) cudaFreeHost(g_trace_host);
    if (cudaHostAlloc((void**)&g_trace_host, sizeof(int) * 16 * (size_t)batch, cudaHostAllocMapped) != cudaSuccess) return nullptr;
    for (size_t k = 0; k < 16 * (size_t)batch; ++k) g_trace_host[k] = -1;
    cudaHostGetDevicePointer((void**)&dptr, g_trace_host, 0);
    cudaMemcpyToSymbol(qpn::qpn_trace_ptr, &dptr, sizeof(dptr));
    return g_trace_host;
}
#endif

static int fail(qpn_handle* h, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return -1;
}

#define CK(call)                                                                           \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) return fail(h, "%s: %s", #call, cudaGetErrorString(e_));    \
    } while (0)

static inline int roundup32(int n) { return n < 32 ? 32 : ((n + 31) / 32) * 32; }

// Kernels that pivot are instantiated per block-size bucket so that __launch_bounds__ can cap the
// registers at 768 resident threads per SM (24 CTAs of 32 threads: one wave for 4,096 instances / 148 SMs... nearly).
// The dynamic shared-memory limit of a kernel is an attribute of the FUNCTION, shared by every host thread of the
// process: handles of different threads launch the same kernels with different sizes, so a per-launch value would race
// (thread A raises the limit, thread B lowers it, A's launch fails with "invalid argument").  Every launch that needs more
// than the default 48 KB therefore sets the same value: the device's opt-in maximum.
static int g_smem_cap = 0;
static inline int qpn_smem_cap() { return g_smem_cap; }
#define QPN_LAUNCH_BUCKETED(KERNEL, threads, grid, smem, stream, ...)                                            \
    do {                                                                                                         \
        const int thr_ = (threads);                                                                              \
        if (thr_ <= 32) { QPN_LAUNCH_ONE(KERNEL<32>, 32, grid, smem, stream, __VA_ARGS__); }                     \
        else if (thr_ <= 64) { QPN_LAUNCH_ONE(KERNEL<64>, thr_, grid, smem, stream, __VA_ARGS__); }              \
        else if (thr_ <= 128) { QPN_LAUNCH_ONE(KERNEL<128>, thr_, grid, smem, stream, __VA_ARGS__); }            \
        else { QPN_LAUNCH_ONE(KERNEL<256>, thr_, grid, smem, stream, __VA_ARGS__); }                             \
    } while (0)
#define QPN_LAUNCH_ONE(KINST, thr, grid, smem, stream, ...)                                                      \
    do {                                                                                                         \
        if ((smem) > 48 * 1024) CK(cudaFuncSetAttribute(KINST, cudaFuncAttributeMaxDynamicSharedMemorySize, qpn_smem_cap())); \
        KINST<<<(grid), (thr), (smem), (stream)>>>(__VA_ARGS__);                                                 \
    } while (0)

extern "C" int qpn_create(int device, qpn_handle** out) {
    qpn_handle* h = nullptr;
    if (!out) return fail(nullptr, "qpn_create: out is NULL");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, "qpn_create: no CUDA device (%s); libqpn_cuda has no CPU fallback",
                    cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(nullptr, "qpn_create: device %d out of range (%d)", device, count);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return fail(nullptr, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, "qpn_create: device %d is sm_%d%d; this library is built for sm_100a only",
                    device, prop.major, prop.minor);
    h = new qpn_handle();
    h->device = device;
    h->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    g_smem_cap = h->max_smem_optin;                  // (every handle of the process sits on the same kind of device)
    h->sm_count = prop.multiProcessorCount;
    if ((e = cudaSetDevice(device)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete h;
        return fail(nullptr, "qpn_create: %s", cudaGetErrorString(e));
    }
    *out = h;
    return 0;
}

extern "C" int qpn_destroy(qpn_handle* h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    if (h->arena) cudaFree(h->arena);
    for (int k = 0; k < 2; ++k) if (h->plan_buf[k]) cudaFree(h->plan_buf[k]);
    if (h->big_work) cudaFree(h->big_work);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

extern "C" const char* qpn_last_error(qpn_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }
extern "C" int qpn_device(qpn_handle* h) { return h ? h->device : -1; }
extern "C" int64_t qpn_launch_count(qpn_handle* h) { return h ? h->launches : 0; }
extern "C" int64_t qpn_big_launch_count(qpn_handle* h) { return h ? h->big_launches : 0; }
extern "C" int qpn_synchronize(qpn_handle* h) {
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int qpn_set_option(qpn_handle* h, const char* name, int64_t value) {
    if (!h || !name) return -1;
    if (!strcmp(name, "force_big")) { h->force_big = value != 0; return 0; }
    if (!strcmp(name, "big_ctas_per_sm")) { h->big_ctas_per_sm = (int)value; return 0; }
    if (!strcmp(name, "big_slot_in_smem")) { h->big_slot_in_smem = value != 0; return 0; }
    if (!strcmp(name, "big_smem_threads")) { h->big_smem_threads = (int)value; return 0; }
    return fail(h, "qpn_set_option: unknown option '%s'", name);
}

extern "C" int qpn_malloc(qpn_handle* h, size_t bytes, void** dptr) {
    CK(cudaSetDevice(h->device));
    CK(cudaMalloc(dptr, bytes ? bytes : 8));
    return 0;
}
extern "C" int qpn_free(qpn_handle* h, void* dptr) {
    CK(cudaSetDevice(h->device));
    CK(cudaFree(dptr));
    return 0;
}
extern "C" int qpn_memcpy_h2d(qpn_handle* h, void* dst, const void* src, size_t bytes) {
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}
extern "C" int qpn_memcpy_d2h(qpn_handle* h, void* dst, const void* src, size_t bytes) {
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int qpn_host_register(qpn_handle* h, void* ptr, size_t bytes) {
    if (!h || !ptr) return -1;
    CK(cudaSetDevice(h->device));
    CK(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    return 0;
}
extern "C" int qpn_host_unregister(qpn_handle* h, void* ptr) {
    if (!h || !ptr) return -1;
    CK(cudaSetDevice(h->device));
    CK(cudaHostUnregister(ptr));
    return 0;
}
extern "C" int qpn_abi_struct_sizes(int32_t* out) {
    if (!out) return -1;
    out[0] = (int32_t)sizeof(qpn_matrix); out[1] = (int32_t)sizeof(qpn_gavi); out[2] = (int32_t)sizeof(qpn_node);
    out[3] = (int32_t)sizeof(qpn_level); out[4] = (int32_t)sizeof(qpn_net_desc); out[5] = 0;
    return 0;
}

// ---- scratch arena -------------------------------------------------------------------------
// Host-pointer entry points stage their operands here.  reset -> reserve (grow once) -> take.
struct Arena {
    qpn_handle* h;
    size_t need = 0;
    std::vector<std::pair<size_t, size_t>> slots;   // (offset, bytes)
    explicit Arena(qpn_handle* hh) : h(hh) {}
    int add(size_t bytes) {
        size_t off = (need + 255) & ~(size_t)255;
        need = off + bytes;
        slots.push_back({off, bytes});
        return (int)slots.size() - 1;
    }
    int commit() {
        if (need > h->arena_bytes) {
            if (h->arena) cudaFree(h->arena);
            h->arena = nullptr; h->arena_bytes = 0;
            size_t cap = need + need / 4 + 4096;
            cudaError_t e = cudaMalloc((void**)&h->arena, cap);
            if (e != cudaSuccess) return fail(h, "cudaMalloc(%zu): %s", cap, cudaGetErrorString(e));
            h->arena_bytes = cap;
        }
        return 0;
    }
    template <typename T> T* ptr(int slot) { return slot < 0 ? nullptr : reinterpret_cast<T*>(h->arena + slots[slot].first); }
};

static int up(qpn_handle* h, void* dst, const void* src, size_t bytes) {
    if (!bytes || !src) return 0;
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
    return 0;
}
static int down(qpn_handle* h, void* dst, const void* src, size_t bytes) {
    if (!bytes || !dst) return 0;
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
    return 0;
}

static int check_matrix(qpn_handle* h, const qpn_matrix* M, int n) {
    if (!M) return fail(h, "matrix descriptor is NULL");
    if (!M->dense && !(M->colptr && M->rowval && M->nzval)) return fail(h, "matrix: neither dense nor CSC given");
    if (n <= 0) return fail(h, "n must be positive");
    return 0;
}

static MatDesc to_desc(const qpn_matrix* M) {
    MatDesc d;
    d.dense = M->dense; d.colptr = M->colptr; d.rowval = M->rowval; d.nzval = M->nzval;
    d.nnz = M->nnz; d.base = M->index_base; d.shared = M->is_shared;
    return d;
}

// Stage a host qpn_matrix into the arena; returns the device descriptor.
struct MatSlots { int dense = -1, colptr = -1, rowval = -1, nzval = -1; };
static MatSlots plan_matrix(Arena& a, const qpn_matrix* M, int n, int batch) {
    MatSlots s;
    size_t reps = M->is_shared ? 1 : (size_t)batch;
    if (M->dense) s.dense = a.add(sizeof(double) * (size_t)n * n * reps);
    else {
        s.colptr = a.add(sizeof(int32_t) * (size_t)(n + 1));
        s.rowval = a.add(sizeof(int32_t) * (size_t)M->nnz);
        s.nzval = a.add(sizeof(double) * (size_t)M->nnz * reps);
    }
    return s;
}
static int stage_matrix(qpn_handle* h, Arena& a, const MatSlots& s, const qpn_matrix* M, int n, int batch, MatDesc* out) {
    size_t reps = M->is_shared ? 1 : (size_t)batch;
    MatDesc d = to_desc(M);
    if (M->dense) {
        if (up(h, a.ptr<double>(s.dense), M->dense, sizeof(double) * (size_t)n * n * reps)) return -1;
        d.dense = a.ptr<double>(s.dense);
    } else {
        if (up(h, a.ptr<int32_t>(s.colptr), M->colptr, sizeof(int32_t) * (size_t)(n + 1))) return -1;
        if (up(h, a.ptr<int32_t>(s.rowval), M->rowval, sizeof(int32_t) * (size_t)M->nnz)) return -1;
        if (up(h, a.ptr<double>(s.nzval), M->nzval, sizeof(double) * (size_t)M->nnz * reps)) return -1;
        d.colptr = a.ptr<int32_t>(s.colptr); d.rowval = a.ptr<int32_t>(s.rowval); d.nzval = a.ptr<double>(s.nzval);
    }
    *out = d;
    return 0;
}

// ---- plans -------------------------------------------------------------------------------------
// One plan: buffers sized for the worst case, filled by plan_build_kernel.
// blob_out != NULL: a fresh allocation owned by the caller (resident levels); blob_out == NULL: the
// handle's reusable buffer for one-off calls.
static bool big_needed(qpn_handle* h, int n, size_t smem_small);
static int build_plan_big(qpn_handle* h, const GaviDesc& g, int kind, int n_avi, const MatDesc& M, const double* l, const double* u,
                          int slot, PlanDesc* out, unsigned char** blob_out);

static int build_plan(qpn_handle* h, const GaviDesc& g, int kind, PlanDesc* out, unsigned char** blob_out) {
    const size_t n = (size_t)g.d1 + 2 * g.d2, dz = (size_t)g.d1 + g.d2;
    if (big_needed(h, (int)n, gavi_smem_bytes(g.d1, g.d2, g.np))) {
        MatDesc m0;
        memset(&m0, 0, sizeof m0);
        return build_plan_big(h, g, kind, 0, m0, nullptr, nullptr, kind, out, blob_out);
    }
    const size_t ldrw = row_stride((int)n + 1);
    size_t need = 0;
    auto add = [&](size_t bytes) { size_t off = (need + 255) & ~(size_t)255; need = off + bytes; return off; };
    const size_t oT0 = add(8 * n * ldrw), oPT = add(8 * n * n), oval = add(8 * n * n), orv = add(4 * n), ocv = add(4 * (n + 1)),
                 optr = add(4 * (n + 1)), ocol = add(4 * n * n), ocols = add(4 * (dz + 1)), ohdr = add(32);
    unsigned char* b = nullptr;
    const size_t smem = gavi_smem_bytes(g.d1, g.d2, g.np);
    if (smem > (size_t)h->max_smem_optin) return fail(h, "plan: d1=%d d2=%d needs %zu B shared memory", g.d1, g.d2, smem);
    if (blob_out) {
        CK(cudaMalloc((void**)&b, need + 256));
    } else {
        if (need + 256 > h->plan_buf_bytes[kind]) {
            if (h->plan_buf[kind]) cudaFree(h->plan_buf[kind]);
            h->plan_buf[kind] = nullptr; h->plan_buf_bytes[kind] = 0;
            CK(cudaMalloc((void**)&h->plan_buf[kind], 2 * (need + 256)));
            h->plan_buf_bytes[kind] = 2 * (need + 256);
        }
        b = h->plan_buf[kind];
    }
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(plan_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, qpn_smem_cap()));
    plan_build_kernel<<<1, roundup32((int)n), smem, h->stream>>>(g, kind, (double*)(b + oT0), (double*)(b + oPT), (int*)(b + orv), (int*)(b + ocv),
                                                                  (int*)(b + optr), (int*)(b + ocol), (double*)(b + oval), (int*)(b + ocols),
                                                                  (int*)(b + ohdr));
    h->launches++;
    CK(cudaGetLastError());
    int hdr[5];
    CK(cudaMemcpyAsync(hdr, b + ohdr, sizeof hdr, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    PlanDesc P;
    P.n = kind == 1 ? hdr[3] + 2 * g.d2 : (int)n;
    P.ncol0 = hdr[0]; P.npiv0 = hdr[1]; P.tcol0 = hdr[2]; P.nact = hdr[4];
    P.T0 = (double*)(b + oT0); P.PT = (double*)(b + oPT); P.rowvar0 = (int*)(b + orv); P.colvar0 = (int*)(b + ocv);
    P.csr_ptr = (int*)(b + optr); P.csr_col = (int*)(b + ocol); P.csr_val = (double*)(b + oval);
    P.cols = (int*)(b + ocols); P.ncols = hdr[3];
    *out = P;
    if (blob_out) *blob_out = b;
    return 0;
}


// Below this batch size a one-off call does not repay the two plan launches and their sync.
static const int QPN_PLAN_MIN_BATCH = 256;


// Plan of a plain AVI with a shared matrix and shared bounds, in the handle's reusable buffer (slot 0).
static int build_plan_avi(qpn_handle* h, int n_, const MatDesc& M, const double* l, const double* u, PlanDesc* out) {
    const size_t n = n_, ldrw = row_stride(n_ + 1);
    size_t need = 0;
    auto add = [&](size_t bytes) { size_t off = (need + 255) & ~(size_t)255; need = off + bytes; return off; };
    const size_t oT0 = add(8 * n * ldrw), oPT = add(8 * n * n), oval = add(8 * n * n), orv = add(4 * n), ocv = add(4 * (n + 1)),
                 optr = add(4 * (n + 1)), ocol = add(4 * n * n), ohdr = add(32);
    const size_t smem = tab_smem_bytes(n_, n_ + 1) + 8 * 3 * n;
    if (smem > (size_t)h->max_smem_optin) return fail(h, "plan: AVI of size n=%d needs %zu B shared memory", n_, smem);
    if (need + 256 > h->plan_buf_bytes[0]) {
        if (h->plan_buf[0]) cudaFree(h->plan_buf[0]);
        h->plan_buf[0] = nullptr; h->plan_buf_bytes[0] = 0;
        CK(cudaMalloc((void**)&h->plan_buf[0], 2 * (need + 256)));
        h->plan_buf_bytes[0] = 2 * (need + 256);
    }
    unsigned char* b = h->plan_buf[0];
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(plan_build_avi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, qpn_smem_cap()));
    plan_build_avi_kernel<<<1, roundup32(n_), smem, h->stream>>>(n_, M, l, u, (double*)(b + oT0), (double*)(b + oPT), (int*)(b + orv),
                                                                  (int*)(b + ocv), (int*)(b + optr), (int*)(b + ocol), (double*)(b + oval),
                                                                  (int*)(b + ohdr));
    h->launches++;
    CK(cudaGetLastError());
    int hdr[5];
    CK(cudaMemcpyAsync(hdr, b + ohdr, sizeof hdr, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    PlanDesc P;
    P.n = n_; P.ncol0 = hdr[0]; P.npiv0 = hdr[1]; P.tcol0 = hdr[2]; P.nact = hdr[4];
    P.T0 = (double*)(b + oT0); P.PT = (double*)(b + oPT); P.rowvar0 = (int*)(b + orv); P.colvar0 = (int*)(b + ocv);
    P.csr_ptr = (int*)(b + optr); P.csr_col = (int*)(b + ocol); P.csr_val = (double*)(b + oval);
    P.cols = nullptr; P.ncols = 0;
    *out = P;
    return 0;
}

// ---- big path (global-memory tableau) ----------------------------------------------------------
// Largest n the big path takes: its vectors must fit shared memory next to a kernel's extras.
static bool big_needed(qpn_handle* h, int n, size_t smem_small) {
    return h->force_big || n > 256 || smem_small > (size_t)h->max_smem_optin;
}

// Grid of persistent CTAs for a big kernel and the workspace for their slots.
// smem_io: in = bytes of the kernel's own layout, out = bytes to launch with (the layout plus room to stage the queued
// pivot rows of a flush, as far as the SM allows; the kernel finds the spare part through %dynamic_smem_size).
template <class K>
static int big_grid_ex(qpn_handle* h, K kernel, int nmax, size_t slot_doubles, size_t* smem_io, int batch, int* grid_out);
template <class K>
static int big_grid(qpn_handle* h, K kernel, int nmax, size_t* smem_io, int batch, int* grid_out) {
    return big_grid_ex(h, kernel, nmax, big_slot_doubles(nmax), smem_io, batch, grid_out);
}
template <class K>
static int big_grid_ex(qpn_handle* h, K kernel, int nmax, size_t slot_doubles, size_t* smem_io, int batch, int* grid_out) {
    size_t smem = *smem_io;
    if (smem > (size_t)h->max_smem_optin)
        return fail(h, "size n=%d needs %zu B of shared memory per CTA on the global-memory tableau path (limit %d)", nmax, smem,
                    h->max_smem_optin);
    smem = ((smem + 15) & ~(size_t)15) + 8 * (size_t)QPN_BIG_PEND * row_stride(nmax + 1);
    if (smem > (size_t)h->max_smem_optin) smem = (size_t)h->max_smem_optin;
    *smem_io = smem;
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, qpn_smem_cap()));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, QPN_BIG_THREADS, smem));
    if (per_sm < 1) return fail(h, "big path: kernel does not fit an SM (n=%d, %zu B shared memory)", nmax, smem);
    if (h->big_ctas_per_sm > 0 && per_sm > h->big_ctas_per_sm) per_sm = h->big_ctas_per_sm;
    int grid = per_sm * h->sm_count;
    if (grid > batch) grid = batch;
    const size_t need = (size_t)grid * slot_doubles;
    if (need > h->big_work_doubles) {
        CK(cudaDeviceSynchronize());
        if (h->big_work) cudaFree(h->big_work);
        h->big_work = nullptr; h->big_work_doubles = 0;
        cudaError_t e = cudaMalloc((void**)&h->big_work, 8 * need);
        if (e != cudaSuccess) return fail(h, "big path: cudaMalloc(%zu B of tableau slots): %s", 8 * need, cudaGetErrorString(e));
        h->big_work_doubles = need;
    }
    *grid_out = grid;
    h->big_launches++;
    return 0;
}

// Plan through the big kernel.  kind 0 / 1: the GAVI's lifted / presolve AVI; kind 2: plain AVI (M, l, u of size n_avi).
// Buffers as in build_plan: blob_out != NULL -> fresh allocation owned by the caller, else the handle's slot `slot`.
static int build_plan_big(qpn_handle* h, const GaviDesc& g, int kind, int n_avi, const MatDesc& M, const double* l, const double* u,
                          int slot, PlanDesc* out, unsigned char** blob_out) {
    const size_t n = kind == 2 ? (size_t)n_avi : (size_t)g.d1 + 2 * g.d2, dz = (size_t)g.d1 + g.d2;
    const size_t ldrw = row_stride((int)n + 1);
    size_t need = 0;
    auto add = [&](size_t bytes) { size_t off = (need + 255) & ~(size_t)255; need = off + bytes; return off; };
    const size_t oT0 = add(8 * n * ldrw), oPT = add(8 * n * n), oval = add(8 * n * n), orv = add(4 * n), ocv = add(4 * (n + 1)),
                 optr = add(4 * (n + 1)), ocol = add(4 * n * n), ocols = add(4 * (dz + 1)), ohdr = add(32);
    const size_t smem = big_smem_bytes((int)n) + (kind == 2 ? 16 * n : gavi_extra_bytes(g.d1, g.d2, g.np)) + 4 * n + 16;
    int grid = 0;
    size_t smem_l = smem;
    if (big_grid(h, plan_build_big_kernel, (int)n, &smem_l, 1, &grid)) return -1;
    unsigned char* b = nullptr;
    if (blob_out) {
        CK(cudaMalloc((void**)&b, need + 256));
    } else {
        if (need + 256 > h->plan_buf_bytes[slot]) {
            if (h->plan_buf[slot]) cudaFree(h->plan_buf[slot]);
            h->plan_buf[slot] = nullptr; h->plan_buf_bytes[slot] = 0;
            CK(cudaMalloc((void**)&h->plan_buf[slot], need + 256));
            h->plan_buf_bytes[slot] = need + 256;
        }
        b = h->plan_buf[slot];
    }
    plan_build_big_kernel<<<1, QPN_BIG_THREADS, smem_l, h->stream>>>(g, kind, n_avi, M, l, u, (double*)(b + oT0), (double*)(b + oPT),
                                                                     (int*)(b + orv), (int*)(b + ocv), (int*)(b + optr), (int*)(b + ocol),
                                                                     (double*)(b + oval), (int*)(b + ocols), (int*)(b + ohdr), h->big_work,
                                                                     (int)smem);
    h->launches++;
    CK(cudaGetLastError());
    int hdr[5];
    CK(cudaMemcpyAsync(hdr, b + ohdr, sizeof hdr, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    PlanDesc P;
    P.n = kind == 1 ? hdr[3] + 2 * g.d2 : (int)n;
    P.ncol0 = hdr[0]; P.npiv0 = hdr[1]; P.tcol0 = hdr[2]; P.nact = hdr[4];
    P.T0 = (double*)(b + oT0); P.PT = (double*)(b + oPT); P.rowvar0 = (int*)(b + orv); P.colvar0 = (int*)(b + ocv);
    P.csr_ptr = (int*)(b + optr); P.csr_col = (int*)(b + ocol); P.csr_val = (double*)(b + oval);
    P.cols = (int*)(b + ocols); P.ncols = hdr[3];
    *out = P;
    if (blob_out) *blob_out = b;
    return 0;
}

static int launch_avi_big(qpn_handle* h, int n, int batch, const MatDesc& M, const double* q, const double* l, const double* u,
                          int lu_shared, const double* z0, int max_pivots, double* z, int32_t* st, int32_t* pv, int8_t* basis,
                          cudaStream_t s) {
    const size_t smem = big_smem_bytes(n) + 8 * 3 * (size_t)n + (((size_t)n + 15) & ~(size_t)15);
    int grid = 0;
    size_t smem_l = smem;
    if (big_grid(h, avi_solve_big_kernel, n, &smem_l, batch, &grid)) return -1;
    PlanDesc P;
    memset(&P, 0, sizeof P);
    int has_plan = 0;
    if (M.shared && lu_shared && batch >= 2 && s == h->stream) {      // plans are built (and synchronised) on the handle's stream
        GaviDesc g0;
        memset(&g0, 0, sizeof g0);
        if (build_plan_big(h, g0, 2, n, M, l, u, 0, &P, nullptr)) return -1;
        has_plan = 1;
    }
    avi_solve_big_kernel<<<grid, QPN_BIG_THREADS, smem_l, s>>>(n, batch, M, P, has_plan, q, l, u, lu_shared, z0, max_pivots, z, st, pv,
                                                               basis, h->big_work, big_slot_doubles(n), (int)smem);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

static int launch_gavi_big(qpn_handle* h, const GaviDesc& g, int batch, const double* w, const double* z0, int presolve,
                           int max_pivots, double* z, double* zfull, int32_t* st, int32_t* pv, int8_t* basis, cudaStream_t s) {
    const int n = g.d1 + 2 * g.d2;
    const size_t smem = big_smem_bytes(n) + gavi_extra_bytes(g.d1, g.d2, g.np);
    int grid = 0;
    size_t smem_l = smem;
    if (big_grid(h, gavi_solve_big_kernel, n, &smem_l, batch, &grid)) return -1;
    GaviPlans plans;
    memset(&plans, 0, sizeof plans);
    if (batch >= 2 && s == h->stream) {
        MatDesc m0;
        memset(&m0, 0, sizeof m0);
        if (build_plan_big(h, g, 0, 0, m0, nullptr, nullptr, 0, &plans.A, nullptr) ||
            build_plan_big(h, g, 1, 0, m0, nullptr, nullptr, 1, &plans.B, nullptr)) return -1;
        plans.has = 1;
    }
    if (max_pivots <= 0) max_pivots = 50 * n + 100;
    gavi_solve_big_kernel<<<grid, QPN_BIG_THREADS, smem_l, s>>>(g, plans, batch, w, z0, presolve, max_pivots, z, zfull, st, pv, basis,
                                                                h->big_work, big_slot_doubles(n), (int)smem);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

// ---- solve_avi -----------------------------------------------------------------------------
static int launch_avi(qpn_handle* h, int n, int batch, const MatDesc& M, const double* q, const double* l,
                      const double* u, int lu_shared, const double* z0, int max_pivots, double* z,
                      int32_t* st, int32_t* pv, int8_t* basis, cudaStream_t s) {
    if (batch <= 0) return 0;
    const size_t smem = tab_smem_bytes(n, n + 1) + 2 * sizeof(double) * (size_t)n;
    if (max_pivots <= 0) max_pivots = 50 * n + 100;
    if (big_needed(h, n, smem)) return launch_avi_big(h, n, batch, M, q, l, u, lu_shared, z0, max_pivots, z, st, pv, basis, s);
    if (M.shared && lu_shared && batch >= QPN_PLAN_MIN_BATCH && s == h->stream) {
        // matrix and bounds shared by a large batch: the crash prefix is computed once
        PlanDesc P;
        if (build_plan_avi(h, n, M, l, u, &P)) return -1;
        const int ldr0 = row_stride(P.ncol0);
        const size_t smem_p = tab_smem_bytes_ex(n, (size_t)n * ldr0, ldr0) + 8 * 3 * (size_t)n;
        QPN_LAUNCH_BUCKETED(avi_solve_plan_kernel, roundup32(n), batch, smem_p, s, P, batch, q, l, u, z0, max_pivots, z, st, pv, basis);
        h->launches++;
        CK(cudaGetLastError());
        return 0;
    }
    QPN_LAUNCH_BUCKETED(avi_solve_kernel, roundup32(n), batch, smem, s, n, batch, M, q, l, u, lu_shared, z0, max_pivots, z, st, pv, basis);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int qpn_avi_solve_batched_dev(qpn_handle* h, int n, int batch, const qpn_matrix* M, const double* q,
                                         const double* l, const double* u, int lu_is_shared, const double* z0,
                                         int max_pivots, double* z_out, int32_t* status_out, int32_t* pivots_out,
                                         int8_t* basis_out, void* stream) {
    if (!h) return -1;
    if (check_matrix(h, M, n)) return -1;
    CK(cudaSetDevice(h->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
    return launch_avi(h, n, batch, to_desc(M), q, l, u, lu_is_shared, z0, max_pivots, z_out, status_out, pivots_out, basis_out, s);
}

extern "C" int qpn_avi_solve_batched(qpn_handle* h, int n, int batch, const qpn_matrix* M, const double* q,
                                     const double* l, const double* u, int lu_is_shared, const double* z0,
                                     int max_pivots, double* z_out, int32_t* status_out, int32_t* pivots_out,
                                     int8_t* basis_out) {
    if (!h) return -1;
    if (check_matrix(h, M, n)) return -1;
    CK(cudaSetDevice(h->device));
    const size_t nb = (size_t)n * batch, lub = lu_is_shared ? (size_t)n : nb;
    Arena a(h);
    MatSlots ms = plan_matrix(a, M, n, batch);
    int sq = a.add(8 * nb), sl = a.add(8 * lub), su = a.add(8 * lub), sz0 = a.add(8 * nb), sz = a.add(8 * nb);
    int sst = a.add(4 * (size_t)batch), spv = a.add(4 * (size_t)batch), sb = basis_out ? a.add(nb) : -1;
    if (a.commit()) return -1;
    MatDesc d;
    if (stage_matrix(h, a, ms, M, n, batch, &d)) return -1;
    if (up(h, a.ptr<double>(sq), q, 8 * nb) || up(h, a.ptr<double>(sl), l, 8 * lub) || up(h, a.ptr<double>(su), u, 8 * lub) ||
        up(h, a.ptr<double>(sz0), z0, 8 * nb)) return -1;
    if (launch_avi(h, n, batch, d, a.ptr<double>(sq), a.ptr<double>(sl), a.ptr<double>(su), lu_is_shared, a.ptr<double>(sz0),
                   max_pivots, a.ptr<double>(sz), a.ptr<int32_t>(sst), a.ptr<int32_t>(spv), a.ptr<int8_t>(sb), h->stream)) return -1;
    if (down(h, z_out, a.ptr<double>(sz), 8 * nb) || down(h, status_out, a.ptr<int32_t>(sst), 4 * (size_t)batch) ||
        down(h, pivots_out, a.ptr<int32_t>(spv), 4 * (size_t)batch) || down(h, basis_out, a.ptr<int8_t>(sb), nb)) return -1;
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

// ---- check_avi_solution ----------------------------------------------------------------------
extern "C" int qpn_check_avi_batched(qpn_handle* h, int n, int batch, const qpn_matrix* M, const double* q,
                                     const double* l, const double* u, int lu_is_shared, const double* z, double tol,
                                     int32_t* bad_out, double* r_out) {
    if (!h) return -1;
    if (check_matrix(h, M, n)) return -1;
    if (batch <= 0) return 0;
    CK(cudaSetDevice(h->device));
    const size_t nb = (size_t)n * batch, lub = lu_is_shared ? (size_t)n : nb;
    Arena a(h);
    MatSlots ms = plan_matrix(a, M, n, batch);
    int sq = a.add(8 * nb), sl = a.add(8 * lub), su = a.add(8 * lub), sz = a.add(8 * nb);
    int sbad = a.add(4 * (size_t)batch), sr = r_out ? a.add(8 * nb) : -1;
    if (a.commit()) return -1;
    MatDesc d;
    if (stage_matrix(h, a, ms, M, n, batch, &d)) return -1;
    if (up(h, a.ptr<double>(sq), q, 8 * nb) || up(h, a.ptr<double>(sl), l, 8 * lub) || up(h, a.ptr<double>(su), u, 8 * lub) ||
        up(h, a.ptr<double>(sz), z, 8 * nb)) return -1;
    const int threads = 128, warps_per_block = threads / 32;
    check_avi_kernel<<<(batch + warps_per_block - 1) / warps_per_block, threads, 0, h->stream>>>(
        n, batch, d, a.ptr<double>(sq), a.ptr<double>(sl), a.ptr<double>(su), lu_is_shared, a.ptr<double>(sz), tol,
        a.ptr<int32_t>(sbad), a.ptr<double>(sr));
    h->launches++;
    CK(cudaGetLastError());
    if (down(h, bad_out, a.ptr<int32_t>(sbad), 4 * (size_t)batch) || down(h, r_out, a.ptr<double>(sr), 8 * nb)) return -1;
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

// ---- GAVI staging ----------------------------------------------------------------------------
struct GaviSlots { int M, N, o, l1, u1, A, B, l2, u2; };
static GaviSlots plan_gavi(Arena& a, const qpn_gavi* g) {
    const size_t d1 = g->d1, d2 = g->d2, np = g->np, dz = d1 + d2;
    GaviSlots s;
    s.M = a.add(8 * d1 * dz); s.N = a.add(8 * d1 * np); s.o = a.add(8 * d1); s.l1 = a.add(8 * d1); s.u1 = a.add(8 * d1);
    s.A = a.add(8 * d2 * dz); s.B = a.add(8 * d2 * np); s.l2 = a.add(8 * d2); s.u2 = a.add(8 * d2);
    return s;
}
static int stage_gavi(qpn_handle* h, Arena& a, const GaviSlots& s, const qpn_gavi* g, GaviDesc* out) {
    const size_t d1 = g->d1, d2 = g->d2, np = g->np, dz = d1 + d2;
    if (up(h, a.ptr<double>(s.M), g->M, 8 * d1 * dz) || up(h, a.ptr<double>(s.N), g->N, 8 * d1 * np) ||
        up(h, a.ptr<double>(s.o), g->o, 8 * d1) || up(h, a.ptr<double>(s.l1), g->l1, 8 * d1) ||
        up(h, a.ptr<double>(s.u1), g->u1, 8 * d1) || up(h, a.ptr<double>(s.A), g->A, 8 * d2 * dz) ||
        up(h, a.ptr<double>(s.B), g->B, 8 * d2 * np) || up(h, a.ptr<double>(s.l2), g->l2, 8 * d2) ||
        up(h, a.ptr<double>(s.u2), g->u2, 8 * d2)) return -1;
    GaviDesc d;
    d.d1 = g->d1; d.d2 = g->d2; d.np = g->np;
    d.M = a.ptr<double>(s.M); d.N = a.ptr<double>(s.N); d.o = a.ptr<double>(s.o); d.l1 = a.ptr<double>(s.l1);
    d.u1 = a.ptr<double>(s.u1); d.A = a.ptr<double>(s.A); d.B = a.ptr<double>(s.B); d.l2 = a.ptr<double>(s.l2);
    d.u2 = a.ptr<double>(s.u2);
    *out = d;
    return 0;
}
static GaviDesc gavi_dev_desc(const qpn_gavi* g) {
    GaviDesc d;
    d.d1 = g->d1; d.d2 = g->d2; d.np = g->np;
    d.M = g->M; d.N = g->N; d.o = g->o; d.l1 = g->l1; d.u1 = g->u1; d.A = g->A; d.B = g->B; d.l2 = g->l2; d.u2 = g->u2;
    return d;
}

// ---- solve_gavi ------------------------------------------------------------------------------
static int launch_gavi(qpn_handle* h, const GaviDesc& g, int batch, const double* w, const double* z0, int presolve,
                       int max_pivots, double* z, double* zfull, int32_t* st, int32_t* pv, int8_t* basis, cudaStream_t s) {
    if (batch <= 0) return 0;
    const int n = g.d1 + 2 * g.d2;
    if (big_needed(h, n, gavi_smem_bytes(g.d1, g.d2, g.np)))
        return launch_gavi_big(h, g, batch, w, z0, presolve, max_pivots, z, zfull, st, pv, basis, s);
    GaviPlans plans;
    plans.has = 0;
    if (batch >= QPN_PLAN_MIN_BATCH && s == h->stream) {       // plans are built (and synchronised) on the handle's stream
        if (build_plan(h, g, 0, &plans.A, nullptr) || build_plan(h, g, 1, &plans.B, nullptr)) return -1;
        plans.has = 1;
    }
    gavi_workspace_shape(g, plans);
    const size_t smem = tab_smem_bytes_ex(n, (size_t)plans.t_doubles, plans.ldr_max) + gavi_extra_bytes(g.d1, g.d2, g.np);
    if (smem > (size_t)h->max_smem_optin)
        return fail(h, "GAVI with d1=%d d2=%d needs %zu B of shared memory per CTA (limit %d)", g.d1, g.d2, smem, h->max_smem_optin);
    if (max_pivots <= 0) max_pivots = 50 * n + 100;
    QPN_LAUNCH_BUCKETED(gavi_solve_kernel, roundup32(n), batch, smem, s, g, plans, batch, w, z0, presolve, max_pivots, z, zfull, st, pv, basis);
    h->launches++;
    CK(cudaGetLastError());
    return 0;                                                  // the handle's plan buffers are reused by the next call (stream order)
}

extern "C" int qpn_gavi_solve_batched_dev(qpn_handle* h, const qpn_gavi* g, int batch, const double* w, const double* z0,
                                          int presolve, int max_pivots, double* z_out, double* zfull_out,
                                          int32_t* status_out, int32_t* pivots_out, int8_t* basis_out, void* stream) {
    if (!h || !g) return -1;
    CK(cudaSetDevice(h->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
    return launch_gavi(h, gavi_dev_desc(g), batch, w, z0, presolve, max_pivots, z_out, zfull_out, status_out, pivots_out, basis_out, s);
}

extern "C" int qpn_gavi_solve_batched(qpn_handle* h, const qpn_gavi* g, int batch, const double* w, const double* z0,
                                      int presolve, int max_pivots, double* z_out, double* zfull_out,
                                      int32_t* status_out, int32_t* pivots_out, int8_t* basis_out) {
    if (!h || !g) return -1;
    CK(cudaSetDevice(h->device));
    const size_t dz = (size_t)g->d1 + g->d2, n = dz + g->d2, B = batch;
    Arena a(h);
    GaviSlots gs = plan_gavi(a, g);
    int sw = a.add(8 * (size_t)g->np * B), sz0 = a.add(8 * dz * B), sz = a.add(8 * dz * B);
    int szf = zfull_out ? a.add(8 * n * B) : -1, sst = a.add(4 * B), spv = a.add(4 * B), sb = basis_out ? a.add(n * B) : -1;
    if (a.commit()) return -1;
    GaviDesc d;
    if (stage_gavi(h, a, gs, g, &d)) return -1;
    if (up(h, a.ptr<double>(sw), w, 8 * (size_t)g->np * B) || up(h, a.ptr<double>(sz0), z0, 8 * dz * B)) return -1;
    if (launch_gavi(h, d, batch, a.ptr<double>(sw), a.ptr<double>(sz0), presolve, max_pivots, a.ptr<double>(sz),
                    a.ptr<double>(szf), a.ptr<int32_t>(sst), a.ptr<int32_t>(spv), a.ptr<int8_t>(sb), h->stream)) return -1;
    if (down(h, z_out, a.ptr<double>(sz), 8 * dz * B) || down(h, zfull_out, a.ptr<double>(szf), 8 * n * B) ||
        down(h, status_out, a.ptr<int32_t>(sst), 4 * B) || down(h, pivots_out, a.ptr<int32_t>(spv), 4 * B) ||
        down(h, basis_out, a.ptr<int8_t>(sb), n * B)) return -1;
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

// ---- comp_indices ------------------------------------------------------------------------------
extern "C" int qpn_comp_indices_batched(qpn_handle* h, const qpn_gavi* g, int batch, const double* z, const double* w,
                                        double tol, int8_t* mask_out) {
    if (!h || !g) return -1;
    if (batch <= 0) return 0;
    CK(cudaSetDevice(h->device));
    const size_t dz = (size_t)g->d1 + g->d2, B = batch;
    Arena a(h);
    GaviSlots gs = plan_gavi(a, g);
    int sz = a.add(8 * dz * B), sw = a.add(8 * (size_t)g->np * B), sm = a.add(dz * B);
    if (a.commit()) return -1;
    GaviDesc d;
    if (stage_gavi(h, a, gs, g, &d)) return -1;
    if (up(h, a.ptr<double>(sz), z, 8 * dz * B) || up(h, a.ptr<double>(sw), w, 8 * (size_t)g->np * B)) return -1;
    const long long total = (long long)dz * B;
    comp_indices_kernel<<<(unsigned)((total + 255) / 256), 256, 0, h->stream>>>(d, batch, a.ptr<double>(sz), a.ptr<double>(sw), tol, a.ptr<int8_t>(sm));
    h->launches++;
    CK(cudaGetLastError());
    if (down(h, mask_out, a.ptr<int8_t>(sm), dz * B)) return -1;
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

// ---- Base.in -------------------------------------------------------------------------------------
extern "C" int qpn_halfspace_in_batched(qpn_handle* h, int npoly, int d, int mtot, const int32_t* poly_ptr,
                                        const double* A, const double* l, const double* u, const uint8_t* rl,
                                        const uint8_t* ru, int npts, const double* x, double tol, uint8_t* in_out) {
    if (!h) return -1;
    if (npoly <= 0 || npts <= 0) return 0;
    CK(cudaSetDevice(h->device));
    Arena a(h);
    const size_t m = mtot;
    int sp = a.add(4 * (size_t)(npoly + 1)), sA = a.add(8 * m * d), sl = a.add(8 * m), su = a.add(8 * m);
    int srl = rl ? a.add(m) : -1, sru = ru ? a.add(m) : -1, sx = a.add(8 * (size_t)d * npts), so = a.add((size_t)npoly * npts);
    if (a.commit()) return -1;
    if (up(h, a.ptr<int32_t>(sp), poly_ptr, 4 * (size_t)(npoly + 1)) || up(h, a.ptr<double>(sA), A, 8 * m * d) ||
        up(h, a.ptr<double>(sl), l, 8 * m) || up(h, a.ptr<double>(su), u, 8 * m) || up(h, a.ptr<uint8_t>(srl), rl, m) ||
        up(h, a.ptr<uint8_t>(sru), ru, m) || up(h, a.ptr<double>(sx), x, 8 * (size_t)d * npts)) return -1;
    const long long warps = (long long)npoly * npts;
    const int threads = 256;
    const size_t smem_t = 8 * (size_t)d * HS_TP + 4 * (size_t)HS_TP * npoly + 4 * (size_t)mtot;
    if (npts >= 2 * HS_TP && smem_t <= (size_t)h->max_smem_optin) {
        // many points: point tiles x all rows, matrix entries reused across 16 points in registers
        if (smem_t > 48 * 1024) CK(cudaFuncSetAttribute(halfspace_in_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, qpn_smem_cap()));
        int tthreads = roundup32(mtot);                    // one row per thread when the rows fit one pass
        if (tthreads < 64) tthreads = 64;
        if (tthreads > 512) tthreads = 512;
        halfspace_in_tiled_kernel<<<(unsigned)((npts + HS_TP - 1) / HS_TP), tthreads, smem_t, h->stream>>>(
            npoly, d, mtot, a.ptr<int32_t>(sp), a.ptr<double>(sA), a.ptr<double>(sl), a.ptr<double>(su), a.ptr<uint8_t>(srl),
            a.ptr<uint8_t>(sru), npts, a.ptr<double>(sx), tol, a.ptr<uint8_t>(so));
    } else {
        halfspace_in_kernel<<<(unsigned)((warps * 32 + threads - 1) / threads), threads, 0, h->stream>>>(
            npoly, d, mtot, a.ptr<int32_t>(sp), a.ptr<double>(sA), a.ptr<double>(sl), a.ptr<double>(su), a.ptr<uint8_t>(srl),
            a.ptr<uint8_t>(sru), npts, a.ptr<double>(sx), tol, a.ptr<uint8_t>(so));
    }
    h->launches++;
    CK(cudaGetLastError());
    if (down(h, in_out, a.ptr<uint8_t>(so), (size_t)npoly * npts)) return -1;
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

#include "qpn_level_host.inc"
#include "qpn_net_host.inc"
