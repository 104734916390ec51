# QPNCuda.jl -- the reference-side binding of libqpn_cuda (include/qpn_cuda.h).
#
# Drop this file next to src/avi.jl of QuadraticProgramNetworks.jl and `include` it from
# src/QuadraticProgramNetworks.jl.  It cannot be executed in the build container (Julia is not
# installed there); tests/test_gpu_parity.py drives the identical C symbols through ctypes.
module QPNCuda

using SparseArrays

const LIB = get(ENV, "QPN_CUDA_LIB", "libqpn_cuda.so")

mutable struct Handle
    ptr::Ptr{Cvoid}
end

function Handle(device::Integer=0)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:qpn_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), device, out)
    rc == 0 || error(unsafe_string(ccall((:qpn_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    h = Handle(out[])
    finalizer(h -> ccall((:qpn_destroy, LIB), Cint, (Ptr{Cvoid},), h.ptr), h)
    h
end

check(h::Handle, rc) = rc == 0 || error(unsafe_string(ccall((:qpn_last_error, LIB), Cstring, (Ptr{Cvoid},), h.ptr)))

# mirror of `qpn_matrix` (CSC with PATH's Cint indices, avi.jl:11-12; 1-based as Julia stores it)
struct QpnMatrix
    dense::Ptr{Cdouble}
    colptr::Ptr{Int32}
    rowval::Ptr{Int32}
    nzval::Ptr{Cdouble}
    nnz::Int32
    index_base::Int32
    is_shared::Int32
end

"""
Batched replacement of `solve_avi` (src/avi.jl:63-77): columns of `Q`, `Z0` are instances.
Returns (Z, status::Vector{Int32}, pivots::Vector{Int32}, basis::Matrix{Int8}).
"""
function solve_avi_batched(h::Handle, M::SparseMatrixCSC{Float64,Int32}, Q::Matrix{Float64},
                           l::Vector{Float64}, u::Vector{Float64}, Z0::Matrix{Float64}; max_pivots=0)
    n, B = size(Q)
    Z = similar(Q); status = Vector{Int32}(undef, B); pivots = Vector{Int32}(undef, B); basis = Matrix{Int8}(undef, n, B)
    GC.@preserve M Q l u Z0 Z status pivots basis begin
        m = Ref(QpnMatrix(C_NULL, pointer(M.colptr), pointer(M.rowval), pointer(M.nzval), Int32(nnz(M)), Int32(1), Int32(1)))
        rc = ccall((:qpn_avi_solve_batched, LIB), Cint,
                   (Ptr{Cvoid}, Cint, Cint, Ref{QpnMatrix}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Cint,
                    Ptr{Cdouble}, Ptr{Int32}, Ptr{Int32}, Ptr{Int8}),
                   h.ptr, n, B, m, Q, l, u, 1, Z0, max_pivots, Z, status, pivots, basis)
        check(h, rc)
    end
    Z, status, pivots, basis
end

# drop-in for the single-instance call site:  (; z, status, info) = solve_avi(avi, z0, w)
function solve_avi(h::Handle, avi, z0::Vector{Float64}, w::Vector{Float64})
    q = avi.N * w + avi.o
    Z, st, pv, basis = solve_avi_batched(h, avi.M, reshape(q, :, 1), avi.l, avi.u, reshape(z0, :, 1))
    (; z = Z[:, 1], status = st[1], info = (; pivots = pv[1], basis = basis[:, 1]))     # status uses StatusCode's values
end

# mirror of `qpn_gavi` (struct GAVI, avi.jl:29-39) with dense column-major blocks
struct QpnGavi
    d1::Int32; d2::Int32; np::Int32
    M::Ptr{Cdouble}; N::Ptr{Cdouble}; o::Ptr{Cdouble}; l1::Ptr{Cdouble}; u1::Ptr{Cdouble}
    A::Ptr{Cdouble}; B::Ptr{Cdouble}; l2::Ptr{Cdouble}; u2::Ptr{Cdouble}
end

"""
Batched replacement of `solve_gavi` (src/avi.jl:101-111): presolve projection, lift and AVI solve in
one kernel.  `W` is np x B, `Z0` is (d1+d2) x B.
"""
function solve_gavi_batched(h::Handle, gavi, W::Matrix{Float64}, Z0::Matrix{Float64}; presolve=true, max_pivots=0)
    d1, d2 = length(gavi.l1), length(gavi.l2); B = size(Z0, 2)
    M, N, A, Bm = Matrix(gavi.M), Matrix(gavi.N), Matrix(gavi.A), Matrix(gavi.B)
    Z = Matrix{Float64}(undef, d1 + d2, B); status = Vector{Int32}(undef, B); pivots = Vector{Int32}(undef, B)
    basis = Matrix{Int8}(undef, d1 + 2d2, B)
    GC.@preserve M N A Bm gavi W Z0 Z status pivots basis begin
        g = Ref(QpnGavi(d1, d2, size(N, 2), pointer(M), pointer(N), pointer(gavi.o), pointer(gavi.l1), pointer(gavi.u1),
                        pointer(A), pointer(Bm), pointer(gavi.l2), pointer(gavi.u2)))
        rc = ccall((:qpn_gavi_solve_batched, LIB), Cint,
                   (Ptr{Cvoid}, Ref{QpnGavi}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble},
                    Ptr{Int32}, Ptr{Int32}, Ptr{Int8}),
                   h.ptr, g, B, W, Z0, presolve, max_pivots, Z, C_NULL, status, pivots, basis)
        check(h, rc)
    end
    Z, status, pivots, basis
end

end # module
