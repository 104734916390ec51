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

# engine options ("force_big", "big_ctas_per_sm") and counters
set_option(h::Handle, name::AbstractString, value::Integer) =
    check(h, ccall((:qpn_set_option, LIB), Cint, (Ptr{Cvoid}, Cstring, Int64), h.ptr, name, value))
launch_count(h::Handle) = ccall((:qpn_launch_count, LIB), Int64, (Ptr{Cvoid},), h.ptr)
big_launch_count(h::Handle) = ccall((:qpn_big_launch_count, LIB), Int64, (Ptr{Cvoid},), h.ptr)

# mirror of `qpn_node` / `qpn_level` (include/qpn_cuda.h)
struct QpnNode
    nd::Int32; nv::Int32; m::Int32
    Qd::Ptr{Cdouble}; qd::Ptr{Cdouble}; A::Ptr{Cdouble}; l::Ptr{Cdouble}; u::Ptr{Cdouble}; dec::Ptr{Int32}
end
struct QpnLevel
    nv::Int32; nplayers::Int32
    players::Ptr{QpnNode}
    gavi::QpnGavi
    dec::Ptr{Int32}; nd_level::Int32
    par::Ptr{Int32}
    max_iters::Int32; num_projections::Int32
    proj::Ptr{Cdouble}
end

"""
`solve(qpn, inits::Matrix)` for a level without children (a flat Nash game, or the bottom level):
the whole iterate loop of `solve_base!` (src/algorithm.jl:13-118) in one call.  `players` are node ids,
`gavi, dec, par` what `solve_qep` assembles for them (src/avi.jl:394-404: `combine_gavis` output, sorted
decision / parameter indices), `proj` the cycle-check vectors (n_vars x num_projections).  `inits` is
n_vars x B.  Returns (X, solved, iters, pivots).
"""
function solve_level_batched(h::Handle, qpn, players, gavi, dec::Vector{Int}, par::Vector{Int}, proj::Matrix{Float64},
                             inits::Matrix{Float64})
    nv, B = size(inits)
    keep = Any[]                                        # every array a pointer is taken of
    nodes = QpnNode[]
    for id in players
        d = decision_inds(qpn, id); qp = qpn.qps[id]
        Qd = Matrix(qp.f.Q[d, :]); qd = qp.f.q[d]
        polys = [qpn.constraints[c].poly for c in qp.constraint_indices]
        Alu = [vectorize(p) for p in polys]
        A = isempty(Alu) ? zeros(0, nv) : Matrix(reduce(vcat, (t[1] for t in Alu)))
        l = isempty(Alu) ? Float64[] : reduce(vcat, (t[2] for t in Alu)); u = isempty(Alu) ? Float64[] : reduce(vcat, (t[3] for t in Alu))
        d0 = Int32.(d .- 1)
        push!(keep, (Qd, qd, A, l, u, d0))
        push!(nodes, QpnNode(length(d), nv, length(l), pointer(Qd), pointer(qd), pointer(A), pointer(l), pointer(u), pointer(d0)))
    end
    d1, d2 = length(gavi.l1), length(gavi.l2)
    M, N, A, Bm = Matrix(gavi.M), Matrix(gavi.N), Matrix(gavi.A), Matrix(gavi.B)
    dec0, par0 = Int32.(dec .- 1), Int32.(par .- 1)
    X = similar(inits); solved = Vector{UInt8}(undef, B); iters = Vector{Int32}(undef, B); pivots = Vector{Int32}(undef, B)
    GC.@preserve keep nodes M N A Bm gavi dec0 par0 proj inits X solved iters pivots begin
        g = QpnGavi(d1, d2, size(N, 2), pointer(M), pointer(N), pointer(gavi.o), pointer(gavi.l1), pointer(gavi.u1),
                    pointer(A), pointer(Bm), pointer(gavi.l2), pointer(gavi.u2))
        lv = Ref(QpnLevel(nv, length(nodes), pointer(nodes), g, pointer(dec0), length(dec0), pointer(par0),
                          qpn.options.max_iters, size(proj, 2), pointer(proj)))
        rc = ccall((:qpn_level_equilibrium_batched, LIB), Cint,
                   (Ptr{Cvoid}, Ref{QpnLevel}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{UInt8}, Ptr{Int32}, Ptr{Int32}, Ptr{Cdouble}),
                   h.ptr, lv, B, inits, X, solved, iters, pivots, C_NULL)
        check(h, rc)
    end
    X, solved .== 1, iters, pivots
end

"""
Write a `QPNet` in the flat-array JSON format of `qpn_b200.load_net` (schema: quadraticprogramnetworks.jl_b200/export.py;
SURVEY.md 8f-4), so that nets built with the symbolic front end (src/programs.jl:147-285) can be solved by
the engine's own host side.  Infinite bounds are written as `null`; indices stay 1-based.
"""
function export_qpnet(qpn, path::AbstractString)
    function csc(M)
        I, J, V = findnz(sparse(M))
        Dict("m" => size(M, 1), "n" => size(M, 2), "I" => I, "J" => J, "V" => V)
    end
    bnd(v) = [isinf(x) ? nothing : x for x in v]
    n = length(qpn.variables)
    qps = Dict(string(id) => Dict("Q" => csc(qp.f.Q), "q" => qp.f.q, "k" => qp.f.k,
                                  "constraint_indices" => qp.constraint_indices, "var_indices" => qp.var_indices) for (id, qp) in qpn.qps)
    cons = Dict{String,Any}()
    for (id, c) in qpn.constraints
        (A, l, u, rl, ru) = vectorize(c.poly)[1:5]
        cons[string(id)] = Dict("A" => csc(A), "l" => bnd(l), "u" => bnd(u), "rl" => Int.(rl .== (<)), "ru" => Int.(ru .== (<)),
                                "group_mapping" => Dict(string(k) => v for (k, v) in c.group_mapping))
    end
    o = qpn.options
    opts = Dict(string(f) => getfield(o, f) for f in fieldnames(typeof(o)) if !(f in (:shared_variable_mode, :levels_to_remove_subsets)))
    lv = o.levels_to_remove_subsets                      # NaturalNumbers() (every level) is written as null
    opts["levels_to_remove_subsets"] = nameof(typeof(lv)) == :NaturalNumbers ? nothing : sort(collect(lv))
    doc = Dict("format" => "qpn-b200/1", "n_vars" => n, "variables" => string.(qpn.variables), "qps" => qps, "constraints" => cons,
               "edges" => [[i, j] for (i, js) in qpn.network_edges for j in js], "options" => opts,
               "default_initialization" => qpn.default_initialization)
    open(path, "w") do io
        write(io, json_string(doc))
    end
end

# minimal JSON writer (the package does not depend on JSON.jl)
json_string(x::Nothing) = "null"
json_string(x::Bool) = x ? "true" : "false"
json_string(x::Integer) = string(x)
json_string(x::AbstractFloat) = isfinite(x) ? repr(Float64(x)) : "null"
json_string(x::AbstractString) = "\"" * escape_string(x) * "\""
json_string(x::Symbol) = json_string(string(x))
json_string(x::Union{AbstractVector,Tuple,AbstractSet}) = "[" * join((json_string(v) for v in x), ",") * "]"
json_string(x::AbstractDict) = "{" * join((json_string(string(k)) * ":" * json_string(v) for (k, v) in x), ",") * "}"

end # module
