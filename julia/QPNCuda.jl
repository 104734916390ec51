# QPNCuda.jl -- the reference-side binding of libqpn_cuda (include/qpn_cuda.h).
#
# Drop this file next to src/avi.jl of QuadraticProgramNetworks.jl and `include` it from
# src/QuadraticProgramNetworks.jl.  It cannot be executed in the build container (Julia is not
# installed there); tests/test_gpu_parity.py drives the identical C symbols through ctypes.
module QPNCuda

using SparseArrays, Random
# inside the package these names are in scope; as a standalone file they come from the package
using ..QuadraticProgramNetworks: Poly, PolyUnion, Slice, Linear, vectorize, decision_inds

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


# ------------------------------------------------------------------------------------------------------------------
# layout check: the mirrors above / below must have the library's struct sizes (qpn_abi_struct_sizes)
# ------------------------------------------------------------------------------------------------------------------
function check_abi()
    out = zeros(Int32, 6)
    ccall((:qpn_abi_struct_sizes, LIB), Cint, (Ptr{Int32},), out)
    mine = Int32[sizeof(QpnMatrix), sizeof(QpnGavi), sizeof(QpnNode), sizeof(QpnLevel), sizeof(QpnNetDesc)]
    out[1:5] == mine || error("libqpn_cuda ABI mismatch: library $(out[1:5]), QPNCuda.jl $(mine)")
    true
end

# ------------------------------------------------------------------------------------------------------------------
# page-locking of Julia arrays.  `GC.@preserve` keeps an Array alive but its memory is pageable: the host-pointer entry
# points then go through the driver's staging copies.  `pin!(h, A)` registers the array's memory with CUDA
# (qpn_host_register) so that batches are read / written in place; `unpin!` before the array can be freed or resized.
# ------------------------------------------------------------------------------------------------------------------
pin!(h::Handle, A::Array) = (check(h, ccall((:qpn_host_register, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), h.ptr, A, sizeof(A))); A)
unpin!(h::Handle, A::Array) = (check(h, ccall((:qpn_host_unregister, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), h.ptr, A)); A)

# ------------------------------------------------------------------------------------------------------------------
# the remaining single-purpose entry points of include/qpn_cuda.h
# ------------------------------------------------------------------------------------------------------------------
"""
Batched `check_avi_solution` (src/avi.jl:148-156).  Returns (bad::Vector{Int32}, R) with `bad[b]` the number of
violated conditions of instance b and `R = M z + q`.
"""
function check_avi_batched(h::Handle, M::SparseMatrixCSC{Float64,Int32}, Q::Matrix{Float64}, l::Vector{Float64},
                           u::Vector{Float64}, Z::Matrix{Float64}; tol=1e-6)
    n, B = size(Q)
    bad = Vector{Int32}(undef, B); R = similar(Q)
    GC.@preserve M Q l u Z bad R begin
        m = Ref(QpnMatrix(C_NULL, pointer(M.colptr), pointer(M.rowval), pointer(M.nzval), Int32(nnz(M)), Int32(1), Int32(1)))
        check(h, ccall((:qpn_check_avi_batched, LIB), Cint,
                       (Ptr{Cvoid}, Cint, Cint, Ref{QpnMatrix}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Cdouble,
                        Ptr{Int32}, Ptr{Cdouble}), h.ptr, n, B, m, Q, l, u, 1, Z, tol, bad, R))
    end
    bad, R
end

# dense column-major copies of a GAVI's blocks + the struct that points at them (keep `keep` alive while `g` is used)
function gavi_struct(gavi)
    d1, d2 = length(gavi.l1), length(gavi.l2)
    M, N, A, Bm = Matrix(gavi.M), Matrix(gavi.N), Matrix(gavi.A), Matrix(gavi.B)
    keep = (M, N, A, Bm, gavi.o, gavi.l1, gavi.u1, gavi.l2, gavi.u2)
    g = QpnGavi(d1, d2, size(N, 2), pointer(M), pointer(N), pointer(gavi.o), pointer(gavi.l1), pointer(gavi.u1),
                pointer(A), pointer(Bm), pointer(gavi.l2), pointer(gavi.u2))
    g, keep
end

"""
Batched `comp_indices(gavi, z, w)` (src/avi_solutions.jl:587-612): `Z` is (d1+d2) x B, `W` np x B.  Returns the
(d1+d2) x B matrix of 4-bit masks (bit k-1 set <=> set k, resp. k+4 in the second block, is admissible).
`masks_to_J` turns one column into the reference's `Dict{Int,Set{Int}}`.
"""
function comp_indices_batched(h::Handle, gavi, Z::Matrix{Float64}, W::Matrix{Float64}; tol=1e-2)
    B = size(Z, 2)
    mask = Matrix{Int8}(undef, size(Z, 1), B)
    g, keep = gavi_struct(gavi)
    GC.@preserve keep Z W mask begin
        check(h, ccall((:qpn_comp_indices_batched, LIB), Cint, (Ptr{Cvoid}, Ref{QpnGavi}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Ptr{Int8}),
                       h.ptr, Ref(g), B, Z, W, tol, mask))
    end
    mask
end
function masks_to_J(mask::AbstractVector{Int8}, d1::Integer)
    J = Dict{Int,Set{Int}}()
    for (i, mk) in enumerate(mask)
        mk == 0 && error("comp_indices: index $i belongs to no set")          # the reference's @assert (avi_solutions.jl:584,609)
        off = i <= d1 ? 0 : 4
        J[i] = Set(k + off for k in 1:4 if (mk >> (k - 1)) & 1 == 1)
    end
    J
end

"""
Batched `x in poly` (src/sets.jl:820-825,850-853) for several polys over the same dimension and several points:
`X` is d x npts.  Returns a npoly x npts Bool matrix.
"""
function in_batched(h::Handle, polys::Vector, X::Matrix{Float64}; tol=1e-6)
    d, npts = size(X)
    parts = [vectorize(p) for p in polys]                                       # (A, l, u, rl, ru) per poly, sets.jl:213-221
    ptr = Int32[0; cumsum(Int32[length(t[2]) for t in parts])]
    A = Matrix(reduce(vcat, (t[1] for t in parts))); l = reduce(vcat, (t[2] for t in parts)); u = reduce(vcat, (t[3] for t in parts))
    rl = UInt8[f === (<) for t in parts for f in t[4]]; ru = UInt8[f === (<) for t in parts for f in t[5]]
    out = Matrix{UInt8}(undef, length(polys), npts)
    GC.@preserve ptr A l u rl ru X out begin
        check(h, ccall((:qpn_halfspace_in_batched, LIB), Cint,
                       (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Int32}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{UInt8}, Ptr{UInt8}, Cint,
                        Ptr{Cdouble}, Cdouble, Ptr{UInt8}), h.ptr, length(polys), d, size(A, 1), ptr, A, l, u, rl, ru, npts, X, tol, out))
    end
    out .== 1
end

# one node's view for verify_solution (keep `keep` alive while the struct is used)
function node_struct(qpn, id, extra_polys=Poly[])
    nv = length(qpn.variables)
    d = decision_inds(qpn, id); qp = qpn.qps[id]
    Qd = Matrix(qp.f.Q[d, :]); qd = qp.f.q[d]
    polys = [[qpn.constraints[c].poly for c in qp.constraint_indices]; extra_polys]
    Alu = [vectorize(p) for p in polys]
    A = isempty(Alu) ? zeros(0, nv) : Matrix(reduce(vcat, (t[1] for t in Alu)))
    l = isempty(Alu) ? Float64[] : reduce(vcat, (t[2] for t in Alu)); u = isempty(Alu) ? Float64[] : reduce(vcat, (t[3] for t in Alu))
    d0 = Int32.(d .- 1)
    QpnNode(length(d), nv, length(l), pointer(Qd), pointer(qd), pointer(A), pointer(l), pointer(u), pointer(d0)), (Qd, qd, A, l, u, d0)
end

"""
Batched `verify_solution(qp, id, constraints, dec_inds, x)` (src/qp_processing.jl:57-149) for the columns of `X`
(n_vars x B); `extra_polys` are the child pieces appended to the node's own constraints (qp_processing.jl:186-187).
Returns (solution::BitVector, Lam (m x B), how::Vector{Int32}, active (m x B)).
"""
function verify_solution_batched(h::Handle, qpn, id, X::Matrix{Float64}; extra_polys=Poly[], tol=1e-4)
    B = size(X, 2)
    node, keep = node_struct(qpn, id, extra_polys)
    m = Int(node.m)
    sol = Vector{UInt8}(undef, B); Lam = Matrix{Float64}(undef, m, B); how = Vector{Int32}(undef, B); act = Matrix{Int8}(undef, m, B)
    GC.@preserve keep X sol Lam how act begin
        check(h, ccall((:qpn_verify_solution_batched, LIB), Cint,
                       (Ptr{Cvoid}, Ref{QpnNode}, Cint, Ptr{Cdouble}, Cdouble, Ptr{UInt8}, Ptr{Cdouble}, Ptr{Int32}, Ptr{Int8}),
                       h.ptr, Ref(node), B, X, tol, sol, Lam, how, act))
    end
    sol .== 1, Lam, how, act
end

# ------------------------------------------------------------------------------------------------------------------
# resident levels: a flat Nash game / bottom level uploaded once (qpn_level_upload), batches then move only x
# ------------------------------------------------------------------------------------------------------------------
mutable struct ResidentLevel
    h::Handle
    ptr::Ptr{Cvoid}
    nv::Int
    lam_total::Int
end

function ResidentLevel(h::Handle, qpn, players, gavi, dec::Vector{Int}, par::Vector{Int}, proj::Matrix{Float64})
    nv = length(qpn.variables)
    keep = Any[]; nodes = QpnNode[]
    for id in players
        nd, k = node_struct(qpn, id)
        push!(nodes, nd); push!(keep, k)
    end
    g, gkeep = gavi_struct(gavi)
    dec0, par0 = Int32.(dec .- 1), Int32.(par .- 1)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep gkeep nodes dec0 par0 proj begin
        lv = Ref(QpnLevel(nv, length(nodes), pointer(nodes), g, pointer(dec0), length(dec0), pointer(par0),
                          qpn.options.max_iters, size(proj, 2), pointer(proj)))
        check(h, ccall((:qpn_level_upload, LIB), Cint, (Ptr{Cvoid}, Ref{QpnLevel}, Ref{Ptr{Cvoid}}), h.ptr, lv, out))
    end
    L = ResidentLevel(h, out[], nv, sum(Int(n.m) for n in nodes))
    finalizer(L -> ccall((:qpn_level_release, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), L.h.ptr, L.ptr), L)
    L
end

"Plan shapes of a resident level (qpn_level_info)."
function level_info(L::ResidentLevel)
    out = zeros(Int32, 8)
    check(L.h, ccall((:qpn_level_info, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int32}), L.h.ptr, L.ptr, out))
    (; n = out[1], live_columns = out[2], plan_pivots = out[3], presolve_n = out[4], presolve_live_columns = out[5],
       presolve_plan_pivots = out[6], global_memory_path = out[7] == 1)
end

"""
The iterate loop of `solve_base!` for a resident level (qpn_level_equilibrium_resident): `inits` n_vars x B.
Pass arrays registered with `pin!` to have the kernel read / write them in place.  Returns (X, solved, iters, pivots, Lam).
"""
function solve_level!(L::ResidentLevel, inits::Matrix{Float64};
                      X=similar(inits), solved=Vector{UInt8}(undef, size(inits, 2)), iters=Vector{Int32}(undef, size(inits, 2)),
                      pivots=Vector{Int32}(undef, size(inits, 2)), Lam=Matrix{Float64}(undef, L.lam_total, size(inits, 2)))
    B = size(inits, 2)
    GC.@preserve inits X solved iters pivots Lam begin
        check(L.h, ccall((:qpn_level_equilibrium_resident, LIB), Cint,
                         (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{UInt8}, Ptr{Int32}, Ptr{Int32}, Ptr{Cdouble}),
                         L.h.ptr, L.ptr, B, inits, X, solved, iters, pivots, Lam))
    end
    X, solved .== 1, iters, pivots, Lam
end

# ------------------------------------------------------------------------------------------------------------------
# solve(qpn, inits::Matrix): the batched form north_star adds.  The whole recursion of solve_base!
# (src/algorithm.jl:1-127) for every column of `inits` runs inside qpn_net_solve_batched; Julia flattens the QPNet
# once (qpn_net_desc: what setup(:name) produced, as plain arrays) and gets back the reference's NamedTuples.
# ------------------------------------------------------------------------------------------------------------------
struct QpnNetDesc
    nv::Int32; nplayers::Int32; nlevels::Int32; npolys::Int32
    Q::Ptr{Cdouble}; q::Ptr{Cdouble}
    var_ptr::Ptr{Int32}; var_idx::Ptr{Int32}; con_ptr::Ptr{Int32}; con_idx::Ptr{Int32}; child_ptr::Ptr{Int32}; child_idx::Ptr{Int32}
    level_of::Ptr{Int32}; poly_ptr::Ptr{Int32}
    poly_A::Ptr{Cdouble}; poly_l::Ptr{Cdouble}; poly_u::Ptr{Cdouble}
    max_iters::Int32; num_projections::Int32; exploration_vertices::Int32; gen_solution_map::Int32; check_for_cycling::Int32
    remove_subsets_at::Ptr{UInt8}; proj::Ptr{Cdouble}
end

mutable struct Net
    h::Handle
    ptr::Ptr{Cvoid}
    ids::Vector{Int}           # player ids in the library's 0-based order
    nv::Int
    nlevels::Int
end

csr(lists) = (Int32[0; cumsum(Int32[length(l) for l in lists])], Int32[v for l in lists for v in l])

"""
Upload a QPNet (qpn_net_create).  `rng` draws the cycle-check vectors exactly as `solve` does
(src/requests.jl:21: `MersenneTwister(1)`, src/algorithm.jl:10-12).
"""
function Net(h::Handle, qpn; rng=MersenneTwister(1))
    nv = length(qpn.variables)
    ids = sort(collect(keys(qpn.qps))); cids = sort(collect(keys(qpn.constraints)))
    ppos = Dict(id => k - 1 for (k, id) in enumerate(ids)); cpos = Dict(id => k - 1 for (k, id) in enumerate(cids))
    np = length(ids)
    Q = zeros(nv, nv, np); q = zeros(nv, np)                           # Q symmetric: row- and column-major coincide
    for (k, id) in enumerate(ids)
        Q[:, :, k] = Matrix(qpn.qps[id].f.Q); q[:, k] = qpn.qps[id].f.q
    end
    var_ptr, var_idx = csr([Int32.(qpn.qps[id].var_indices .- 1) for id in ids])
    con_ptr, con_idx = csr([Int32[cpos[c] for c in qpn.qps[id].constraint_indices] for id in ids])
    child_ptr, child_idx = csr([Int32[ppos[j] for j in sort(collect(qpn.network_edges[id]))] for id in ids])
    level_of = zeros(Int32, np)
    for (lv, players) in qpn.network_depth_map, id in players
        level_of[ppos[id] + 1] = lv - 1
    end
    nl = length(qpn.network_depth_map)
    parts = [vectorize(qpn.constraints[c].poly) for c in cids]
    poly_ptr = Int32[0; cumsum(Int32[length(t[2]) for t in parts])]
    PA = Matrix(transpose(Matrix(reduce(vcat, (t[1] for t in parts)))))   # rows x nv, ROW-major = nv x rows column-major
    pl = reduce(vcat, (t[2] for t in parts)); pu = reduce(vcat, (t[3] for t in parts))
    o = qpn.options
    remove = UInt8[(lv in o.levels_to_remove_subsets) ? 1 : 0 for lv in 1:nl]
    nproj = o.check_for_cycling ? o.num_projections : 0
    proj = reduce(hcat, [randn(rng, nv) for _ in 1:nproj]; init=zeros(nv, 0))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve Q q var_ptr var_idx con_ptr con_idx child_ptr child_idx level_of poly_ptr PA pl pu remove proj begin
        d = Ref(QpnNetDesc(nv, np, nl, length(cids), pointer(Q), pointer(q), pointer(var_ptr), pointer(var_idx), pointer(con_ptr),
                           pointer(con_idx), pointer(child_ptr), pointer(child_idx), pointer(level_of), pointer(poly_ptr),
                           pointer(PA), pointer(pl), pointer(pu), o.max_iters, nproj, o.exploration_vertices, o.gen_solution_map,
                           o.check_for_cycling, pointer(remove), pointer(proj)))
        check(h, ccall((:qpn_net_create, LIB), Cint, (Ptr{Cvoid}, Ref{QpnNetDesc}, Ref{Ptr{Cvoid}}), h.ptr, d, out))
    end
    net = Net(h, out[], ids, nv, nl)
    finalizer(n -> ccall((:qpn_net_destroy, LIB), Cint, (Ptr{Cvoid},), n.ptr), net)
    net
end

set_option(net::Net, name::AbstractString, value::Integer) =
    ccall((:qpn_net_set_option, LIB), Cint, (Ptr{Cvoid}, Cstring, Int64), net.ptr, name, value) == 0 || error("unknown net option $name")

const NET_ERRORS = Dict(1 => "Cycling detected (noticed solution iterate returned to a previous value).",
                        2 => "AVI solve error. This might be because one of the qps is unbounded or ill-conditioned.",
                        3 => "Detected disagreement in solution status between qp solution processer and equilibrium solver.",
                        4 => "Can't find solution", 5 => "This shouldn't happen. Solution graph is empty.",
                        6 => "comp_indices assertion", 7 => "Too many solutions to combine.",
                        8 => "Solution graphs were not properly populated.", 9 => "Cycling check requested, but num_projections == 0.")

# a solution-graph piece back as a Poly (sets.jl:121-139): rows of qpn_net_piece_get
function piece(net::Net, id::Integer)
    m = ccall((:qpn_net_piece_rows, LIB), Cint, (Ptr{Cvoid}, Cint), net.ptr, id)
    At = Matrix{Float64}(undef, net.nv, m); l = Vector{Float64}(undef, m); u = similar(l); rl = Vector{UInt8}(undef, m); ru = similar(rl)
    ccall((:qpn_net_piece_get, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{UInt8}, Ptr{UInt8}),
          net.ptr, id, At, l, u, rl, ru)
    Poly(Set(Slice(sparsevec(At[:, i]), l[i], u[i], rl[i] == 1 ? (<) : (≤), ru[i] == 1 ? (<) : (≤)) for i in 1:m))
end

"""
    solve(net::Net, inits::Matrix{Float64}; keep_sol=false)

`solve(qpn, inits::Matrix)`: one result per column of `inits`, with the NamedTuple shapes of src/algorithm.jl:116,125 --
`(; solved=true, x_opt, Sol, identified_request, x_alts)` or `(; solved=false, x_fail, x_opt=nothing)`.
`Sol` (Dict id => PolyUnion) is materialised only with `keep_sol=true`.
"""
function solve(net::Net, inits::Matrix{Float64}; keep_sol=false)
    nv, B = size(inits)
    nv == net.nv || error("inits must be n_vars x B")
    X = similar(inits); solved = Vector{UInt8}(undef, B); iters = Matrix{Int32}(undef, net.nlevels, B); err = Vector{Int32}(undef, B)
    GC.@preserve inits X solved iters err begin
        rc = ccall((:qpn_net_solve_batched, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{UInt8}, Ptr{Int32}, Ptr{Int32}),
                   net.ptr, B, inits, X, solved, iters, err)
        rc == 0 || error(unsafe_string(ccall((:qpn_net_last_error, LIB), Cstring, (Ptr{Cvoid},), net.ptr)))
    end
    map(1:B) do b
        if solved[b] == 1
            Sol = Dict{Int,Any}()
            if keep_sol
                for (k, id) in enumerate(net.ids)
                    n = ccall((:qpn_net_sol_count, LIB), Cint, (Ptr{Cvoid}, Cint, Cint), net.ptr, b - 1, k - 1)
                    n >= 0 && (Sol[id] = PolyUnion([piece(net, ccall((:qpn_net_sol_piece, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cint),
                                                                     net.ptr, b - 1, k - 1, j - 1)) for j in 1:n]))
                end
            end
            (; solved = true, x_opt = X[:, b], Sol, identified_request = Set{Linear}(), x_alts = Vector{Float64}[], level_iters = iters[:, b])
        else
            (; solved = false, x_fail = X[:, b], x_opt = nothing, error = get(NET_ERRORS, err[b] & 0xff, ""), level_iters = iters[:, b])
        end
    end
end

# the method the package exports:  solve(qpn, inits::Matrix)  next to  solve(qpn, x_init)  (src/requests.jl:1-22)
const _nets = IdDict{Any,Net}()
const _handles = Dict{Int,Handle}()
function solve(qpn, inits::Matrix{Float64}; device=0, keep_sol=false)
    h = get!(() -> Handle(device), _handles, device)
    net = get!(() -> Net(h, qpn), _nets, qpn)
    solve(net, inits; keep_sol)
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
