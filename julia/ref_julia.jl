# ref_julia.jl -- produce TRUE reference goldens and timings where Julia exists (SURVEY.md 8c, 8d).
#
# Neither Julia nor PATH / OSQP is in the build image, so no fixture in this repository comes from the
# reference itself ("parity unpinned", DESIGN.md 3).  On a machine with Julia >= 1.8, the package
# installed (`] add QuadraticProgramNetworks`) and a PATH licence in ENV["PATH_LICENSE_STRING"], run
#
#     julia -t auto julia/ref_julia.jl out_dir [batch]
#
# It writes, per example, the model in the flat-array format of `qpn_b200.load_net`
# (julia/QPNCuda.jl: export_qpnet -- so the problem data is the reference's own, seed for seed), the inits,
# the reference's results, and its wall-clock equilibria/sec with one `deepcopy(qpn)` per thread (QPNet is
# not re-entrant: iterate_cache, SURVEY F6).  `python scripts/compare_with_julia_goldens.py out_dir` then
# solves the same nets from the same inits on the B200 engine and reports parity.
using QuadraticProgramNetworks, Random, Printf
include(joinpath(@__DIR__, "QPNCuda.jl"))

out_dir = length(ARGS) >= 1 ? ARGS[1] : "julia_goldens"
batch = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 256
mkpath(out_dir)

js = QPNCuda.json_string

function run_batch(name::Symbol, qpn, inits::Vector{Vector{Float64}})
    n = length(inits)
    results = Vector{Any}(undef, n)
    nt = Threads.nthreads()
    nets = [deepcopy(qpn) for _ in 1:nt]
    solve(nets[1], inits[1])                                   # compile
    t0 = time()
    Threads.@threads :static for k in 1:n
        q = nets[Threads.threadid()]
        results[k] = try
            r = solve(q, inits[k])
            (; solved = r.solved, x = r.solved ? r.x_opt : r.x_fail)
        catch err
            (; solved = false, x = inits[k])
        end
    end
    dt = time() - t0
    @printf("%s: %d solves in %.3f s on %d threads = %.1f equilibria/s, solved %d\n", name, n, dt, nt, n / dt,
            count(r -> r.solved, results))
    QPNCuda.export_qpnet(qpn, joinpath(out_dir, "$(name)_model.json"))
    open(joinpath(out_dir, "$(name)_goldens.json"), "w") do io
        write(io, js(Dict("example" => string(name), "threads" => nt, "seconds" => dt, "equilibria_per_s" => n / dt,
                          "inits" => inits, "solved" => [r.solved for r in results], "x" => [r.x for r in results])))
    end
end

rng = MersenneTwister(0xB200)

# config 1: simple_bilevel, the 8 known answers of test/simple_bilevel.jl:4-11 (+ random w)
qpn = setup(:simple_bilevel)
ws = [[0.0, 1.0], [1.0, 0.0], [1.0, 1.0], [-1.0, 1.0], [1.0, -1.0], [-1.0, -1.0], [2.0, 0.5], [0.5, 2.0]]
inits = [[w; 0.0; 0.0] for w in ws]
append!(inits, [[2 .* randn(rng, 2); 0.0; 0.0] for _ in 1:max(batch - 8, 0)])
run_batch(:simple_bilevel, qpn, inits)

# config 2: four_player_matrix_game as a flat Nash game, inits ~ U(-5, 5)^8
qpn = setup(:four_player_matrix_game; edge_list = [])
run_batch(:four_player_matrix_game, qpn, [10 .* rand(rng, 8) .- 5 for _ in 1:batch])

# config 3: robust_avoid_simple, perturbed positions (parameters) and random controls
qpn = setup(:robust_avoid_simple)
x0 = qpn.default_initialization
inits = map(1:batch) do _
    x = copy(x0)
    x[1:6] .+= 0.5 .* randn(rng, 6)
    x[7:12] .= 2 .* rand(rng, 6) .- 1
    x
end
run_batch(:robust_avoid_simple, qpn, inits)
