# bench_reference.jl — the true reference number for anyone who has Julia (this image has none, so
# this script has NEVER been executed here; bench.py --impl reference times the C++ restatement
# in oracle/ instead).
#
#   julia -t <cores> --project=<env with DataInterpolationsND> baseline/bench_reference.jl [c1|c2|c3|c4|c5]
#
# It builds the same synthetic inputs as bench.py (non-uniform knots cumsum(0.5 .+ rand(n)), uniform
# random u, sorted linspace grids / uniform random points) with the UNMODIFIED DataInterpolationsND.jl
# and times the reference's own eval_grid! / eval_unstructured! on the KernelAbstractions CPU backend
# (one KA work item per output point, statically split over Threads.nthreads()).
using DataInterpolationsND
using Random

knots(n, T) = T.(cumsum(0.5 .+ rand(n)))

function timeit(f; warmup = 1, reps = 3)
    for _ in 1:warmup
        f()
    end
    best = Inf
    for _ in 1:reps
        t = @elapsed f()
        best = min(best, t)
    end
    return best
end

function c2(; n_u = 256, n_g = 512, T = Float32)
    Random.seed!(1234)
    ts = ntuple(_ -> knots(n_u, T), 3)
    dims = ntuple(d -> LinearInterpolationDimension(ts[d]; t_eval = T.(range(ts[d][1], ts[d][end]; length = n_g))), 3)
    u = rand(T, n_u, n_u, n_u)
    itp = NDInterpolation(u, dims)
    out = similar(u, (n_g, n_g, n_g))
    t = timeit(() -> eval_grid!(out, itp))
    return n_g^3 / t
end

function c1(; n = 10^6, T = Float64)
    Random.seed!(1234)
    ts = ntuple(_ -> knots(64, T), 2)
    xs = ntuple(d -> ts[d][1] .+ (ts[d][end] - ts[d][1]) .* rand(T, n), 2)
    dims = ntuple(d -> LinearInterpolationDimension(ts[d]; t_eval = xs[d]), 2)
    itp = NDInterpolation(rand(T, 64, 64, 2), dims)
    out = zeros(T, n, 2)
    return n / timeit(() -> eval_unstructured!(out, itp))
end

function c3(; n = 10^7, T = Float64)
    Random.seed!(1234)
    ts = ntuple(_ -> knots(126, T), 3)
    xs = ntuple(d -> ts[d][1] .+ (ts[d][end] - ts[d][1]) .* rand(T, n), 3)
    dims = ntuple(d -> BSplineInterpolationDimension(ts[d], 3; t_eval = xs[d], max_derivative_order_eval = 1), 3)
    itp = NDInterpolation(rand(T, 128, 128, 128, 3), dims)
    out = zeros(T, n, 3)
    return n / timeit(() -> eval_unstructured!(out, itp; derivative_orders = (0, 0, 0)))
end

function c4(; n_g = 32, T = Float64)      # reduced grid (32^4) of the 128^4 config
    Random.seed!(1234)
    ts = ntuple(_ -> knots(63, T), 4)
    dims = ntuple(d -> BSplineInterpolationDimension(ts[d], 2; t_eval = T.(range(ts[d][1], ts[d][end]; length = n_g))), 4)
    itp = NDInterpolation(rand(T, 64, 64, 64, 64), dims; cache = NURBSWeights(0.5 .+ rand(T, 64, 64, 64, 64)))
    out = zeros(T, n_g, n_g, n_g, n_g)
    return n_g^4 / timeit(() -> eval_grid!(out, itp))
end

function c5(; n = 10^7, T = Float32)
    Random.seed!(1234)
    ts = ntuple(_ -> knots(32, T), 5)
    xs = ntuple(d -> ts[d][1] .+ (ts[d][end] - ts[d][1]) .* rand(T, n), 5)
    dims = ntuple(d -> LinearInterpolationDimension(ts[d]; t_eval = xs[d]), 5)
    itp = NDInterpolation(rand(T, 32, 32, 32, 32, 32, 4), dims)
    out = zeros(T, n, 4)
    return n / timeit(() -> eval_unstructured!(out, itp))
end

# ---- fixtures: pin the oracle to the reference itself ----------------------------------------------------------
#   julia --project=<env> baseline/bench_reference.jl fixtures [tests/golden/julia]
# writes small inputs and the UNMODIFIED reference's outputs (values, derivatives, idx_eval, grid and unstructured,
# every dimension kind, NURBS) as raw little-endian arrays + manifest.json.  tests/test_oracle_golden.py checks the
# oracle against them when the directory exists (it does not in this repository: no Julia where it was built).
function write_fixtures(dir)
    mkpath(dir)
    Random.seed!(20260101)
    entries = String[]
    function dump(name, a)
        open(joinpath(dir, name * ".bin"), "w") do io
            write(io, a)
        end
        push!(entries, """{"name": "$name", "eltype": "$(eltype(a))", "shape": [$(join(size(a), ", "))]}""")
    end
    cases = [
        ("lin2", (:linear, :linear), (0, 0), (7, 6), (2,), [(0, 0), (1, 0), (0, 1), (1, 1)], false),
        ("lin5", (:linear, :linear, :linear, :linear, :linear), (0, 0, 0, 0, 0), (4, 3, 4, 3, 5), (4,), [(0, 0, 0, 0, 0), (0, 0, 1, 0, 0)], false),
        ("const2", (:constant, :constant), (0, 0), (6, 7), (), [(0, 0)], false),
        ("bs3", (:bspline, :bspline, :bspline), (3, 3, 3), (6, 5, 7), (3,), [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)], false),
        ("bs2d", (:bspline, :bspline), (2, 5), (7, 8), (), [(0, 0), (1, 1), (2, 1)], false),
        ("nurbs4", (:bspline, :bspline, :bspline, :bspline), (2, 2, 2, 2), (5, 4, 5, 4), (), [(0, 0, 0, 0)], true),
    ]
    for (name, kinds, degs, nk, outd, orders_list, nurbs) in cases
        N = length(kinds)
        ts = [knots(nk[d], Float64) for d in 1:N]
        n_u = [kinds[d] == :bspline ? nk[d] + degs[d] - 1 : nk[d] for d in 1:N]
        u = rand(n_u..., outd...)
        w = nurbs ? 0.5 .+ rand(n_u...) : nothing
        for (mode, n) in (("unst", 400), ("grid", N <= 2 ? 30 : (N == 3 ? 9 : 5)))
            # inside the span, a few outside, and the knots themselves (interval-index edge cases)
            xs = [vcat(ts[d][1] .+ (ts[d][end] - ts[d][1]) .* (1.2 .* rand(n) .- 0.1), mode == "unst" ? ts[d][1:min(end, 4)] : Float64[]) for d in 1:N]
            mode == "unst" && (m = minimum(length.(xs)); xs = [x[1:m] for x in xs])
            mode == "grid" && (xs = sort.(xs))
            maxd = maximum(maximum.(orders_list))
            dims = ntuple(N) do d
                kinds[d] == :linear ? LinearInterpolationDimension(ts[d]; t_eval = xs[d]) :
                kinds[d] == :constant ? ConstantInterpolationDimension(ts[d]; t_eval = xs[d]) :
                BSplineInterpolationDimension(ts[d], degs[d]; t_eval = xs[d], max_derivative_order_eval = maxd)
            end
            itp = nurbs ? NDInterpolation(u, dims; cache = NURBSWeights(w)) : NDInterpolation(u, dims)
            for d in 1:N
                dump("$(name)_$(mode)_t$(d)", ts[d]); dump("$(name)_$(mode)_x$(d)", xs[d])
                dump("$(name)_$(mode)_idx$(d)", Int64.(dims[d].idx_eval))
            end
            dump("$(name)_$(mode)_u", u)
            nurbs && dump("$(name)_$(mode)_w", w)
            for o in orders_list
                out = mode == "unst" ? eval_unstructured(itp; derivative_orders = o) : eval_grid(itp; derivative_orders = o)
                dump("$(name)_$(mode)_out_" * join(o, ""), out)
            end
        end
    end
    open(joinpath(dir, "manifest.json"), "w") do io
        println(io, "{\"reference\": \"DataInterpolationsND.jl\", \"arrays\": [\n  " * join(entries, ",\n  ") * "\n]}")
    end
    println("wrote $(length(entries)) arrays to $dir")
end

which = isempty(ARGS) ? "c2" : ARGS[1]
if which == "fixtures"
    write_fixtures(length(ARGS) > 1 ? ARGS[2] : joinpath(@__DIR__, "..", "tests", "golden", "julia"))
    exit(0)
end
rate = Dict("c1" => c1, "c2" => c2, "c3" => c3, "c4" => c4, "c5" => c5)[which]()
println("""{"impl": "reference (Julia, KernelAbstractions CPU)", "workload": "$which", "value": $(rate / 1e9), "unit": "Gpoints/s", "threads": $(Threads.nthreads())}""")
