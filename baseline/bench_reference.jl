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

which = isempty(ARGS) ? "c2" : ARGS[1]
rate = Dict("c1" => c1, "c2" => c2, "c3" => c3, "c4" => c4, "c5" => c5)[which]()
println("""{"impl": "reference (Julia, KernelAbstractions CPU)", "workload": "$which", "value": $(rate / 1e9), "unit": "Gpoints/s", "threads": $(Threads.nthreads())}""")
