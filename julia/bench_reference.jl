# bench_reference.jl -- the reference's own CPU number for the headline workload, for sites that
# have Julia + ODE.jl (the build container does not): Lorenz / ode45 over an ensemble, threaded
# over trajectories.  Usage: julia -t auto --project julia/bench_reference.jl [n_traj]
using ODE, Random
n = length(ARGS) > 0 ? parse(Int, ARGS[1]) : 1 << 14
rng = MersenneTwister(20261019)
y0 = [[20rand(rng)-10, 20rand(rng)-10, 5+30rand(rng)] for _ in 1:n]
lorenz(t, y) = [10.0*(y[2]-y[1]); y[1]*(28.0-y[3])-y[2]; y[1]*y[2]-(8/3)*y[3]]
steps = zeros(Int, n)
ODE.ode45(lorenz, y0[1], [0.0, 10.0])   # compile
el = @elapsed Threads.@threads for i in 1:n
    t, _ = ODE.ode45(lorenz, y0[i], [0.0, 10.0])
    steps[i] = length(t) - 1            # accepted steps (rejected ones are not observable from outside)
end
println("threads=$(Threads.nthreads()) n=$n accepted_steps=$(sum(steps)) seconds=$el accepted_steps_per_s=$(sum(steps)/el)")
