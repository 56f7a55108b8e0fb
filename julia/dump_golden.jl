# dump_golden.jl -- run the REAL ODE.jl and write attempt-level golden data in the format the
# parity tests consume (tests/golden/*.json).  Needs Julia + ODE.jl; cannot run in the build
# container (SURVEY.md F2).  Usage:  julia --project julia/dump_golden.jl out.json
using ODE, JSON
lorenz(t, y) = [10.0*(y[2]-y[1]); y[1]*(28.0-y[3])-y[2]; y[1]*y[2]-(8/3)*y[3]]
cases = Dict{String,Any}()
for (name, solver) in (("ode23", ODE.ode23), ("ode45", ODE.ode45), ("ode45_fe", ODE.ode45_fe), ("ode78", ODE.ode78))
    t, y = solver(lorenz, [1.0, 1.0, 1.0], [0.0, 2.0])
    cases["lorenz_" * name] = Dict("t" => t, "y" => y)
end
rober(t, y) = (d1 = -0.04*y[1] + 1.0e4*y[2]*y[3]; d3 = 3.0e7*y[2]*y[2]; [d1; -d1-d3; d3])
t, y = ODE.ode23s(rober, [1.0, 0.0, 0.0], [0.0, 1e11]; abstol=1e-8, reltol=1e-8, maxstep=1e10, minstep=1e-7)
cases["rober_ode23s"] = Dict("t" => t, "y" => y)
open(ARGS[1], "w") do io
    JSON.print(io, cases)
end
