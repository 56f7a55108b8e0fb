# dump_golden.jl -- run the REAL ODE.jl on the cases of tests/golden/oracle_runs.json and write
# tests/golden/reference_runs.json in the same schema (accepted-step count, number of output points, final time and
# final state as exact hex floats).  Needs Julia + ODE.jl + JSON; it cannot run in the build container (no Julia
# there, SURVEY.md F2).  Once the file exists, tests/test_oracle.py::test_oracle_against_real_julia_if_present
# compares the CPU oracle with it and reports every ulp of difference -- the one place where the ulp-level
# questions (Base `^`, LAPACK `lu`, BLAS `nrm2`) can be settled.
# Usage:  julia --project julia/dump_golden.jl tests/golden/oracle_runs.json tests/golden/reference_runs.json
using ODE, JSON, Printf

hexf(x::Float64) = @sprintf("%a", x)

# the registered right-hand sides, written exactly as csrc/rhs.cuh / oracle/ode_oracle.cpp state them
function rhs(name::String, p)
    name == "linear" && return (t, y) -> p[1] * y
    name == "lorenz" && return (t, y) -> [p[1]*(y[2]-y[1]); y[1]*(p[2]-y[3])-y[2]; y[1]*y[2]-p[3]*y[3]]
    name == "vdp" && return (t, y) -> [y[2]; p[1]*((1.0-y[1]*y[1])*y[2]) - y[1]]
    name == "robertson" && return (t, y) -> (d1 = -p[1]*y[1] + p[3]*y[2]*y[3]; d3 = p[2]*y[2]*y[2]; [d1; -d1-d3; d3])
    name == "kepler" && return (t, y) -> (r2 = y[1]*y[1]+y[2]*y[2]; r3 = r2*sqrt(r2); [y[3]; y[4]; -y[1]/r3; -y[2]/r3])
    if name == "reaction_diffusion"
        return function (t, u)
            n = length(u)
            [p[1]*((u[i == 1 ? n : i-1] - 2.0*u[i]) + u[i == n ? 1 : i+1]) + (p[2]*u[i])*(1.0-u[i]) for i in 1:n]
        end
    end
    error("unknown rhs $name")
end

solvers = Dict("ode1"=>ODE.ode1, "ode2_midpoint"=>ODE.ode2_midpoint, "ode2_heun"=>ODE.ode2_heun, "ode4"=>ODE.ode4,
               "ode21"=>ODE.ode21, "ode23"=>ODE.ode23, "ode45_fe"=>ODE.ode45_fe, "ode45"=>ODE.ode45, "ode78"=>ODE.ode78,
               "ode23s"=>ODE.ode23s, "ode4s_s"=>ODE.ode4s_s, "ode4s_kr"=>ODE.ode4s_kr, "ode4ms"=>ODE.ode4ms)
fixed = Set(["ode1", "ode2_midpoint", "ode2_heun", "ode4", "ode4s_s", "ode4s_kr", "ode4ms"])

out = Any[]
for entry in JSON.parsefile(ARGS[1])
    c = entry["case"]
    haskey(c["kw"], "jacobian") && continue       # the reference never uses a user Jacobian (DESIGN.md section 2)
    p = c["params"] === nothing ? Float64[] : Float64.(c["params"])
    F = rhs(c["rhs"], p)
    y0 = length(c["y0"]) == 1 ? Float64(c["y0"][1]) : Float64.(c["y0"])
    ts = Float64.(c["tspan"])
    kw = c["method"] in fixed ? Dict{Symbol,Any}() : Dict{Symbol,Any}(Symbol(k) => v for (k, v) in c["kw"])
    t, y = solvers[c["method"]](F, y0, ts; kw...)
    yl = y[end] isa AbstractVector ? y[end] : [y[end]]
    push!(out, Dict("case" => c, "result" => Dict("n_out" => length(t), "t_final" => hexf(Float64(t[end])),
                                                  "y_final" => [hexf(Float64(v)) for v in yl])))
end
open(ARGS[2], "w") do io
    JSON.print(io, out, 1)
end
println("wrote $(length(out)) runs to $(ARGS[2])")
