out=gpurun_out
for g in 148 296 74; do echo "=== stream grid $g"; ODE_B200_STREAM_GRID=$g ODE_B200_DEBUG_STREAM=1 ODE_B200_STREAM_INIT=1 python bench.py --workload lorenz --no-cpu-baseline --steps 3 > $out/bench_r2j.json 2> $out/bench_r2j.err; tail -2 $out/bench_r2j.err; python - <<P
import json
d=json.loads(open("$out/bench_r2j.json").read().strip().splitlines()[-1])
print("value %.4g  ms %.3f   e2e %.4g  ms %.3f  kernels inside e2e %.3f ms" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["kernels_ms_inside_last_call"]))
P
done
