out=gpurun_out
python -m pytest tests/test_gpu_ensemble.py -m gpu -x -q -k "streaming or pinned or chunked or hinit" 2>&1 | tail -3
for g in 0 8 16 32; do echo "=== STREAM grid $g (0 = staged)"; if [ $g = 0 ]; then export ODE_B200_STREAM_INIT=0; else export ODE_B200_STREAM_INIT=1 ODE_B200_STREAM_GRID=$g; fi; python bench.py --workload lorenz --no-cpu-baseline --steps 5 > $out/bench_r2i_g$g.json 2> $out/bench_r2i_g$g.err; tail -2 $out/bench_r2i_g$g.err; python - <<P
import json
d=json.loads(open("$out/bench_r2i_g$g.json").read().strip().splitlines()[-1])
print("value %.4g  ms %.3f   e2e %.4g  ms %.3f  kernels inside e2e %.3f ms launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["kernels_ms_inside_last_call"], d["gpu_launches"]))
P
done
