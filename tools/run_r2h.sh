out=gpurun_out
python -m pytest tests/test_gpu_ensemble.py -m gpu -x -q -k "streaming or pinned or chunked or hinit" 2>&1 | tail -5
for s in 0 1; do echo "=== ODE_B200_STREAM_INIT=$s"; ODE_B200_STREAM_INIT=$s python bench.py --workload lorenz --no-cpu-baseline > $out/bench_r2h_s$s.json 2> $out/bench_r2h_s$s.err; python - <<P
import json
d=json.loads(open("$out/bench_r2h_s$s.json").read().strip().splitlines()[-1])
print("value %.4g  ms %.3f   e2e %.4g  ms %.3f  launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["gpu_launches"]))
P
done
python -m pytest tests -m gpu -x -q > $out/tests_r2h.log 2>&1; echo "pytest rc=$?"; tail -3 $out/tests_r2h.log
