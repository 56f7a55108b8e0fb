out=gpurun_out
python -m pytest tests/test_gpu_ensemble.py -m gpu -x -q 2>&1 | tail -3
for v in "direct ODE_B200_BENCH_DIRECT=1" "staged_progressive ODE_B200_BENCH_DIRECT=0" "staged_phase ODE_B200_BENCH_DIRECT=0 ODE_B200_PROGRESSIVE=0"; do set -- $v; tag=$1; shift; env "$@" python bench.py --workload lorenz --no-cpu-baseline > $out/bench_r2k_$tag.json 2> $out/bench_r2k_$tag.err; tail -2 $out/bench_r2k_$tag.err; python - <<P
import json
d=json.loads(open("$out/bench_r2k_$tag.json").read().strip().splitlines()[-1])
print("%-20s value %.4g ms %.3f | e2e %.4g ms %.3f kernels inside %.3f (%s)" % ("$tag", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["kernels_ms_inside_last_call"], d["e2e"]["output_path"][:30]))
P
done
