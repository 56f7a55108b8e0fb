out=gpurun_out
python -m pytest tests/test_gpu_large.py tests/test_gpu_user_rhs.py -m gpu -x -q 2>&1 | tail -3
for n in 65536 1048576 8388608 67108864; do python bench.py --workload large --large-n $n --no-cpu-baseline --steps 5 > $out/bench_r2n_$n.json 2> $out/bench_r2n_$n.err; python - <<P
import json
d=json.loads(open("$out/bench_r2n_$n.json").read().strip().splitlines()[-1])
print("N=%9d  ms/attempt %.4f  attempts %d  value %.5g  e2e %.5g" % ($n, d["config"]["ms_per_attempt"], d["config"]["attempts_per_step"], d["value"], d["e2e"]["value"]))
P
done
