out=gpurun_out
python -m pytest tests -m gpu -x -q > $out/tests_r2f.log 2>&1; echo "pytest rc=$?"; tail -5 $out/tests_r2f.log
for d in 0 1; do echo "=== ODE_B200_DRAIN=$d"; ODE_B200_DRAIN=$d python tools/bench_configs.py 262144 2>&1 | grep -v "^{" ; ODE_B200_DRAIN=$d python tools/one_config.py c2 3 2>&1 | tail -1; done > $out/variants_f.log 2>&1
cat $out/variants_f.log
python bench.py --workload lorenz --no-cpu-baseline > $out/bench_r2f_lorenz.json 2> $out/bench_r2f_lorenz.err; tail -c 1500 $out/bench_r2f_lorenz.json
