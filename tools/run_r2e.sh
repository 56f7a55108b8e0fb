out=gpurun_out
python -m pytest tests -m gpu -x -q > $out/tests_r2e.log 2>&1; echo "pytest rc=$?"; tail -3 $out/tests_r2e.log
for l in libode_b200 libvar_cnb0 libvar_cnb; do echo "=== $l"; ODE_B200_LIB=$PWD/ode.jl_b200/$l.so python tools/bench_configs.py 262144 2>&1 | grep -v "^{" ; done > $out/variants_e.log 2>&1
cat $out/variants_e.log
for l in libode_b200 libvar_cnb0; do ODE_B200_LIB=$PWD/ode.jl_b200/$l.so python tools/one_config.py c2 3 2>&1 | tail -1; done
echo "== parity bulk4 (two slots)"; ODE_B200_LIB=$PWD/ode.jl_b200/libvar_bulk4.so python -m pytest tests/test_gpu_large.py -m gpu -q 2>&1 | tail -4
for lib in libode_b200 libvar_bulk4; do echo "== probe $lib"; ODE_B200_LIB=$PWD/ode.jl_b200/$lib.so python tools/large_probe.py 26 20 | grep "N=2" | tail -1;  ODE_B200_LIB=$PWD/ode.jl_b200/$lib.so python tools/large_probe.py 22 40 | grep "N=2" | tail -1; done
ODE_B200_LIB=$PWD/ode.jl_b200/libvar_bulk4.so ncu --clock-control none --set full --import-source on -k regex:large_attempt_kernel -s 2 -c 1 -f -o $out/prof_large_bulk4_r2 python tools/large_probe.py 26 6 > $out/ncu_large_bulk4_r2.log 2>&1
