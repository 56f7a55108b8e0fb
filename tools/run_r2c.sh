out=gpurun_out
python -m pytest tests -m gpu -x -q > $out/tests_r2c.log 2>&1; echo "pytest rc=$?"; tail -3 $out/tests_r2c.log
rm -f $out/variants_s.log
bash tools/variant_run.sh variants_s ode.jl_b200/libvar_s0.so ode.jl_b200/libvar_s1.so ode.jl_b200/libvar_s2.so ode.jl_b200/libode_b200.so > /dev/null 2>&1
cat $out/variants_s.log
for v in bulk4 bulk5; do
  echo "== parity $v"; ODE_B200_LIB=$PWD/ode.jl_b200/libvar_$v.so python -m pytest tests/test_gpu_large.py -m gpu -q 2>&1 | tail -15
done
for lib in libode_b200 libvar_bulk4 libvar_bulk5; do echo "== probe $lib"; ODE_B200_LIB=$PWD/ode.jl_b200/$lib.so python tools/large_probe.py 26 20 | grep "N=2" | tail -1;  ODE_B200_LIB=$PWD/ode.jl_b200/$lib.so python tools/large_probe.py 22 40 | grep "N=2" | tail -1; done
ODE_B200_LIB=$PWD/ode.jl_b200/libvar_bulk4.so ncu --clock-control none --set full --import-source on -k regex:large_attempt_kernel -s 2 -c 1 -f -o $out/prof_large_bulk4_r2 python tools/large_probe.py 26 6 > $out/ncu_large_bulk4_r2.log 2>&1
