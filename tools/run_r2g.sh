out=gpurun_out
rm -f $out/variants_g.log
for l in libode_b200 libvar_s23a libvar_s23b libvar_s23c; do echo "=== $l" >> $out/variants_g.log; for c in c4; do ODE_B200_LIB=$PWD/ode.jl_b200/$l.so python tools/one_config.py $c 3 2>&1 | tail -1 >> $out/variants_g.log; done; ODE_B200_LIB=$PWD/ode.jl_b200/$l.so python tools/bench_configs.py 262144 2>&1 | grep -E "^vdp mu=50|^C4" >> $out/variants_g.log; done
cat $out/variants_g.log
ODE_B200_LIB=$PWD/ode.jl_b200/libvar_s23a.so python -m pytest tests/test_gpu_ensemble.py tests/test_gpu_fuzz.py -m gpu -q -k "ode23s or rober or robertson or fuzz or golden or analytic" 2>&1 | tail -2
