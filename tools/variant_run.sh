#!/bin/bash
# usage: tools/variant_run.sh out_prefix lib1 lib2 ...   -- bench_configs.py once per library build (kernel variants)
out=$1; shift
for lib in "$@"; do
  name=$(basename $lib .so)
  echo "=== $name $VARIANT_NOTE" | tee -a gpurun_out/${out}.log
  ODE_B200_LIB=$PWD/$lib python tools/bench_configs.py 262144 2>&1 | grep -E "^C3|^C4|^lorenz ode45 |^lorenz ode78|^vdp|energy" | tee -a gpurun_out/${out}.log
done
