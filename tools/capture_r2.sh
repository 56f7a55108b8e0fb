#!/bin/bash
# Round-2 evidence run on one B200 (under gpurun): tests, bench lines, launch lists, one ncu --set full capture per
# headline kernel.  usage: tools/capture_r2.sh tag [skip_tests]
tag=${1:-r2}; out=gpurun_out
mkdir -p $out
if [ -z "$2" ]; then
  python -m pytest tests -m gpu -x -q > $out/tests_$tag.log 2>&1; echo "pytest rc=$?" >> $out/tests_$tag.log; tail -3 $out/tests_$tag.log
fi
python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_${tag}_ref.json 2> $out/bench_${tag}_ref.err
python bench.py > $out/bench_${tag}_n1.json 2> $out/bench_${tag}_n1.err; echo "bench rc=$?"
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $out/launches_lorenz_$tag.csv python bench.py --workload lorenz --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_ll_lorenz_$tag.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 1500 --csv --log-file $out/launches_large_$tag.csv python bench.py --workload large --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_ll_large_$tag.log 2>&1
FULL="$NCU --set full --import-source on"
$FULL -k regex:ens_adapt_kernel -s 4 -c 1 -f -o $out/prof_lorenz_$tag python bench.py --workload lorenz --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_lorenz_$tag.log 2>&1
$FULL -k regex:ens_adapt_kernel -s 2 -c 1 -f -o $out/prof_c3_$tag python tools/config_probe.py c3 > $out/ncu_c3_$tag.log 2>&1
$FULL -k regex:ens_ode23s_kernel -s 1 -c 1 -f -o $out/prof_c4_$tag python tools/config_probe.py c4 > $out/ncu_c4_$tag.log 2>&1
$FULL -k regex:large_attempt_kernel -s 2 -c 1 -f -o $out/prof_large_$tag python tools/large_probe.py 26 6 > $out/ncu_large_$tag.log 2>&1
$FULL -k regex:fp64_peak_kernel -s 1 -c 1 -f -o $out/prof_peak_dfma_$tag python tools/peak_probe.py > $out/ncu_peak_$tag.log 2>&1
$FULL -k regex:fp64_peak_kernel -s 6 -c 1 -f -o $out/prof_peak_mix_$tag python tools/peak_probe.py >> $out/ncu_peak_$tag.log 2>&1
ls -la $out/*$tag*
