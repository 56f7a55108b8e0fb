out=gpurun_out
python -m pytest tests/test_gpu_ensemble.py tests/test_gpu_large.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -3
python tools/bench_configs.py 262144 2>&1 | grep -v "^{"
python tools/one_config.py c2 3 | tail -1
python tools/large_probe.py 26 20 | grep "N=2" | tail -1
python tools/large_probe.py 22 40 | grep "N=2" | tail -1
