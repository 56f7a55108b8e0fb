# usage: tools/run_multi.sh N tag   (under gpurun --gpus N)
N=$1; tag=$2; out=gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $out/bench_${tag}_ref_n$N.json 2> $out/bench_${tag}_ref_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > $out/bench_${tag}_n$N.json 2> $out/bench_${tag}_n$N.err; echo "bench rc=$?"
tail -c 600 $out/bench_${tag}_n$N.json; tail -5 $out/bench_${tag}_n$N.err
