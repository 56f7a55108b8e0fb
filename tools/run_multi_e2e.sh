# usage: tools/run_multi_e2e.sh N   -- e2e variants of the Lorenz record on N ranks
N=$1; out=gpurun_out
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --workload lorenz --no-cpu-baseline > $out/bench_e2e_${tag}_n$N.json 2> $out/bench_e2e_${tag}_n$N.err; python - <<P
import json
d=json.loads(open("$out/bench_e2e_${tag}_n$N.json").read().strip().splitlines()[-1])
print("%-28s value %.4g ms %.3f | e2e %.4g ms %.3f (%s)" % ("$tag", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["output_path"][:30]))
P
}
run progressive X=1
run copy_phase ODE_B200_PROGRESSIVE=0
run direct ODE_B200_BENCH_DIRECT=1
nvidia-smi topo -m 2>/dev/null | head -14
lscpu | grep -E "NUMA|Socket|^CPU\(s\)" | head
