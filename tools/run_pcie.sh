for n in 1 2 4 8; do python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29514 tools/pcie_probe.py 2>/dev/null | grep ranks; done
