# Host-link experiment at N ranks (gpurun --gpus N): plain, core-bound, write-combined.
N=${1:-8}
O=gpurun_out
nvidia-smi topo -m > $O/r02_hostlink_topo.log 2>&1
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" >> $O/r02_hostlink_topo.log
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/exp_hostlink.py 2>/dev/null; }
{ run; EXP_BIND=1 run; PB2_HOST_ALLOC_WC=1 run; } > $O/r02_hostlink_n$N.log 2>&1
cat $O/r02_hostlink_n$N.log
