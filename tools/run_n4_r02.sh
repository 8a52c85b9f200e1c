# N = 4 check (gpurun --gpus 4): bench.py as the driver launches it.
set -x
O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 10 --warmup 3 > $O/r02au_n4_bench.json 2> $O/r02au_n4_bench.err
echo "bench rc=$?"
