# N = 8 check (gpurun --gpus 8): bench.py as the driver launches it.
set -x
O=gpurun_out
nvidia-smi -L > $O/r02an_n8_gpus.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 > $O/r02an_n8_bench.json 2> $O/r02an_n8_bench.err
echo "bench rc=$?"
tail -3 $O/r02an_n8_bench.err
