# N = 2 validation (gpurun --gpus 2): the NCCL film-reduce test that a 1-GPU box skips, then bench.py at N = 2 (C5 parity field).
set -x
O=gpurun_out
nvidia-smi -L > $O/r02am_n2_gpus.log
python -m pytest tests/test_gpu_multi.py -q -rs > $O/r02am_n2_pytest_multi.log 2>&1; echo "rc=$?" >> $O/r02am_n2_pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r02am_n2_bench.json 2> $O/r02am_n2_bench.err
echo "bench rc=$?"
tail -3 $O/r02am_n2_pytest_multi.log
