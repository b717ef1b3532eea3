set -x
mkdir -p gpurun_out/r2g
nvidia-smi -L > gpurun_out/r2g/gpus.txt
timeout 1200 python -m pytest tests/test_gpu_drivers.py tests/test_gpu_optim.py -m gpu -q --timeout 900 -p no:cacheprovider -rA -k "two_gpu or allreduce" > gpurun_out/r2g/pytest_2gpu.log 2>&1
echo "rc pytest $?"; tail -5 gpurun_out/r2g/pytest_2gpu.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2g/bench_n2.json 2> gpurun_out/r2g/bench_n2.err
echo "rc bench $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2g/bench_ref_n2.json 2> gpurun_out/r2g/bench_ref_n2.err
echo "rc benchref $?"
