set -x
mkdir -p gpurun_out/r2final2
timeout 1200 python -m pytest tests/test_gpu_drivers.py tests/test_gpu_optim.py -m gpu -q --timeout 900 -p no:cacheprovider -rA -k "two_gpu or allreduce" > gpurun_out/r2final2/pytest_2gpu.log 2>&1
echo "rc pytest $?"; tail -6 gpurun_out/r2final2/pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 20 --warmup 5 --no-aten-gpu --no-cpu-baseline --legs "" > gpurun_out/r2final2/bench_n2.json 2> gpurun_out/r2final2/bench_n2.err
echo "rc bench n2 $?"; cut -c1-300 gpurun_out/r2final2/bench_n2.json
