set -x
mkdir -p gpurun_out/r2f
timeout 900 python -m pytest tests/test_gpu_token_kernels.py tests/test_gpu_kernels.py -m gpu -q --timeout 600 -p no:cacheprovider -rf > gpurun_out/r2f/pytest_kernels.log 2>&1
echo "rc kernels $?"; tail -3 gpurun_out/r2f/pytest_kernels.log
timeout 1500 python -m pytest tests/test_gpu_models.py -m gpu -q --timeout 900 -p no:cacheprovider -rf -k "mvit or maskfeat" > gpurun_out/r2f/pytest_models_mvit.log 2>&1
echo "rc models $?"; tail -3 gpurun_out/r2f/pytest_models_mvit.log
timeout 900 python -m pytest tests/test_gpu_drivers.py -m gpu -q --timeout 900 -p no:cacheprovider -rf -k "precise or mvit" > gpurun_out/r2f/pytest_drivers.log 2>&1
echo "rc drivers $?"; tail -3 gpurun_out/r2f/pytest_drivers.log
timeout 1500 python bench.py --steps 10 --warmup 3 --no-aten-gpu --no-cpu-baseline --legs mvitv2_s,maskfeat_s > gpurun_out/r2f/bench.json 2> gpurun_out/r2f/bench.err
echo "rc bench $?"
SFB_DWPOOL_RING=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-aten-gpu --no-cpu-baseline --legs mvitv2_s > gpurun_out/r2f/bench_noring.json 2> gpurun_out/r2f/bench_noring.err
echo "rc bench2 $?"
