set -x
mkdir -p gpurun_out/r2i
timeout 900 python -m pytest tests/test_gpu_token_kernels.py tests/test_gpu_kernels.py -m gpu -q --timeout 600 -p no:cacheprovider -rf > gpurun_out/r2i/pytest_kernels.log 2>&1
echo "rc kernels $?"; tail -3 gpurun_out/r2i/pytest_kernels.log
timeout 1500 python -m pytest tests/test_gpu_models.py -m gpu -q --timeout 900 -p no:cacheprovider -rf -k "mvit or maskfeat or x3d" > gpurun_out/r2i/pytest_models.log 2>&1
echo "rc models $?"; tail -3 gpurun_out/r2i/pytest_models.log
timeout 1500 python bench.py --steps 20 --warmup 5 --no-aten-gpu --no-cpu-baseline > gpurun_out/r2i/bench.json 2> gpurun_out/r2i/bench.err
echo "rc bench $?"
timeout 600 python tests/probes/dw_probe.py > gpurun_out/r2i/dw_probe.log 2>&1
