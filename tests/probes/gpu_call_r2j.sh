set -x
mkdir -p gpurun_out/r2j
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -p no:cacheprovider -rf -k "dwconv" > gpurun_out/r2j/pytest_kernels.log 2>&1
echo "rc kernels $?"; tail -3 gpurun_out/r2j/pytest_kernels.log
timeout 1500 python -m pytest tests/test_gpu_models.py tests/test_gpu_replay.py -m gpu -q --timeout 900 -p no:cacheprovider -rf -k "x3d" > gpurun_out/r2j/pytest_models.log 2>&1
echo "rc models $?"; tail -3 gpurun_out/r2j/pytest_models.log
timeout 600 python tests/probes/dw_probe.py > gpurun_out/r2j/dw_probe.log 2>&1
timeout 1500 python bench.py --steps 20 --warmup 5 --no-aten-gpu --no-cpu-baseline --legs x3d_m,mvitv2_s > gpurun_out/r2j/bench.json 2> gpurun_out/r2j/bench.err
echo "rc bench $?"
