set -x
mkdir -p gpurun_out/r2k
timeout 1200 python -m pytest tests/test_gpu_drivers.py -m gpu -q --timeout 900 -p no:cacheprovider -rA -k "two_gpu" > gpurun_out/r2k/pytest_2gpu.log 2>&1
echo "rc pytest $?"; tail -8 gpurun_out/r2k/pytest_2gpu.log
