set -x
mkdir -p gpurun_out/r2q
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -p no:cacheprovider -rf -k "wgrad_direct" > gpurun_out/r2q/pytest_wgrad.log 2>&1
rc=$?; echo "rc wgrad $rc"; tail -5 gpurun_out/r2q/pytest_wgrad.log
timeout 600 python tests/probes/layer_profile.py 8 3 > gpurun_out/r2q/layer_profile.log 2>&1; cp gpurun_out/layer_profile_b8_n3.json gpurun_out/r2q/layer_profile_direct.json
SFB_WGRAD_DIRECT=0 timeout 600 python tests/probes/layer_profile.py 8 3 > gpurun_out/r2q/layer_profile_off.log 2>&1; cp gpurun_out/layer_profile_b8_n3.json gpurun_out/r2q/layer_profile_tensor.json
timeout 900 python bench.py --steps 20 --warmup 5 --no-aten-gpu --no-cpu-baseline --legs "" > gpurun_out/r2q/bench.json 2> gpurun_out/r2q/bench.err
echo "rc bench $?"; cut -c1-200 gpurun_out/r2q/bench.json
SFB_WGRAD_DIRECT=0 timeout 900 python bench.py --steps 20 --warmup 5 --no-aten-gpu --no-cpu-baseline --legs "" > gpurun_out/r2q/bench_wgd_off.json 2> gpurun_out/r2q/bench_wgd_off.err
echo "rc bench off $?"; cut -c1-200 gpurun_out/r2q/bench_wgd_off.json
