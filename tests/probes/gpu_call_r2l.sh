set -x
mkdir -p gpurun_out/r2l
timeout 600 python -m pytest tests/test_gpu_token_kernels.py -m gpu -q --timeout 300 -p no:cacheprovider -rf -k "fused_attention" > gpurun_out/r2l/pytest_attn.log 2>&1
echo "rc attn $?"; tail -5 gpurun_out/r2l/pytest_attn.log
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q --timeout 600 -p no:cacheprovider -rf -k "mvit or maskfeat" > gpurun_out/r2l/pytest_models.log 2>&1
echo "rc models $?"; tail -3 gpurun_out/r2l/pytest_models.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-aten-gpu --no-cpu-baseline --legs mvitv2_s,maskfeat_s > gpurun_out/r2l/bench.json 2> gpurun_out/r2l/bench.err
echo "rc bench $?"
SFB_ATTN_FUSED=0 timeout 900 python bench.py --steps 20 --warmup 5 --no-aten-gpu --no-cpu-baseline --legs mvitv2_s > gpurun_out/r2l/bench_unfused.json 2> gpurun_out/r2l/bench_unfused.err
echo "rc bench2 $?"
