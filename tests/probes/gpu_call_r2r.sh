set -x
mkdir -p gpurun_out/r2r
timeout 600 python -m pytest tests/test_gpu_token_kernels.py -m gpu -q --timeout 300 -p no:cacheprovider -rf -k "fused_attention" > gpurun_out/r2r/pytest_attn.log 2>&1
echo "rc attn $?"; tail -15 gpurun_out/r2r/pytest_attn.log
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q --timeout 600 -p no:cacheprovider -rf -k "mvit or maskfeat" > gpurun_out/r2r/pytest_models.log 2>&1
echo "rc models $?"; tail -5 gpurun_out/r2r/pytest_models.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-aten-gpu --no-cpu-baseline --legs mvitv2_s,maskfeat_s > gpurun_out/r2r/bench.json 2> gpurun_out/r2r/bench.err
echo "rc bench $?"; cut -c1-200 gpurun_out/r2r/bench.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2r/mvit_times.csv python tests/probes/ncu_step.py mvit > gpurun_out/r2r/t1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_fwd -s 4 -c 1 -o gpurun_out/r2r/ncu_attn_fwd python tests/probes/ncu_step.py mvit > gpurun_out/r2r/t2.log 2>&1
echo "rc ncu $?"
