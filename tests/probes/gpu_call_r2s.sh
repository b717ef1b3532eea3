set -x
mkdir -p gpurun_out/r2s
timeout 600 python -m pytest tests/test_gpu_token_kernels.py -m gpu -q --timeout 300 -p no:cacheprovider -rf -k "fused_attention" > gpurun_out/r2s/pytest_attn.log 2>&1
rc=$?; echo "rc attn $rc"; tail -25 gpurun_out/r2s/pytest_attn.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_replay.py tests/test_gpu_stochastic.py -m gpu -q --timeout 600 -p no:cacheprovider -rf -k "mvit or maskfeat or drop" > gpurun_out/r2s/pytest_models.log 2>&1
echo "rc models $?"; tail -8 gpurun_out/r2s/pytest_models.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-aten-gpu --no-cpu-baseline --legs mvitv2_s,maskfeat_s > gpurun_out/r2s/bench.json 2> gpurun_out/r2s/bench.err
echo "rc bench $?"
SFB_ATTN_FUSED_BWD=0 timeout 900 python bench.py --steps 20 --warmup 5 --no-aten-gpu --no-cpu-baseline --legs mvitv2_s > gpurun_out/r2s/bench_bwd_off.json 2> gpurun_out/r2s/bench_bwd_off.err
echo "rc bench off $?"
python - <<'PY'
import json
for f in ('gpurun_out/r2s/bench.json','gpurun_out/r2s/bench_bwd_off.json'):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line)
            for k in ('mvitv2_s','maskfeat_s'):
                if d.get(k): print(f,k,d[k]['value'],d[k]['ms_per_step'])
PY
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum -k regex:"attn_" --clock-control none --csv --log-file gpurun_out/r2s/attn_times.csv python tests/probes/ncu_step.py mvit > gpurun_out/r2s/t1.log 2>&1
echo "rc ncu $?"
