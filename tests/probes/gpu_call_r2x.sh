set -x
mkdir -p gpurun_out/r2x
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -p no:cacheprovider -rf -k "conv or bn or stem" > gpurun_out/r2x/pytest_kernels.log 2>&1
rc=$?; echo "rc kernels $rc"; tail -8 gpurun_out/r2x/pytest_kernels.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q --timeout 600 -p no:cacheprovider -rf > gpurun_out/r2x/pytest_models.log 2>&1
echo "rc models $?"; tail -4 gpurun_out/r2x/pytest_models.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-aten-gpu --no-cpu-baseline --legs "mvitv2_s,x3d_m" > gpurun_out/r2x/bench.json 2> gpurun_out/r2x/bench.err
echo "rc bench $?"
SFB_EPI_COALESCED=0 SFB_ATOMIC_DGRAD=0 timeout 900 python bench.py --steps 20 --warmup 5 --no-aten-gpu --no-cpu-baseline --legs "mvitv2_s,x3d_m" > gpurun_out/r2x/bench_off.json 2> gpurun_out/r2x/bench_off.err
echo "rc bench off $?"
SFB_ATOMIC_DGRAD=0 timeout 900 python bench.py --steps 20 --warmup 5 --no-aten-gpu --no-cpu-baseline --legs "" > gpurun_out/r2x/bench_epi_only.json 2> gpurun_out/r2x/bench_epi_only.err
echo "rc bench epi only $?"
python - <<'PY'
import json
for f in ('gpurun_out/r2x/bench.json','gpurun_out/r2x/bench_off.json','gpurun_out/r2x/bench_epi_only.json'):
    d=json.loads([l for l in open(f) if l.startswith('{')][0])
    print(f, d['value'], d['ms_per_step'], d.get('mvitv2_s',{}).get('value'), d.get('x3d_m',{}).get('value'))
PY
timeout 600 python tests/probes/layer_profile.py 8 3 > gpurun_out/r2x/layer_profile.log 2>&1; cp gpurun_out/layer_profile_b8_n3.json gpurun_out/r2x/layer_profile_coalesced.json
