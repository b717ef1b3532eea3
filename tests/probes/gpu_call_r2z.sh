set -x
mkdir -p gpurun_out/r2z
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -p no:cacheprovider -rf -k "conv" > gpurun_out/r2z/pytest_kernels.log 2>&1
rc=$?; echo "rc kernels $rc"; tail -8 gpurun_out/r2z/pytest_kernels.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-aten-gpu --no-cpu-baseline --legs "x3d_m,mvitv2_s" > gpurun_out/r2z/bench.json 2> gpurun_out/r2z/bench.err
echo "rc bench $?"
SFB_EPI_COALESCED=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-aten-gpu --no-cpu-baseline --legs "x3d_m,mvitv2_s" > gpurun_out/r2z/bench_always.json 2> gpurun_out/r2z/bench_always.err
echo "rc bench always $?"
python - <<'PY'
import json
for f in ('gpurun_out/r2z/bench.json','gpurun_out/r2z/bench_always.json'):
    d=json.loads([l for l in open(f) if l.startswith('{')][0])
    print(f, d['value'], d['ms_per_step'], d.get('mvitv2_s',{}).get('value'), d.get('x3d_m',{}).get('value'))
PY
