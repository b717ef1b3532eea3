set -x
mkdir -p gpurun_out/r2v
timeout 1800 python -m pytest tests -x -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/r2v/pytest_gpu.log 2>&1
echo "rc pytest $?"; tail -6 gpurun_out/r2v/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-aten-gpu --no-cpu-baseline --legs "x3d_m" > gpurun_out/r2v/bench.json 2> gpurun_out/r2v/bench.err
echo "rc bench $?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2v/bench.json') if l.startswith('{')][0])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['x3d_m']['value'], d['x3d_m']['ms_per_step'])
PY
