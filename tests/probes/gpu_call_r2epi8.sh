set -x
mkdir -p gpurun_out/r2epi8
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -p no:cacheprovider -rf -k "conv or stem" > gpurun_out/r2epi8/pytest_kernels.log 2>&1
rc=$?; echo "rc kernels $rc"; tail -6 gpurun_out/r2epi8/pytest_kernels.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q --timeout 300 -p no:cacheprovider -rf > gpurun_out/r2epi8/pytest_models.log 2>&1
rc=$?; echo "rc models $rc"; tail -4 gpurun_out/r2epi8/pytest_models.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 600 python bench.py --steps 20 --warmup 5 --no-aten-gpu --no-cpu-baseline --legs "x3d_m,mvitv2_s" > gpurun_out/r2epi8/bench.json 2> gpurun_out/r2epi8/bench.err
echo "rc bench $?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2epi8/bench.json') if l.startswith('{')][0])
print('RESULT', d['value'], d['ms_per_step'], d['e2e']['value'], d['mvitv2_s']['value'], d['x3d_m']['value'])
PY
