set -x
mkdir -p gpurun_out/r2o
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -p no:cacheprovider -rf -k "stem8" > gpurun_out/r2o/pytest_stem8.log 2>&1
rc=$?; echo "rc stem8 $rc"; tail -30 gpurun_out/r2o/pytest_stem8.log
if [ $rc -ne 0 ]; then
  # which half fails? fprop-only / wgrad-only visibility comes from the assertion line in the log; stop here
  exit 0
fi
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_replay.py -m gpu -q --timeout 600 -p no:cacheprovider -rf -k "slowfast" > gpurun_out/r2o/pytest_models.log 2>&1
echo "rc models $?"; tail -5 gpurun_out/r2o/pytest_models.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-aten-gpu --no-cpu-baseline --legs "" > gpurun_out/r2o/bench.json 2> gpurun_out/r2o/bench.err
echo "rc bench $?"; cut -c1-400 gpurun_out/r2o/bench.json
SFB_STEM_T8=0 timeout 900 python bench.py --steps 20 --warmup 5 --no-aten-gpu --no-cpu-baseline --legs "" > gpurun_out/r2o/bench_t8off.json 2> gpurun_out/r2o/bench_t8off.err
echo "rc bench off $?"; cut -c1-400 gpurun_out/r2o/bench_t8off.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum -k regex:"stem" --clock-control none --csv --log-file gpurun_out/r2o/stem_times.csv python tests/probes/ncu_step.py slowfast > gpurun_out/r2o/t1.log 2>&1
echo "rc ncu $?"; grep -c stem gpurun_out/r2o/stem_times.csv
