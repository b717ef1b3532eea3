set -x
mkdir -p gpurun_out/r2h
timeout 900 python -m pytest tests/test_gpu_token_kernels.py -m gpu -q --timeout 600 -p no:cacheprovider -rf -k "input_pipeline" > gpurun_out/r2h/pytest_pipeline.log 2>&1
echo "rc pipeline $?"; tail -3 gpurun_out/r2h/pytest_pipeline.log
timeout 900 python -m pytest tests/test_gpu_drivers.py -m gpu -q --timeout 900 -p no:cacheprovider -rf -k "precise" > gpurun_out/r2h/pytest_drivers.log 2>&1
echo "rc drivers $?"; tail -3 gpurun_out/r2h/pytest_drivers.log
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2h/bench.json 2> gpurun_out/r2h/bench.err
echo "rc bench $?"
for w in mvit x3d slowfast; do
  timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2h/traffic_$w.csv python tests/probes/ncu_step.py $w > gpurun_out/r2h/traffic_$w.log 2>&1
  echo "rc traffic $w $?"
done
