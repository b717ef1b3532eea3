set -x
mkdir -p gpurun_out/r2final
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -rf > gpurun_out/r2final/pytest_gpu.log 2>&1
echo "rc pytest $?"; tail -12 gpurun_out/r2final/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2final/smoke.log 2>&1
echo "rc smoke $?"; tail -2 gpurun_out/r2final/smoke.log
timeout 1500 python bench.py > gpurun_out/r2final/bench.json 2> gpurun_out/r2final/bench.err
echo "rc bench $?"; cut -c1-300 gpurun_out/r2final/bench.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2final/bench_reference.json 2> gpurun_out/r2final/bench_reference.err
echo "rc bench ref $?"; cut -c1-300 gpurun_out/r2final/bench_reference.json
for m in slowfast mvit x3d; do
  timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2final/traffic_$m.csv python tests/probes/ncu_step.py $m > gpurun_out/r2final/ncu_$m.log 2>&1
  echo "rc ncu $m $?"
done
