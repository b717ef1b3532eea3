set -x
mkdir -p gpurun_out/r2e
for f in test_gpu_replay test_gpu_optim test_gpu_drivers; do
  timeout 1500 python -m pytest tests/$f.py -m gpu -q --timeout 900 -p no:cacheprovider -rA > gpurun_out/r2e/pytest_$f.log 2>&1
  echo "rc $f $?"
  tail -3 gpurun_out/r2e/pytest_$f.log
done
timeout 600 python -m pytest tests/test_oracle.py -q -p no:cacheprovider > gpurun_out/r2e/pytest_oracle.log 2>&1
echo "rc oracle $?"
python __graft_entry__.py --smoke > gpurun_out/r2e/smoke.log 2>&1
echo "rc smoke $?"
