set -x
mkdir -p gpurun_out/r2b
for f in test_gpu_kernels test_gpu_token_kernels test_gpu_models test_gpu_replay test_gpu_stochastic test_gpu_optim test_gpu_drivers; do
  timeout 1200 python -m pytest tests/$f.py -m gpu -q --timeout 900 -p no:cacheprovider -rA > gpurun_out/r2b/pytest_$f.log 2>&1
  echo "rc $f $?"
  tail -3 gpurun_out/r2b/pytest_$f.log
done
timeout 600 python -m pytest tests/test_oracle.py -q -p no:cacheprovider > gpurun_out/r2b/pytest_oracle.log 2>&1
echo "rc oracle $?"
SFB_SIMT_SMALLC=1 SFB_SIMT_MAX_MACS=1000000 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "conv" --timeout 600 -p no:cacheprovider > gpurun_out/r2b/pytest_simt.log 2>&1
echo "rc simt $?"
timeout 600 python tests/probes/smallc_probe.py > gpurun_out/r2b/smallc_probe.log 2>&1
echo "rc probe $?"
timeout 600 python tests/probes/dw_probe.py > gpurun_out/r2b/dw_probe.log 2>&1
echo "rc dwprobe $?"
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b/bench.json 2> gpurun_out/r2b/bench.err
echo "rc bench $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2b/bench_ref.json 2> gpurun_out/r2b/bench_ref.err
echo "rc benchref $?"
