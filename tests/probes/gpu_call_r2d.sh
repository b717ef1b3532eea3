set -x
mkdir -p gpurun_out/r2d
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_token_kernels.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r2d/pytest_kernels.log 2>&1
echo "rc kernels $?"
for f in test_gpu_replay test_gpu_optim test_gpu_drivers test_gpu_models test_gpu_stochastic; do
  timeout 1500 python -m pytest tests/$f.py -m gpu -q --timeout 900 -p no:cacheprovider -rA > gpurun_out/r2d/pytest_$f.log 2>&1
  echo "rc $f $?"
  tail -3 gpurun_out/r2d/pytest_$f.log
done
timeout 600 python -m pytest tests/test_oracle.py -q -p no:cacheprovider > gpurun_out/r2d/pytest_oracle.log 2>&1
echo "rc oracle $?"
timeout 600 python tests/probes/dw_probe.py > gpurun_out/r2d/dw_probe.log 2>&1
echo "rc dwprobe $?"
timeout 600 python tests/probes/smallc_probe.py > gpurun_out/r2d/smallc_probe.log 2>&1
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2d/bench.json 2> gpurun_out/r2d/bench.err
echo "rc bench $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dw3_conv -c 1 -o gpurun_out/r2d/ncu_dw3_conv python tests/probes/dw_probe.py > gpurun_out/r2d/ncu_dw3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:bn_bwd_apply -s 20 -c 1 -o gpurun_out/r2d/ncu_bn_bwd_apply python tests/probes/ncu_step.py slowfast > gpurun_out/r2d/ncu_bn1.log 2>&1
echo "rc ncu $?"
