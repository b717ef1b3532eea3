set -x
mkdir -p gpurun_out/r2p
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -p no:cacheprovider -rf -k "wgrad" > gpurun_out/r2p/pytest_wgrad.log 2>&1
rc=$?; echo "rc wgrad $rc"; tail -30 gpurun_out/r2p/pytest_wgrad.log
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_replay.py -m gpu -q --timeout 600 -p no:cacheprovider -rf -k "slowfast or c2d or x3d" > gpurun_out/r2p/pytest_models.log 2>&1
echo "rc models $?"; tail -8 gpurun_out/r2p/pytest_models.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-aten-gpu --no-cpu-baseline --legs "x3d_m" > gpurun_out/r2p/bench.json 2> gpurun_out/r2p/bench.err
echo "rc bench $?"; cut -c1-300 gpurun_out/r2p/bench.json
SFB_WGRAD_DIRECT=0 timeout 900 python bench.py --steps 20 --warmup 5 --no-aten-gpu --no-cpu-baseline --legs "" > gpurun_out/r2p/bench_wgd_off.json 2> gpurun_out/r2p/bench_wgd_off.err
echo "rc bench off $?"; cut -c1-300 gpurun_out/r2p/bench_wgd_off.json
timeout 600 python tests/probes/layer_profile.py 8 3 > gpurun_out/r2p/layer_profile.log 2>&1; cp gpurun_out/layer_profile.json gpurun_out/r2p/layer_profile_b8_n3.json
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2p/traffic_slowfast.csv python tests/probes/ncu_step.py slowfast > gpurun_out/r2p/t1.log 2>&1
echo "rc ncu $?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:stem8_fprop -c 1 -o gpurun_out/r2p/ncu_stem8_fprop python tests/probes/ncu_step.py slowfast > gpurun_out/r2p/t2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:stem8_wgrad -c 1 -o gpurun_out/r2p/ncu_stem8_wgrad python tests/probes/ncu_step.py slowfast > gpurun_out/r2p/t3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_wgrad_direct -s 2 -c 1 -o gpurun_out/r2p/ncu_wgrad_direct python tests/probes/ncu_step.py slowfast > gpurun_out/r2p/t4.log 2>&1
echo "rc ncu full $?"
