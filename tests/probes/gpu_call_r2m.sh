set -x
mkdir -p gpurun_out/r2m
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum -k regex:"attn_fwd|gemm_batched|softmax_relpos_fwd" --clock-control none --csv --log-file gpurun_out/r2m/attn_times.csv python tests/probes/ncu_step.py mvit > gpurun_out/r2m/t1.log 2>&1
echo "rc t1 $?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_fwd -s 4 -c 1 -o gpurun_out/r2m/ncu_attn_fwd python tests/probes/ncu_step.py mvit > gpurun_out/r2m/t2.log 2>&1
echo "rc t2 $?"
