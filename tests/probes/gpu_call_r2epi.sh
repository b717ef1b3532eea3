set -x
mkdir -p gpurun_out/r2epi
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_igemm -s 2 -c 1 -o gpurun_out/r2epi/ncu_igemm_s2_coalesced python tests/probes/ncu_step.py slowfast > gpurun_out/r2epi/t.log 2>&1
echo "rc $?"
