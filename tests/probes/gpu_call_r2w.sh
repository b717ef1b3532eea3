set -x
mkdir -p gpurun_out/r2w
for s in 1 2 41; do
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_igemm -s $s -c 1 -o gpurun_out/r2w/ncu_igemm_s$s python tests/probes/ncu_step.py slowfast > gpurun_out/r2w/t$s.log 2>&1
echo "rc $s $?"
done
