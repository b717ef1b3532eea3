set -x
mkdir -p gpurun_out/r2a
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for w in slowfast mvit x3d; do
  timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r2a/traffic_$w.csv python tests/probes/ncu_step.py $w > gpurun_out/r2a/traffic_$w.log 2>&1
  echo "rc $w $?"
done
nvidia-smi > gpurun_out/r2a/smi.txt
python -c "import os; print(len(os.sched_getaffinity(0)), os.cpu_count())" > gpurun_out/r2a/cores.txt
