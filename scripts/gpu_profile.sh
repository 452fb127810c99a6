# ncu evidence for the bench command: (1) per-launch device time of one timed step, (2) a full-set
# capture of the kernels named in $NCU_KERNELS.  Each ncu run follows a plain run of the same command.
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-profile-pass"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_SKIP:-1300} -c ${NCU_COUNT:-420} --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "exit launches: $?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNELS:-agg_|combine_}" -s ${NCU_FULL_SKIP:-200} -c ${NCU_FULL_COUNT:-12} \
    -f -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
echo "exit full: $?"
ls -la gpurun_out
