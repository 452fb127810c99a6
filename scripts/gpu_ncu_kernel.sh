# full-set ncu capture of one kernel of the bench step (after a plain run of the same command)
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile-pass"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${2:-102} -c ${3:-2} -f -o gpurun_out/${4:-prof_kernel} $CMD > gpurun_out/ncu_kernel.log 2>&1
echo "exit ncu: $?"; tail -3 gpurun_out/ncu_kernel.log
