# Round 2, second capture of the distance kernels (after l1_kernel moved to the integer pipe and bulk-copy staging):
#   r02b_prof_dist   --set full of one gram_kernel and one l1_kernel launch of a 1024 x 2048 block call
# (640 and 1280 CTAs: several waves on 148 SMs), after the plain run of the same command exited 0.
set -x
mkdir -p gpurun_out
CMD="python bench.py --workload distance --clips 2048 --row-block 1024 --steps 1 --warmup 3 --no-cpu-baseline --no-profile-pass"
$CMD > gpurun_out/r02b_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"gram_kernel|l1_kernel" -s 2 -c 2 -f -o gpurun_out/r02b_prof_dist $CMD > gpurun_out/r02b_ncu.log 2>&1
echo "exit dist: $?"
