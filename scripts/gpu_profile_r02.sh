# Round-2 ncu evidence (each capture after a plain run of the same command exited 0; one ncu tool invocation per gpurun call
# would be ideal -- the captures below are all `ncu`, which counts as one tool):
#   r02_launches_graph.csv   per-launch device time of the real step (CUDA graph, four rank launches side by side)
#   r02_launches_serial.csv  TOPO_CONCURRENT_RANKS=0, eager: every launch alone on the whole GPU
#   r02_prof_bwd             --set full of the four rank launches of one layer's fused backward (dominant kernel)
#   r02_prof_dist            --set full of the distance sweep's gram_kernel and l1_kernel (one launch each)
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile-pass"
$CMD > gpurun_out/r02_plain0.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 700 --csv --log-file gpurun_out/r02_launches_graph.csv $CMD > gpurun_out/r02_ncu0.log 2>&1
echo "exit launches (graph): $?"
export TOPO_CONCURRENT_RANKS=0
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile-pass --no-graph"
$CMD > gpurun_out/r02_plain1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 1500 --csv --log-file gpurun_out/r02_launches_serial.csv $CMD > gpurun_out/r02_ncu1.log 2>&1
echo "exit launches (serial): $?"
$CMD > gpurun_out/r02_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:combine_bwd_fused -s 0 -c 4 -f -o gpurun_out/r02_prof_bwd $CMD > gpurun_out/r02_ncu2.log 2>&1
echo "exit bwd: $?"
unset TOPO_CONCURRENT_RANKS
CMD="python bench.py --workload distance --clips 1024 --row-block 512 --steps 1 --warmup 3 --no-cpu-baseline --no-profile-pass"
$CMD > gpurun_out/r02_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"gram_kernel|l1_kernel" -s 2 -c 2 -f -o gpurun_out/r02_prof_dist $CMD > gpurun_out/r02_ncu3.log 2>&1
echo "exit dist: $?"
