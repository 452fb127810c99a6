# Round-end ncu evidence for the bench command (after a plain run of the same command):
#   launches_final.csv   per-launch device time of one step; TOPO_CONCURRENT_RANKS=0 so that every launch has the whole
#                        GPU, as in bench.py's profile pass (serialised, cold cache: read the shares)
#   prof_bwd_final       --set full of the four rank launches of one layer's fused backward (the dominant kernel), full grid
#   prof_fwd_final       --set full of the four rank launches of one layer's forward, full grid
set -x
mkdir -p gpurun_out
export TOPO_CONCURRENT_RANKS=0
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile-pass --no-graph"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 700 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "exit launches: $?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:combine_bwd_fused -s 0 -c 4 -f -o gpurun_out/prof_bwd_final $CMD > gpurun_out/ncu_bwd.log 2>&1
echo "exit bwd: $?"
$CMD > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:combine_fwd16 -s 0 -c 4 -f -o gpurun_out/prof_fwd_final $CMD > gpurun_out/ncu_fwd.log 2>&1
echo "exit fwd: $?"
