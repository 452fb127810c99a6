# Round-end ncu evidence for the bench command (each capture after a plain run of the same command exited 0):
#   launches_serial.csv  per-launch device time, TOPO_CONCURRENT_RANKS=0: every launch has the whole GPU, as in bench.py's
#                        profile pass (serialised, cold cache: read the shares) -> profiles/r01_launches_final.md (a)
#   launches_final.csv   the real step (CUDA graph, four rank launches of a layer on SM partitions) -> (b)
#   prof_bwd_final       --set full of the four rank launches of one layer's fused backward (the dominant kernel), full grid
#   prof_fwd_final       --set full of the four rank launches of one layer's forward, full grid
# Tables: scripts/launch_table.py, scripts/ncu_summary.py, scripts/ncu_lines.py, scripts/ncu_stalls.py.
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile-pass"
$CMD > gpurun_out/plain0.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 700 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_launches_final.log 2>&1
echo "exit launches (concurrent): $?"
export TOPO_CONCURRENT_RANKS=0
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile-pass --no-graph"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 1500 --csv --log-file gpurun_out/launches_serial.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "exit launches (serial): $?"
if [ -n "$LISTS_ONLY" ]; then exit 0; fi     # LISTS_ONLY=1: the two launch lists, no --set full captures
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:combine_bwd_fused -s 0 -c 4 -f -o gpurun_out/prof_bwd_final $CMD > gpurun_out/ncu_bwd.log 2>&1
echo "exit bwd: $?"
$CMD > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:combine_fwd16 -s 0 -c 4 -f -o gpurun_out/prof_fwd_final $CMD > gpurun_out/ncu_fwd.log 2>&1
echo "exit fwd: $?"
