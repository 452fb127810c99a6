# End-of-round-2 ncu evidence for the shipped build (each capture after a plain run of the same command exited 0):
#   r02f_launches_graph.csv  per-launch device time of the real step (CUDA graph, four rank launches side by side)
#   r02f_prof_fwd            --set full of the four rank launches of one layer's combine forward, each alone on the GPU
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile-pass"
$CMD > gpurun_out/r02f_plain0.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 700 --csv --log-file gpurun_out/r02f_launches_graph.csv $CMD > gpurun_out/r02f_ncu0.log 2>&1
echo "exit launches (graph): $?"
export TOPO_CONCURRENT_RANKS=0
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile-pass --no-graph"
$CMD > gpurun_out/r02f_plain1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:combine_fwd16 -s 0 -c 4 -f -o gpurun_out/r02f_prof_fwd $CMD > gpurun_out/r02f_ncu1.log 2>&1
echo "exit fwd: $?"
