# ncu --set full of the forward's rank-3 launch (whole GPU)
set -x
mkdir -p gpurun_out
export TOPO_CONCURRENT_RANKS=0
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile-pass --no-graph"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:combine_fwd16 -s 3 -c 1 -f -o gpurun_out/prof_fwd_r3 $CMD > gpurun_out/ncu_fwd.log 2>&1
echo "exit fwd: $?"
