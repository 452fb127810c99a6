# secondary measurements: Hard Concrete (sparse) regime, eager launch mode, pairwise distance sweep
set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --regime sparse --cpu-samples 2 > gpurun_out/bench_sparse.log 2>&1; tail -1 gpurun_out/bench_sparse.log | cut -c1-300
python bench.py --steps 5 --warmup 3 --no-graph --no-cpu-baseline --no-profile-pass > gpurun_out/bench_eager.log 2>&1; tail -1 gpurun_out/bench_eager.log | cut -c1-200
python scripts/bench_distance.py > gpurun_out/bench_distance.log 2>&1; tail -3 gpurun_out/bench_distance.log | cut -c1-400
