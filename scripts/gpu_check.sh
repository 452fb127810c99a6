# Runs every GPU parity test file in its own process (a CUDA fault stays contained), then smoke and bench.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
rm -f gpurun_out/parity_report.txt gpurun_out/summary.txt
for f in rectifier gate builder tc sccn stage distance graph_and_scale; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -m gpu -q --timeout 800 > gpurun_out/test_$f.log 2>&1
  echo "exit $f: $?" >> gpurun_out/summary.txt
  tail -3 gpurun_out/test_$f.log
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "exit smoke: $?" >> gpurun_out/summary.txt
timeout 900 python bench.py ${BENCH_ARGS:---steps 5 --warmup 3 --cpu-samples 2} > gpurun_out/bench.log 2>&1; echo "exit bench: $?" >> gpurun_out/summary.txt
tail -1 gpurun_out/bench.log | cut -c1-400
cat gpurun_out/summary.txt
