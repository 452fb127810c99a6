# quick iteration loop: tensor-core + SCCN + stage parity, then a bench line with the per-entry breakdown
set -x
mkdir -p gpurun_out
for f in ${TESTS:-tc sccn stage}; do
  timeout 600 python -m pytest tests/test_gpu_$f.py -m gpu -q -x --timeout 500 > gpurun_out/test_$f.log 2>&1
  echo "exit $f: $?"; tail -2 gpurun_out/test_$f.log | cut -c1-200
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/bench_quick.log 2>&1; echo "exit bench: $?"
python - <<'PY'
import json
l = open('gpurun_out/bench_quick.log').read().strip().splitlines()[-1]
d = json.loads(l)
print('samples/s', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1))
for k, v in (d.get('breakdown') or {}).items():
    if v['share'] > 0.01:
        print(f"  {k:36s} {v['ms_per_step']:7.3f} ms  {v['share']:.3f}")
PY
