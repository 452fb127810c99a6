# A/B of library variants built into variants/libtopo_b200_<name>.so (same sources, different -D switches): one bench line each.
# VARIANTS="v0 v3" TESTS="tc sccn" bash scripts/gpu_ab.sh
set -x
mkdir -p gpurun_out
LIB=topo_audio_autoencoder_b200/libtopo_b200.so
cp $LIB /tmp/lib_keep.so
for f in ${TESTS:-}; do
  timeout 600 python -m pytest tests/test_gpu_$f.py -m gpu -q -x --timeout 500 > gpurun_out/ab_test_$f.log 2>&1
  echo "exit $f: $?"; tail -2 gpurun_out/ab_test_$f.log | cut -c1-200
done
for v in ${VARIANTS:-v0 v3}; do
  cp variants/libtopo_b200_$v.so $LIB
  for regime in ${REGIMES:-full}; do
    timeout 300 python bench.py --steps ${STEPS:-20} --warmup 5 --no-cpu-baseline --regime $regime ${BENCH_ARGS:-} > gpurun_out/ab_${v}_$regime.log 2>&1; echo "exit bench $v $regime: $?"
    python - $v $regime <<'PY'
import json, sys
l = open(f'gpurun_out/ab_{sys.argv[1]}_{sys.argv[2]}.log').read().strip().splitlines()[-1]
d = json.loads(l)
print('AB', sys.argv[1], sys.argv[2], 'samples/s', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1))
for k, v in (d.get('breakdown') or {}).items():
    if v['share'] > 0.05:
        print(f"   {k:36s} {v['ms_per_step']:7.3f} ms  {v['share']:.3f}")
PY
  done
done
cp /tmp/lib_keep.so $LIB
