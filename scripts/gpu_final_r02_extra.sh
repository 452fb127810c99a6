# End-of-round numbers of the other BASELINE configurations on the final build (one B200)
set -x
mkdir -p gpurun_out
timeout 400 python bench.py --vertices 28 --batch 16 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/final_bench_n28.log 2>&1; echo "exit n28: $?"
timeout 400 python bench.py --vertices 32 --batch 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/final_bench_n32.log 2>&1; echo "exit n32: $?"
timeout 600 python bench.py --workload distance --no-cpu-baseline > gpurun_out/final_bench_distance.log 2>&1; echo "exit distance: $?"
timeout 600 python bench.py --workload full_step --steps 3 --no-cpu-baseline > gpurun_out/final_bench_full_step.log 2>&1; echo "exit full_step: $?"
python - <<'PY'
import json
for name in ("n28", "n32", "distance", "full_step"):
    try:
        d = json.loads(open(f"gpurun_out/final_bench_{name}.log").read().strip().splitlines()[-1])
        print(name, round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 3), "e2e", (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(name, "unreadable:", e)
PY
