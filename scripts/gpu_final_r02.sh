# End-of-round check of the shipped build: the whole GPU suite the way the driver runs it, smoke, the default bench line, the
# sparse regime, and the reference arm.  Logs under gpurun_out/final_*.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/final_tests.log 2>&1; echo "exit tests: $?"; tail -2 gpurun_out/final_tests.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "exit smoke: $?"; tail -2 gpurun_out/final_smoke.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/final_bench_stage.log 2>&1; echo "exit bench: $?"
timeout 600 python bench.py --regime sparse --no-cpu-baseline > gpurun_out/final_bench_sparse.log 2>&1; echo "exit sparse: $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.log 2>&1; echo "exit reference: $?"
python - <<'PY'
import json
for name in ("stage", "sparse", "reference"):
    try:
        d = json.loads(open(f"gpurun_out/final_bench_{name}.log").read().strip().splitlines()[-1])
        print(name, round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 3), "e2e", d.get("e2e", {}).get("value"),
              "roofline", (d.get("roofline") or {}).get("frac"), "stage", (d.get("roofline_stage") or {}).get("frac"))
    except Exception as e:
        print(name, "unreadable:", e)
PY
