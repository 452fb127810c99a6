# A/B of environment knobs on the shipped library: SETTINGS is a ';'-separated list of "NAME=VAL NAME=VAL" groups
set -x
mkdir -p gpurun_out
IFS=';' read -ra GROUPS_ <<< "${SETTINGS:-}"
n=0
for grp in "default" "${GROUPS_[@]}"; do
  n=$((n+1))
  for regime in ${REGIMES:-full sparse}; do
    if [ "$grp" = "default" ]; then
      timeout 300 python bench.py --steps ${STEPS:-30} --warmup 5 --no-cpu-baseline --no-profile-pass --regime $regime > gpurun_out/env_${n}_$regime.log 2>&1
    else
      env $grp timeout 300 python bench.py --steps ${STEPS:-30} --warmup 5 --no-cpu-baseline --no-profile-pass --regime $regime > gpurun_out/env_${n}_$regime.log 2>&1
    fi
    python - "$grp" $regime $n <<'PY'
import json, sys
d = json.loads(open(f'gpurun_out/env_{sys.argv[3]}_{sys.argv[2]}.log').read().strip().splitlines()[-1])
print('ENV', sys.argv[1], '|', sys.argv[2], 'samples/s', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 3))
PY
  done
done
