run() { # label, args...
  label=$1; shift
  for i in 1 2 3 4 5 6; do python bench.py --no-cpu-baseline --steps 50 "$@" > gpurun_out/s3_ab.json 2>/dev/null; python - <<PY
import json
d=json.loads(open("gpurun_out/s3_ab.json").read().strip().splitlines()[-1])
s=d["step_ms"]
print("$label", round(d["value"]), round(s["median_ms"],2), [x for x in s["slowest_steps"] if x[1] > 32], round(d["e2e"]["value"]))
PY
  done
}
run sparse --config sparse
run dense21 --config dense21
