set -x
python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
python bench.py --no-cpu-baseline --grad-type analytic > gpurun_out/r02f_bench_analytic.json 2> gpurun_out/r02f_bench_analytic.err
python bench.py --no-cpu-baseline --config dense21 > gpurun_out/r02f_bench_dense21.json 2> gpurun_out/r02f_bench_dense21.err
IA_TABLE_FP16=1 python bench.py --no-cpu-baseline --config dense21 > gpurun_out/r02f_bench_dense21_fp16.json 2> gpurun_out/r02f_bench_dense21_fp16.err
python bench.py --no-cpu-baseline --config wreflection --steps 30 > gpurun_out/r02f_bench_wreflection.json 2> gpurun_out/r02f_bench_wreflection.err
for r in 256 1024 32768; do python bench.py --no-cpu-baseline --steps 30 --rays $r > gpurun_out/r02f_bench_rays$r.json 2> gpurun_out/r02f_bench_rays$r.err; done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_ncu_bench.log 2>&1
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r02f_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], round(d.get("value",0)), d.get("ms_per_step"), (d.get("step_ms") or {}).get("median_ms"), round((d.get("e2e") or {}).get("value",0)), (d.get("roofline") or {}).get("frac"))
    except Exception as e: print(f, "ERR", e)
PY
