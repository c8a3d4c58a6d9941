#!/bin/bash
# scratch: the standard GPU pass -- parity tests, bench lines, ncu launch list, one full capture of the step kernel
# usage (on the GPU box, from the repo root): bash scratch/gpu_round.sh <tag> [full]
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
python bench.py --steps 300 --warmup 20 > gpurun_out/bench_pred16_$TAG.json 2> gpurun_out/bench_pred16_$TAG.err; echo "bench rc=$?"
python bench.py --workload gt32 --steps 300 --warmup 20 --skip-cpu --skip-e2e > gpurun_out/bench_gt32_$TAG.json 2> gpurun_out/bench_gt32_$TAG.err
python - <<PY
import json
for w in ("pred16","gt32"):
    try:
        j=json.loads(open(f"gpurun_out/bench_{w}_$TAG.json").read().strip().splitlines()[-1])
        r=j["roofline"]; print(w, "ms/step", round(j["ms_per_step"],4), "value", round(j["value"]), "frac", round(r["frac"],3), "phase", {k:(round(v,1) if isinstance(v,float) else None) for k,v in (r.get("phase_us") or {}).items() if not isinstance(v,dict)}, "e2e", (j.get("e2e") or {}).get("value"))
    except Exception as e: print(w, "ERR", e)
PY
if [ "$2" = "full" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 20 --warmup 3 --skip-cpu --skip-e2e > gpurun_out/ncu_l_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 2 -f -o gpurun_out/step_$TAG python bench.py --steps 20 --warmup 3 --skip-cpu --skip-e2e > gpurun_out/ncu_f_$TAG.log 2>&1
ls -la gpurun_out | tail -5
fi
