#!/bin/bash
# scratch: round-2 collection on one B200 -- parity tests, driver-style bench (both arms), launch list, full ncu captures, timelines
TAG=${1:-r2f}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_$TAG.err
python scratch/brief.py gpurun_out/bench_$TAG.json
python bench.py --steps 300 --warmup 20 --skip-cpu > gpurun_out/bench_k300_$TAG.json 2> gpurun_out/bench_k300_$TAG.err; echo "bench300 rc=$?"
python scratch/brief.py gpurun_out/bench_k300_$TAG.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 20 --warmup 5 --skip-cpu --skip-e2e --skip-configs > gpurun_out/ncu_l_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step_overlap -s 25 -c 2 -f -o gpurun_out/step_pred_$TAG python bench.py --steps 20 --warmup 5 --skip-cpu --skip-e2e --skip-configs > gpurun_out/ncu_f_pred_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step_overlap -s 25 -c 2 -f -o gpurun_out/step_gt32_$TAG python bench.py --workload gt32 --steps 20 --warmup 5 --skip-cpu --skip-e2e --skip-configs > gpurun_out/ncu_f_gt32_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_raster_known -s 25 -c 2 -f -o gpurun_out/raster_known_$TAG python bench.py --steps 20 --warmup 5 --skip-cpu --skip-e2e --only known64 > gpurun_out/ncu_f_known_$TAG.log 2>&1
for w in pred16 gt32 gt1; do NWALK=4412 WARM=4100 REPS=8 IVM_DEBUG_FLAGS=4096 python scratch/pdl_timeline.py $w > gpurun_out/tl_${w}_$TAG.txt 2>&1; python scratch/steplog.py $w > gpurun_out/steplog_${w}_$TAG.txt 2>&1; done
ls -la gpurun_out | grep $TAG
