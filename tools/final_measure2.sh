#!/bin/bash
# Shorter round-end batch: benches in every mode / workload + ncu launch list of the bench command + one full capture.
TAG=${1:-final}
O=gpurun_out
python bench.py > $O/bench_${TAG}_bf16.json 2> $O/bench_${TAG}.err
python bench.py --mode tf32 --no-cpu-baseline > $O/bench_${TAG}_tf32.json 2>> $O/bench_${TAG}.err
python bench.py --mode exact --no-cpu-baseline > $O/bench_${TAG}_exact.json 2>> $O/bench_${TAG}.err
for w in dc_multi dc_adaptive_only pems07_scale; do
  python bench.py --workload $w --no-cpu-baseline > $O/bench_${TAG}_$w.json 2>> $O/bench_${TAG}.err
done
for f in $O/bench_${TAG}_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], round(d.get("value",0),1), "ms/step", round(d.get("ms_per_step",0),2), "e2e", round(d.get("e2e",{}).get("value",0),1), "win", round(d.get("e2e_device_windows",{}).get("value",0),1), "roof", d.get("roofline",{}).get("frac"), "traffic", d.get("roofline",{}).get("traffic"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/plain_${TAG}.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 2900 -c 1400 --csv --log-file $O/launches_${TAG}_bf16.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_launches_${TAG}.log 2>&1
python tools/torch_profile_step.py bf16 > $O/torch_prof_${TAG}.log 2>&1
ls -la $O/ | grep ${TAG} | head -20
