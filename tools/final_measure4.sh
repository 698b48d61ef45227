#!/bin/bash
# Round-end batch (one B200): parity tests, benches in every mode / workload, reference arm, ncu launch list of the bench command
# (eager issue, so that every launch is a kernel node ncu lists), per-kernel step breakdown, ncu --set full of the two persistent kernels.
TAG=${1:-final}
O=gpurun_out
python -m pytest tests -m gpu -q > $O/gpu_tests_$TAG.log 2>&1; tail -2 $O/gpu_tests_$TAG.log
python __graft_entry__.py --smoke > $O/smoke_$TAG.log 2>&1; tail -4 $O/smoke_$TAG.log
python bench.py > $O/bench_${TAG}_bf16.json 2> $O/bench_${TAG}.err
python bench.py --mode tf32 --no-cpu-baseline --no-exact-leg --no-strong-leg > $O/bench_${TAG}_tf32.json 2>> $O/bench_${TAG}.err
python bench.py --mode exact --no-cpu-baseline --no-exact-leg --no-strong-leg --steps 5 --warmup 2 > $O/bench_${TAG}_exact.json 2>> $O/bench_${TAG}.err
for w in dc_multi dc_adaptive_only pems07_scale; do
  python bench.py --workload $w --no-cpu-baseline --no-exact-leg --no-strong-leg > $O/bench_${TAG}_$w.json 2>> $O/bench_${TAG}.err
done
python bench.py --workload pems07_scale --batch 32 --no-cpu-baseline --no-exact-leg --no-strong-leg > $O/bench_${TAG}_pems07_b32.json 2>> $O/bench_${TAG}.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_${TAG}_reference.json 2>> $O/bench_${TAG}.err
for f in $O/bench_${TAG}_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], round(d.get("value",0),1), "ms/step", round(d.get("ms_per_step",0),2), "e2e", round(d.get("e2e",{}).get("value",0),1), "eager", round((d.get("eager") or {}).get("ms_per_step",0),2), "roof", d.get("roofline",{}).get("frac"), "launches", d.get("gpu_launches"), "clk", d.get("clocks",{}).get("sm_mhz"), d.get("clocks",{}).get("reasons"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
L="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-exact-leg --no-strong-leg --no-graph-leg"
$L > $O/plain_${TAG}.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches_${TAG}_bf16.csv $L > $O/ncu_launches_${TAG}.log 2>&1
python tools/summarize_launches.py $O/launches_${TAG}_bf16.csv > $O/launches_${TAG}_summary.txt 2>&1; head -12 $O/launches_${TAG}_summary.txt
python tools/torch_profile_step.py bf16 > $O/torch_prof_${TAG}.log 2>&1
python tools/rec_once.py > $O/plain_rec_${TAG}.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:rec_ -s 2 -c 2 -o $O/prof_rec_${TAG} -f python tools/rec_once.py > $O/ncu_rec_${TAG}.log 2>&1
ncu -i $O/prof_rec_${TAG}.ncu-rep --page raw --csv > $O/prof_rec_${TAG}_raw.csv 2>/dev/null
ncu -i $O/prof_rec_${TAG}.ncu-rep --page details > $O/prof_rec_${TAG}_details.txt 2>/dev/null
rm -f $O/prof_rec_${TAG}.ncu-rep
ls -la $O | grep ${TAG} | head -40
