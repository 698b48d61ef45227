#!/bin/bash
# Round-end measurement batch (one B200): parity tests, bench in every mode / workload, reference arm, ncu launch list of the
# bench command and ncu --set full captures of the top kernels.  Outputs under gpurun_out/ with the given tag.
TAG=${1:-final}
O=gpurun_out
python -m pytest tests -m gpu -q > $O/gpu_tests_$TAG.log 2>&1; tail -2 $O/gpu_tests_$TAG.log
python bench.py > $O/bench_${TAG}_bf16.json 2> $O/bench_${TAG}.err; cut -c1-200 $O/bench_${TAG}_bf16.json
python bench.py --mode tf32 --no-cpu-baseline > $O/bench_${TAG}_tf32.json 2>> $O/bench_${TAG}.err
python bench.py --mode exact --no-cpu-baseline > $O/bench_${TAG}_exact.json 2>> $O/bench_${TAG}.err
for w in dc_multi dc_adaptive_only pems07_scale; do
  python bench.py --workload $w --no-cpu-baseline > $O/bench_${TAG}_$w.json 2>> $O/bench_${TAG}.err
done
python bench.py --impl reference --steps 2 --warmup 1 --cpu-budget 40 > $O/bench_${TAG}_reference.json 2>> $O/bench_${TAG}.err
for f in $O/bench_${TAG}_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], round(d.get("value",0),1), d.get("unit"), "ms/step", round(d.get("ms_per_step",0),2), "e2e", round(d.get("e2e",{}).get("value",0),1), "roof", d.get("roofline",{}).get("frac"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
# ncu: launch list of the bench command (after its plain run above exited 0), then the top kernels once each
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/plain_${TAG}.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 2900 -c 1400 --csv --log-file $O/launches_${TAG}_bf16.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_launches_${TAG}.log 2>&1
cap() {  # cap <name> <kernel regex> <skip> <command...>: one ncu --set full capture, kept as text (the .ncu-rep is ~37 MB)
  local name=$1 rx=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o $O/prof_${name}_${TAG} -f "$@" > $O/ncu_${name}_${TAG}.log 2>&1
  ncu -i $O/prof_${name}_${TAG}.ncu-rep --page raw --csv > $O/prof_${name}_${TAG}_raw.csv 2>/dev/null
  ncu -i $O/prof_${name}_${TAG}.ncu-rep --page details > $O/prof_${name}_${TAG}_details.txt 2>/dev/null
  rm -f $O/prof_${name}_${TAG}.ncu-rep
}
python tools/prop_bench.py 1 twin > $O/plain_prop_${TAG}.log 2>&1 && cap prop_twin gemm_tc_kernel 2 python tools/prop_bench.py 1 twin
python tools/trace_step.py bf16 > $O/plain_trace_${TAG}.log 2>&1 && cap resbwd res_bwd_fused 60 python tools/trace_step.py bf16
# (a capture of the fused tail needs --kernel-name-base demangled to select the EpiCandRes instantiation: not done in round 1)

ls -la $O/ | grep ${TAG} | head -40
