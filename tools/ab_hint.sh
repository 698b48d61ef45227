# A/B of the L2 eviction hints in the persistent kernels (MATGCN_REC_HINT / MATGCN_REC_BWD_HINT)
B="python bench.py --no-cpu-baseline --no-exact-leg --no-strong-leg --steps 10 --warmup 3"
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d["roofline"]; print(sys.argv[1], "graph ms", round(d["ms_per_step"],3), "eager", round(d["eager"]["ms_per_step"],3), "bwd", round(r["launch_ms"],3), "fwd", round(r["rec_fwd_kernel"]["launch_ms"],3))'
for cfg in "0 0" "1 0" "0 1" "1 1" "0 0" "1 1"; do
  set -- $cfg
  MATGCN_REC_HINT=$1 MATGCN_REC_BWD_HINT=$2 $B 2>/dev/null | python -c "$P" "hint_fwd=$1 hint_bwd=$2"
done
