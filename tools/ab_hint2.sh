B="python bench.py --no-cpu-baseline --no-exact-leg --no-strong-leg --steps 10 --warmup 3"
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d["roofline"]; print(sys.argv[1], "graph ms", round(d["ms_per_step"],3), "eager", round(d["eager"]["ms_per_step"],3), "bwd", round(r["launch_ms"],3), "fwd", round(r["rec_fwd_kernel"]["launch_ms"],3))'
for h in 1 3 0 3 1; do
  MATGCN_REC_HINT=$h $B 2>/dev/null | python -c "$P" "hint_fwd=$h"
done
