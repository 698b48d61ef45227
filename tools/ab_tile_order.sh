# A/B of the dense-phase tile order at the census-tract shape (N = 8192, 64 samples per GPU as four slices of 16): full train steps
for v in 0 1; do
  echo "MATGCN_REC_TN_FAST=$v"
  MATGCN_REC_TN_FAST=$v python tools/large_n_probe.py 8192 64 bf16 4 2>&1 | grep "train step"
done
