#!/bin/bash
# tuning sweep of the K7' mapping (run on the GPU box): warps per CTA (adjacent wavenumbers), profiles per lane, unroll
for v in "4 4 4" "32 2 4" "32 1 4" "24 2 4" "16 2 4" "32 2 2"; do
  set -- $v
  PTT_NVCC_EXTRA="-DPTT_GC_WARPS=$1 -DPTT_GC_PPL=$2 -DPTT_GC_UNROLL=$3" python -m pttools_b200.build --force > /dev/null || exit 1
  echo "warps=$1 ppl=$2 unroll=$3: $(PYTHONPATH=. python tools/gw_time.py 2>&1 | grep 'cells True\|max rel' | tail -2 | tr '\n' ' ')"
done
python -m pttools_b200.build --force > /dev/null
