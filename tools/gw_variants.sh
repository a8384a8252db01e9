#!/bin/bash
# tuning sweep of the K7' mapping (run on the GPU box): warps per CTA (adjacent wavenumbers), profiles per lane
for v in "4 8" "4 4" "8 4" "2 8" "8 2"; do
  set -- $v
  PTT_NVCC_EXTRA="-DPTT_GC_WARPS=$1 -DPTT_GC_PPL=$2" python -m pttools_b200.build --force > /dev/null || exit 1
  echo "warps=$1 ppl=$2: $(PYTHONPATH=. python tools/gw_time.py 2>&1 | grep 'cells True\|max rel' | tail -2 | tr '\n' ' ')"
done
python -m pttools_b200.build --force > /dev/null
