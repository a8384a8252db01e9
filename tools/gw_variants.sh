#!/bin/bash
# tuning sweep of the K7' mapping (run on the GPU box): "warps per CTA (adjacent wavenumbers), profiles per lane,
# minimum resident CTAs per SM (register cap)" triples as arguments, e.g. tools/gw_variants.sh "4 6 4" "4 3 4"
for v in "$@"; do
  set -- $v
  PTT_NVCC_EXTRA="-DPTT_GC_WARPS=$1 -DPTT_GC_PPL=$2 -DPTT_GC_MINB=$3" python -m pttools_b200.build --force > /dev/null || exit 1
  echo "warps=$1 ppl=$2 minb=$3: $(PYTHONPATH=. python tools/gw_time.py 2>&1 | tail -1)"
done
python -m pttools_b200.build --force > /dev/null
