#!/bin/bash
# runs tools/exp_traverse_noprof.py for every library variant under csrc/variants (GPU box)
# usage: run_variants.sh ["carveout list"]   e.g. "default 40 60"
for lib in yocto-hair_b200/csrc/variants/libv_*.so; do
  n=$(basename $lib .so)
  for c in ${1:-default}; do
    if [ "$c" = default ]; then unset YH_CARVEOUT; else export YH_CARVEOUT=$c; fi
    echo "== $n carveout=$c"; YH_B200_LIB=$PWD/$lib python tools/exp_traverse_noprof.py ${WL:-c3} 16 2>&1 | tail -1
  done
done
