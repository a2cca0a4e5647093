#!/bin/bash
lib=$PWD/yocto-hair_b200/csrc/variants/libv_g.so
YH_B200_LIB=$lib timeout 900 python -m pytest tests/test_trace_gpu.py -x -q 2>&1 | tail -3
for g in 1 2 3 4; do for d in 1 2; do
  echo "== groups=$g div=$d"; YH_WAVE_GROUPS=$g YH_GROUP_GRID_DIV=$d YH_B200_LIB=$lib python tools/exp_traverse_noprof.py c3 16 2>&1 | tail -1
done; done
