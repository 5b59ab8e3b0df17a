#!/bin/bash
T="tests/test_gpu_fullsize.py -m gpu -q --timeout 400 -k loss_bounds"
for v in "A=1" "WN_BWD6_TMA=1" "WN_WGRAD_SIDE=1" "WN_BWD6=0"; do
  echo "== $v"; env $v timeout 300 python -m pytest $T 2>&1 | tail -1
done
