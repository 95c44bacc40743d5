#!/bin/bash
# Builds the library with -DSB_BOUNDS_CHECK next to the real one (here, nvcc cross-compiles) and, on the GPU box, runs the
# GPU parity tests and one bench step of every workload with it: any data-dependent load of the keypoint kernels outside
# the layer it reads traps with a message.    tools/bounds_check.sh build | run
L=$(cd $(dirname $0)/.. && pwd)/sift_features_b200
if [ "$1" = build ]; then
  SB200_NVCC_EXTRA="-DSB_BOUNDS_CHECK" SB200_BUILD_OUT=$L/libsift_b200_check.so python -m sift_features_b200.build --force
else
  export SB200_LIB=$L/libsift_b200_check.so
  python -m pytest tests -m gpu -x -q 2>&1 | tail -4
  for w in 1080p vga; do python bench.py --workload $w --steps 2 --warmup 3 --no-cpu --no-extra --no-profile-stages 2>&1 | grep -c "SB_BOUNDS_CHECK"; done
  python bench.py --workload desc --steps 2 --no-cpu --no-extra 2>&1 | tail -1 | cut -c1-200
fi
