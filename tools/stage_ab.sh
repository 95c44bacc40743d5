#!/bin/bash
# serialised stage times (fine_profile's last line) of several builds of the library, interleaved (development aid):
#   stage_ab.sh WORKLOAD lib...
W=$1; shift
for i in 1 2; do
for v in "$@"; do
  echo -n "$(basename $v) "; SB200_LIB=$v python tools/fine_profile.py $W | tail -1
done
done
