#!/bin/bash
# per-octave blur launch times as a function of the marching run length (development aid):
#   seg_fine_sweep.sh WORKLOAD rows...   (rows = SB200_SEG_ROWS; 0 = the production choice)
W=$1; shift
for r in "$@"; do
  echo "## SEG_ROWS=$r"
  if [ "$r" = 0 ]; then python tools/fine_profile.py $W | grep -E "blur1|blur3|blur5|sum"
  else SB200_SEG_ROWS=$r python tools/fine_profile.py $W | grep -E "blur1|blur3|blur5|sum"; fi
done
