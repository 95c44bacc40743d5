#!/bin/bash
# per-octave blur launch times as a function of the pieces per column of the aligned distribution (development aid):
#   pieces_sweep.sh WORKLOAD k...   (k = SB200_PIECES; 0 = the production choice)
W=$1; shift
for r in "$@"; do
  echo "## PIECES=$r"
  if [ "$r" = 0 ]; then python tools/fine_profile.py $W | grep -E "blur1|blur3|blur5|sum"
  else SB200_PIECES=$r python tools/fine_profile.py $W | grep -E "blur1|blur3|blur5|sum"; fi
done
