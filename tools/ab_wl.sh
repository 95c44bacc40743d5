#!/bin/bash
# A/B of several builds of the library on one bench workload, interleaved (development aid): ab_wl.sh WORKLOAD REPS lib...
W=$1; R=$2; shift; shift
for i in $(seq $R); do
  for v in "$@"; do
    SB200_LIB=$v python bench.py --workload $W --no-cpu --no-extra --no-profile-stages --steps 10 2>/dev/null > /tmp/ab.json
    python - "$W $(basename $v)" <<'P'
import json, sys
d = json.loads(open('/tmp/ab.json').read().strip().splitlines()[-1])
print("%-36s value %.0f  e2e %.0f  sm_mhz %s" % (sys.argv[1], d["value"], d["e2e"]["value"], d["clocks"]["sm_mhz"]))
P
  done
done
