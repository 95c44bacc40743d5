#!/bin/bash
# A/B of two environment settings on the bench workloads, interleaved (development aid): ab_env.sh "ENV_A" "ENV_B" [reps] [workloads]
A="$1"; B="$2"; R=${3:-2}; WL=${4:-"1080p vga 4k"}
for wl in $WL; do
for i in $(seq $R); do
  for v in "$A" "$B"; do
    env $v python bench.py --workload $wl --no-cpu --no-extra --no-profile-stages --steps 10 2>/dev/null > /tmp/ab.json
    python - "$wl $v" <<'P'
import json, sys
d = json.loads(open('/tmp/ab.json').read().strip().splitlines()[-1])
print("%-36s value %.0f  e2e %.0f  sm_mhz %s" % (sys.argv[1], d["value"], d["e2e"]["value"], d["clocks"]["sm_mhz"]))
P
  done
done
done
