#!/bin/bash
# A/B of two environment settings on the default bench, interleaved (development aid): ab.sh "ENV_A" "ENV_B" [reps]
A="$1"; B="$2"; R=${3:-3}
for i in $(seq $R); do
  for v in "$A" "$B"; do
    env $v python bench.py --no-cpu --no-profile-stages --steps 10 2>/dev/null > /tmp/ab.json
    python - "$v" <<'P'
import json, sys
d = json.loads(open('/tmp/ab.json').read().strip().splitlines()[-1])
print("%-24s value %.0f  e2e %.0f  sm_mhz %s" % (sys.argv[1], d["value"], d["e2e"]["value"], d["clocks"]["sm_mhz"]))
P
  done
done
