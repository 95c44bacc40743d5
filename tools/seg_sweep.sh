#!/bin/bash
# sweeps the marching-blur segment height at the bench's batch size (development aid): [WORKLOAD=vga] seg_sweep.sh rows...
for r in "$@"; do
  SB200_SEG_ROWS=$r python bench.py --workload ${WORKLOAD:-1080p} --no-cpu --steps 6 2>/dev/null > /tmp/segsweep.json
  python - "$r" <<'P'
import json, sys
d = json.loads(open('/tmp/segsweep.json').read().strip().splitlines()[-1])
print(sys.argv[1], "value %.0f" % d["value"], "blur %.4f top %.4f" % (d["stages_ms_per_image"]["blur"], d["stages_ms_per_image"]["top_blur"]))
P
done
