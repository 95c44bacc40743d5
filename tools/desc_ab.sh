#!/bin/bash
# descriptor-only workload (200k keypoints) with several builds of the library, interleaved: desc_ab.sh lib...
for i in 1 2; do
for v in "$@"; do
  SB200_LIB=$v python bench.py --workload desc --no-cpu --no-extra --steps 6 2>/dev/null | tail -1 > /tmp/d.json
  python - "$(basename $v)" <<'P'
import json, sys
d = json.loads(open('/tmp/d.json').read())
print("%-28s %.1f M keypoints/s  %.3f ns/keypoint" % (sys.argv[1], d["value"] / 1e6, d["ns_per_keypoint"]))
P
done
done
