for cfg in "--batch 32 --groups 4" "--batch 64 --groups 2" "--batch 64 --groups 4" "--batch 16 --groups 8" "--workload 4k --batch 8 --groups 2" "--workload 4k --batch 16 --groups 2" "--workload vga --batch 128 --groups 4" "--workload vga --batch 256 --groups 2" "--workload vga --batch 256 --groups 4"; do
  python bench.py $cfg --no-cpu --no-extra --no-profile-stages --steps 10 2>/dev/null > /tmp/ab.json
  python - "$cfg" <<'P'
import json, sys
d = json.loads(open('/tmp/ab.json').read().strip().splitlines()[-1])
print("%-40s value %.0f  e2e %.0f  sm_mhz %s" % (sys.argv[1], d["value"], d["e2e"]["value"], d["clocks"]["sm_mhz"]))
P
done
