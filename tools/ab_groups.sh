#!/bin/bash
# A/B of group partitions / chunk lengths of the grouped persistent scans on ONE box:
#   tools/ab_groups.sh "64,40,44 64,40,44 16" "52,48,48 52,48,48 16" ...     (forward groups, backward groups, Tc)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "$@"; do
  set -- $cfg
  export PARROT_GROUPS_F=$1 PARROT_GROUPS_B=$2 PARROT_TC=$3
  tag=$(echo "$cfg" | tr ' ,' '__')
  timeout 150 python bench.py --steps 4 --warmup 3 --no_cpu_baseline 2>gpurun_out/abg_$tag.err | grep "^{" > gpurun_out/abg_$tag.json
  python - "$cfg" "gpurun_out/abg_$tag.json" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); s=d['sections_ms_persistent']
    print('%-28s %7.2f ms  fwd %.2f bwd %.2f' % (sys.argv[1], d['ms_per_step'], s['sec_scan_fwd']['ms'], s['sec_scan_bwd']['ms']))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
done
