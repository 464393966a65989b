#!/bin/bash
# A/B of variant builds (tools/build_variants.sh) on ONE box: tools/ab_variants.sh base tmpl0 tmpl2 ...
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for name in "$@"; do
  if [ "$name" = base ]; then unset PARROT_B200_LIB; else export PARROT_B200_LIB=$PWD/build_variants/$name.so; fi
  timeout 150 python bench.py --no_cpu_baseline 2>gpurun_out/abv_$name.err | grep "^{" > gpurun_out/abv_$name.json
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.load(open('gpurun_out/abv_%s.json'%n))
    s=d['sections_ms_persistent']
    print('%-10s %9.0f f/s %7.2f ms  fwd %.2f bwd %.2f tail %.2f' % (n, d['value'], d['ms_per_step'], s['sec_scan_fwd']['ms'], s['sec_scan_bwd']['ms'], s['sec_grads_tail']['ms']))
except Exception as e:
    print(n, 'FAILED', e)
PY
done
