#!/bin/bash
# Round-end evidence run (one gpurun call): bench line, launch list, ncu captures of the persistent scans, smoke.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python bench.py 2>gpurun_out/final_bench.err | grep "^{" > gpurun_out/final_bench.json
python -c "
import json; d=json.load(open('gpurun_out/final_bench.json')); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_final.csv \
  python bench.py --steps 2 --warmup 1 --no_cpu_baseline > gpurun_out/launches_final.log 2>&1
echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name regex:"scan_(fwd|bwd)_persistent" \
  --launch-skip 2 -c 2 -f -o gpurun_out/persist_T40 python bench.py --frames 40 --steps 1 --warmup 1 --no_cpu_baseline --profile_steps 0 > gpurun_out/persist_T40.log 2>&1
echo "ncu T40 rc=$?"
timeout 400 ncu --set full --clock-control none --kernel-name regex:"scan_(fwd|bwd)_persistent" \
  --launch-skip 2 -c 2 -f -o gpurun_out/persist_T800 python bench.py --steps 1 --warmup 1 --no_cpu_baseline --profile_steps 0 > gpurun_out/persist_T800.log 2>&1
echo "ncu T800 rc=$?"
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 250 python bench.py --impl reference --steps 1 --warmup 1 2>/dev/null | grep "^{" | head -c 600
