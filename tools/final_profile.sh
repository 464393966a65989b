#!/bin/bash
# Round-end evidence run (one gpurun call): bench line, launch list, ncu capture of the grouped scans, per-group
# timelines, smoke, reference arm.  Summaries are made afterwards with tools/summarize_ncu.py (no GPU needed).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 python bench.py 2>gpurun_out/final_bench.err | grep "^{" > gpurun_out/final_bench.json
python -c "
import json; d=json.load(open('gpurun_out/final_bench.json')); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d.get('sample', {}).get('us_per_step'))"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_final.csv \
  python bench.py --steps 2 --warmup 1 --no_cpu_baseline --no_sample > gpurun_out/launches_final.log 2>&1
echo "launch list rc=$?"
timeout 500 ncu --set full --clock-control none --kernel-name regex:"scan_(fwd|bwd)_grouped" \
  --launch-skip 2 -c 2 -f -o gpurun_out/grouped_T800 python bench.py --steps 1 --warmup 1 --no_cpu_baseline --no_sample --profile_steps 0 > gpurun_out/grouped_T800.log 2>&1
echo "ncu T800 rc=$?"
timeout 200 python tools/group_timeline.py 256 fwd > gpurun_out/final_gt_fwd.txt 2>&1
timeout 200 python tools/group_timeline.py 256 bwd > gpurun_out/final_gt_bwd.txt 2>&1
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 250 python bench.py --impl reference --steps 1 --warmup 1 2>/dev/null | grep "^{" | head -c 700
