"""Lists the source lines of local-memory spill instructions (STL/LDL) per kernel of a variant build.
    python tools/spill_lines.py "-DPB_ENGINE_THREADS=384 -DPB_SPLIT_BATCH=3" scan_fwd_grouped"""
import collections, re, subprocess, sys
flags = sys.argv[1].split()
pat = sys.argv[2] if len(sys.argv) > 2 else 'persistent'
subprocess.check_call(['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-cubin',
                       '-o', 'gpurun_out/spill.cubin'] + flags + ['parrot_b200/csrc/api.cu'])
sass = subprocess.run(['nvdisasm', '--print-line-info', 'gpurun_out/spill.cubin'], capture_output=True, text=True,
                      errors='ignore').stdout
cur = fn = None
cnt = collections.Counter()
for line in sass.splitlines():
    m = re.match(r'\s*\.text\.(\S+):', line)
    if m:
        fn = m.group(1)
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if fn and pat in fn and re.search(r'\b(STL|LDL)(\.\w+)*\b', line):
        cnt[(fn[:40], cur, 'STL' if 'STL' in line else 'LDL')] += 1
for k, n in sorted(cnt.items(), key=lambda x: str(x[0])):
    print(k, n)
