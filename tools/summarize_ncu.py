"""Turns an .ncu-rep capture into the short text summary committed under profiles/ (run here, no GPU needed).

    python tools/summarize_ncu.py gpurun_out/prof_persist.ncu-rep profiles/r01_scan_fwd_persistent_ncu.txt
    python tools/summarize_ncu.py --launches gpurun_out/launches.csv profiles/r01_launch_list.txt
    python tools/summarize_ncu.py --traffic gpurun_out/grouped_T800.ncu-rep profiles/r02_scan_grouped_ncu_T800.txt \
        profiles/r02_scan_traffic.json        (dram bytes per launch of the scan kernels -> bench.py roofline.traffic)
"""
import csv
import collections
import json
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__cycles_active.avg', 'sm__cycles_elapsed.max',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__inst_executed.sum',
        'smsp__warp_issue_stalled_barrier_per_warp_active.pct', 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct']


def summarize_rep(rep, out):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, 'w') as f:
        f.write('# ncu --set full --clock-control none capture: %s\n' % rep)
        for r in rows[2:]:
            f.write('kernel: %s\n' % r[idx['Kernel Name']])
            for k in KEYS:
                if k in idx:
                    f.write('  %-75s %s %s\n' % (k, r[idx[k]], units[idx[k]]))
            f.write('\n')
        sass = subprocess.run(['cuobjdump', '-sass', 'parrot_b200/libparrot_b200.so'], capture_output=True, text=True).stdout
        cnt = collections.Counter(w for w in sass.replace(';', ' ').split() if w.split('.')[0] in
                                  ('UTCHMMA', 'UTMALDG', 'LDTM', 'STTM', 'UTCBAR', 'UBLKCP', 'HMMA'))
        f.write('SASS mnemonics in libparrot_b200.so: %s\n' % dict(cnt))


def summarize_launches(path, out):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(row['Metric Value'].replace(',', ''))
        v = v / 1000 if row['Metric Unit'] == 'ns' else (v * 1000 if row['Metric Unit'] == 'ms' else v)
        name = row['Kernel Name'].split('(')[0]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, 'w') as f:
        f.write('# ncu --metrics gpu__time_duration.sum --clock-control none launch list: %s\n' % path)
        f.write('# (cold-cache, serialised: compare SHARES, not absolutes)\n')
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('%-42s n=%6d total %11.1f us  avg %9.2f us  share %5.1f%%\n' % (k[:42], v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
        f.write('total %.1f us\n' % tot)


def traffic_json(rep, txt, out):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
    res = {'source': txt}
    for r in rows[2:]:
        name = r[idx['Kernel Name']].split('(')[0]
        tot = 0.0
        for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            tot += float(r[idx[k]].replace(',', '')) * scale.get(units[idx[k]], 1.0)
        res[name] = tot
    json.dump(res, open(out, 'w'), indent=1)


if __name__ == '__main__':
    if sys.argv[1] == '--launches':
        summarize_launches(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == '--traffic':
        summarize_rep(sys.argv[2], sys.argv[3])
        traffic_json(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        summarize_rep(sys.argv[1], sys.argv[2])
