"""Turn gpurun_out/<session> artefacts into the committed summaries under profiles/.
usage: python tools/summarise_profiles.py <session_dir> <tag>"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
src, tag = sys.argv[1], sys.argv[2]
out = os.path.join(ROOT, 'profiles')
os.makedirs(out, exist_ok=True)

# 1. launch list (ncu --metrics gpu__time_duration.sum --clock-control none): shares per kernel
for f in sorted(os.listdir(src)):
    if f.startswith('launches') and f.endswith('.csv'):
        rows = list(csv.DictReader(l for l in open(os.path.join(src, f)) if l.startswith('"')))
        agg = collections.defaultdict(lambda: [0, 0.0])
        for r in rows:
            n = r['Kernel Name'].split('(')[0]
            agg[n][0] += 1
            agg[n][1] += float(r['Metric Value']) / 1e6
        tot = sum(v[1] for v in agg.values())
        with open(os.path.join(out, '%s_%s.txt' % (tag, f[:-4])), 'w') as o:
            o.write('# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n')
            o.write('# source: %s ; %d launches, %.2f ms total\n' % (f, len(rows), tot))
            for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                o.write('%-70s n=%5d ms=%9.3f share=%5.1f%%\n' % (k[:70], v[0], v[1], 100 * v[1] / tot))

# 2. per-kernel CUDA-event breakdown of one frame (tools/profile_frame.py)
for f in sorted(os.listdir(src)):
    if f.startswith('profile_') and f.endswith('.json'):
        rows = json.load(open(os.path.join(src, f)))
        tot = sum(r['ms'] for r in rows)
        agg = {}
        for r in rows:
            opname = r.get('op', r.get('kind'))
            if 'Cin' not in r:
                key = ('scale%d' % r['scale'], opname if isinstance(opname, str) else 'conv')
            else:
                key = ('scale%d' % r['scale'], 'head' if r['Cout'] <= 3 else 'conv', r['Cin'], r['Cout'], '%dx%d' % tuple(r['k']),
                       's%d' % r['stride'], 'T' if r['transposed'] else '-', '%dx%d' % tuple(r['grid']), 'R%d' % r['R'],
                       '%dx%d' % (r['TH'], r['TW']))
            a = agg.setdefault(key, [0, 0.0, 0.0])
            a[0] += 1
            a[1] += r['ms']
            a[2] += r.get('macs', 0.0)
        with open(os.path.join(out, '%s_%s.txt' % (tag, f[:-5])), 'w') as o:
            o.write('# CUDA events after every kernel of one steady-state frame (v2v_plan_profile, eager, min of 3)\n')
            o.write('# total %.3f ms over %d kernels\n' % (tot, len(rows)))
            for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                tf = 2 * a[2] / (a[1] * 1e-3) / 1e12 if a[2] else 0
                o.write('%-78s n=%3d ms=%8.3f GMAC=%8.2f TFLOP/s=%7.1f\n' % (' '.join(str(x) for x in k), a[0], a[1], a[2] / 1e9, tf))

# 3. ncu --set full captures: the metrics the roofline uses
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']
for f in sorted(os.listdir(src)):
    if f.endswith('.ncu-rep'):
        txt = subprocess.run(['ncu', '-i', os.path.join(src, f), '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        with open(os.path.join(out, '%s_%s.txt' % (tag, f[:-8])), 'w') as o:
            o.write('# ncu --set full --clock-control none --import-source on ; source %s\n' % f)
            for r in rows[2:]:
                d = dict(zip(hdr, r))
                o.write('kernel %s grid %s block %s\n' % (d.get('Kernel Name', '?')[:60], d.get('Grid Size'), d.get('Block Size')))
                for i, h in enumerate(hdr):
                    if h in WANT:
                        o.write('   %-72s %s %s\n' % (h, r[i], units[i]))
print(os.listdir(out))
