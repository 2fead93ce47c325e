"""profiles/conv_dram_traffic.json from an ncu CSV of `bench.py` (metrics dram__bytes_read.sum, dram__bytes_write.sum,
gpu__time_duration.sum, --print-units base): mean DRAM bytes per conv_umma_kernel launch, plus the per-kernel shares."""
import csv, json, os, sys
src, workload, tag = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(open(src, errors='replace')) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
H = rows[hdr]
ki, mi, vi, ii = H.index('Kernel Name'), H.index('Metric Name'), H.index('Metric Value'), H.index('ID')
per = {}
for r in rows[hdr + 1:]:
    try:
        per.setdefault((r[ii], r[ki]), {})[r[mi]] = float(r[vi].replace(',', ''))
    except (ValueError, IndexError):
        pass
agg = {}
for (_, k), m in per.items():
    name = k.split('(')[0].split('<')[0].split('::')[-1]
    a = agg.setdefault(name, dict(n=0, ns=0.0, rd=0.0, wr=0.0))
    a['n'] += 1; a['ns'] += m.get('gpu__time_duration.sum', 0.0)
    a['rd'] += m.get('dram__bytes_read.sum', 0.0); a['wr'] += m.get('dram__bytes_write.sum', 0.0)
tot = sum(a['ns'] for a in agg.values())
out_path = os.path.join(os.path.dirname(__file__), '..', 'profiles', 'conv_dram_traffic.json')
try:
    out = json.load(open(out_path))
except Exception:
    out = {}
c = agg['conv_umma_kernel']
out[workload] = {'dram_bytes_per_launch_avg': (c['rd'] + c['wr']) / c['n'], 'dram_read_bytes_per_launch_avg': c['rd'] / c['n'],
                 'dram_write_bytes_per_launch_avg': c['wr'] / c['n'], 'launches_captured': c['n'],
                 'source': 'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum of `python bench.py --steps 2 --warmup 3` (%s)' % tag}
json.dump(out, open(out_path, 'w'), indent=1)
print('# %s: kernel, launches, share of device time, total ms, DRAM GB read, GB written' % tag)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['ns']):
    print('%-34s n=%5d share=%5.1f%% ms=%9.3f rd=%8.3f wr=%8.3f' % (k, a['n'], 100 * a['ns'] / tot, a['ns'] / 1e6, a['rd'] / 1e9, a['wr'] / 1e9))
