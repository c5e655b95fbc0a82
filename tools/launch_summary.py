"""Aggregate an ncu launch list (csv with gpu__time_duration.sum, sm__cycles_active.sum, smsp__inst_executed.sum) of
tools/ncu_solve.py: per kernel launches, summed duration and SM-active time of the LAST (warm) solve.
usage: python tools/launch_summary.py gpurun_out/solve_launches_r1.csv [nsolves=2]"""
import csv, collections, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
nsolves = int(sys.argv[2]) if len(sys.argv) > 2 else 2
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
recs = collections.OrderedDict()
for r in rows[1:]:
    k = int(r[ix['ID']])
    recs.setdefault(k, {'name': r[ix['Kernel Name']]})[r[ix['Metric Name']]] = float(r[ix['Metric Value']].replace(',', ''))
ids = sorted(recs); n = len(ids); last = ids[n - n // nsolves:]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for k in last:
    d = recs[k]; nm = re.sub(r'<.*', '', d['name']).replace('void ', '').replace('pf::', '')
    a = agg[nm]; a[0] += 1; a[1] += d['gpu__time_duration.sum'] / 1e3; a[2] += d.get('sm__cycles_active.sum', 0); a[3] += d.get('smsp__inst_executed.sum', 0)
tt = sum(a[1] for a in agg.values()); tc = sum(a[2] for a in agg.values())
print("launches in the last solve: %d; sum of kernel durations %.2f ms; SM-active time / (148 SMs x 1.965 GHz) = %.2f ms" % (len(last), tt / 1e3, tc / 148 / 1.965e6))
print("| kernel | launches | sum of durations us | share | SM-active equivalent us | share | warp instructions (M) |")
print("|---|---:|---:|---:|---:|---:|---:|")
for nm, a in sorted(agg.items(), key=lambda t: -t[1][1]):
    print("| `%s` | %d | %.1f | %.1f %% | %.1f | %.1f %% | %.1f |" % (nm, a[0], a[1], 100 * a[1] / tt, a[2] / 148 / 1965, 100 * a[2] / max(tc, 1), a[3] / 1e6))
