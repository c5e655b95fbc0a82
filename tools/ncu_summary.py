"""Summarise an ncu report: per-kernel headline metrics and the SASS opcode mix / stall reasons.
usage: python tools/ncu_summary.py report.ncu-rep [kernel-regex]"""
import collections, csv, subprocess, sys, io
rep = sys.argv[1]; rx = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units = rows[0], rows[1]; idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
for r in rows[2:]:
    print("==", r[idx['Kernel Name']][:70])
    print("   " + "  ".join("%s=%s%s" % (w.split('.')[0].split('__')[-1], r[idx[w]], units[idx[w]][:6]) for w in want if w in idx))
if rx:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-count", "1"] + (["--launch-skip", sys.argv[3]] if len(sys.argv) > 3 else []),
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
    tot = stot = 0; byop = collections.Counter(); samp = collections.Counter(); reasons = collections.Counter()
    st = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    for r in rows[2:]:
        try: n = int(r[idx['Instructions Executed']]); s = int(r[idx['# Samples']])
        except Exception: continue
        t = r[idx['Source']].strip(); op = (t.split()[1] if t.startswith('@') else t.split()[0]).split('.')[0]
        byop[op] += n; samp[op] += s; tot += n; stot += s
        for i, h in st:
            try: reasons[h] += int(r[i])
            except Exception: pass
    print("total warp instr", tot, "samples", stot)
    for op, n in byop.most_common(16): print("   %-10s %5.1f%% of instr   %5.1f%% of samples" % (op, 100 * n / tot, 100 * samp[op] / max(stot, 1)))
    rt = sum(reasons.values())
    print("   stalls: " + "  ".join("%s %.0f%%" % (h.replace('stall_', ''), 100 * v / rt) for h, v in reasons.most_common(8)))
