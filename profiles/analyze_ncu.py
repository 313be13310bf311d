"""Summarise an .ncu-rep (raw + source pages) the way the profiles/*.md notes quote it.
usage: python profiles/analyze_ncu.py gpurun_out/prof.ncu-rep"""
import csv, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__thread_inst_executed_per_inst_executed.ratio']
stalls = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
seen = set()
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    if name in seen: continue
    seen.add(name)
    print('====', name[:70])
    for w in want:
        if w in hdr: print('   %-62s %s %s' % (w, r[hdr.index(w)][:16], rows[1][hdr.index(w)]))
    st = sorted(((float(r[hdr.index(h)]), h) for h in stalls), reverse=True)[:6]
    print('   stalls/issue:', ', '.join('%s=%.2f' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v) for v, h in st))
rows = list(csv.reader(io.StringIO(src)))
ks = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
seen = set()
for n, k in enumerate(ks):
    if rows[k][1] in seen: continue
    seen.add(rows[k][1])
    end = ks[n + 1] if n + 1 < len(ks) else len(rows)
    h2 = rows[k + 1]; body = rows[k + 2:end]; ci = {h: i for i, h in enumerate(h2)}
    mx = max(int(r[ci['Instructions Executed']]) for r in body)
    hot = [r for r in body if int(r[ci['Instructions Executed']]) >= 0.5 * mx]
    tot = sum(int(r[ci['Instructions Executed']]) for r in body)
    print('====', rows[k][1][:70]); print('   total inst %d, max line count %d, inst/step %.1f' % (tot, mx, tot / mx))
    ops = Counter()
    for r in hot:
        s = r[ci['Source']].strip().split()
        ops[s[1] if s[0].startswith('@') else s[0]] += int(r[ci['Instructions Executed']]) / mx
    print('   hot mix:', [(o, round(c, 1)) for o, c in ops.most_common(14)])
    for r in sorted(body, key=lambda r: -int(r[ci['# Samples']]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 8]:
        st = {h: int(r[ci[h]]) for h in h2 if h.startswith('stall_') and 'Not' not in h and int(r[ci[h]]) > 0}
        print('   %6s %9s %-58s %s' % (r[ci['# Samples']], r[ci['Instructions Executed']], r[ci['Source']].strip()[:58],
                                       sorted(st.items(), key=lambda kv: -kv[1])[:2]))
