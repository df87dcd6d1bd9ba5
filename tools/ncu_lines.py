"""Per-source-line hot spots of an ncu report (needs -lineinfo + --import-source on).
   python tools/ncu_lines.py report.ncu-rep [topN]"""
import csv, subprocess, sys, io, os
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
H = None; fname = '?'; items = []
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fname = os.path.basename(r[1]); continue
    if r[0] == 'Line No': H = r; continue
    if H is None or not r[0].isdigit(): continue
    try:
        ii, ss = H.index('Instructions Executed'), H.index('Warp Stall Sampling (All Samples)')
        items.append((float(r[ss] or 0), float(r[ii] or 0), fname, int(r[0]), r[1].strip()[:100]))
    except (ValueError, IndexError):
        pass
ti, ts = sum(i[1] for i in items) or 1, sum(i[0] for i in items) or 1
print('total warp instructions %.3g, stall samples %.3g' % (ti, ts))
for s, i, f, ln, txt in sorted(items, reverse=True)[:topn]:
    print('%-16s %4d  inst %5.1f%%  stall %5.1f%%  %s' % (f, ln, 100 * i / ti, 100 * s / ts, txt))
