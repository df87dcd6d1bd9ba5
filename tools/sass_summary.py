"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native path (B200_PROFILING.md): UTC*MMA (tcgen05.mma),
LDTM/STTM (tcgen05.ld/st), UTMALDG/UTMASTG (TMA tensor copies), UBLKCP (bulk copies), UTCBAR (tcgen05.commit), SYNCS (mbarrier),
plus registers / shared memory from the ELF.  Runs here (no GPU): python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'bodyfitting_b200', 'libbodyfit_b200.so')
sass = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
pats = collections.OrderedDict([('UTC*MMA', r'\bUTC[A-Z]*MMA'), ('LDTM', r'\bLDTM'), ('STTM', r'\bSTTM'), ('UTMALDG', r'\bUTMALDG'),
                                ('UTMASTG', r'\bUTMASTG'), ('UBLKCP', r'\bUBLKCP'), ('UTCBAR', r'\bUTCBAR'), ('SYNCS', r'\bSYNCS'),
                                ('HMMA', r'\bHMMA'), ('FFMA', r'\bFFMA'), ('ATOM/RED', r'\b(ATOMG|ATOMS|RED|ATOM)\b')])
cur, rows = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = subprocess.run(['cu++filt', m.group(1)], capture_output=True, text=True).stdout.strip().replace('(int)', '').replace('void ', '').split('(')[0]
        rows[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for name, pat in pats.items():
        if re.search(pat, line):
            rows[cur][name] += 1
res = subprocess.run(['cuobjdump', '-res-usage', so], capture_output=True, text=True).stdout
usage = {}
for m in re.finditer(r'Function (\S+):\s*\n\s*REG:(\d+).*?SHARED:(\d+)', res):
    usage[subprocess.run(['cu++filt', m.group(1)], capture_output=True, text=True).stdout.strip().replace('(int)', '').replace('void ', '').split('(')[0]] = (m.group(2), m.group(3))
print('# SASS evidence for %s (sm_100a), `cuobjdump -sass` mnemonic counts per kernel' % os.path.basename(so))
print('%-44s %5s %7s ' % ('kernel', 'regs', 'smem_st') + ' '.join('%8s' % k for k in pats))
tot = collections.Counter()
for k, c in rows.items():
    r, sm = usage.get(k, ('?', '?'))
    print('%-44s %5s %7s ' % (k[:44], r, sm) + ' '.join('%8d' % c[n] for n in pats))
    tot.update(c)
print('%-44s %5s %7s ' % ('TOTAL', '', '') + ' '.join('%8d' % tot[n] for n in pats))
