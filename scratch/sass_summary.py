#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that show which hardware path a kernel of libwc_b200.so uses
(B200_PROFILING.md "What proves a Blackwell-native kernel"):  cuobjdump -sass wc_gan_b200/libwc_b200.so"""
import collections, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else 'wc_gan_b200/libwc_b200.so'
out = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
pats = collections.OrderedDict([
    ('UTC*MMA (tcgen05.mma)', r'\bUTC[A-Z]*MMA'), ('LDTM/STTM (tcgen05.ld/st)', r'\b(LDTM|STTM)'),
    ('UTMALDG (TMA tensor load)', r'\bUTMALDG'), ('UTMASTG (TMA tensor store)', r'\bUTMASTG'),
    ('UTMAREDG (TMA reduce)', r'\bUTMAREDG'), ('UBLKCP (cp.async.bulk)', r'\bUBLKCP'),
    ('SYNCS (mbarrier)', r'\bSYNCS'), ('HMMA (mma.sync)', r'\bHMMA'), ('LDGSTS (cp.async)', r'\bLDGSTS'),
    ('FFMA', r'\bFFMA'), ('UCGABAR / cluster', r'\bUCGABAR'), ('instructions', r'^\s+/\*[0-9a-f]{4,5}\*/')])
cur, rows = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1); rows[cur] = collections.Counter(); continue
    if cur is None: continue
    for k, p in pats.items():
        if re.search(p, line): rows[cur][k] += 1
def short(n):
    d = subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip()
    d = re.sub(r'\(anonymous namespace\)::', '', d)
    return re.sub(r'\(.*', '', d)[:44]
print(f'# {so}: SASS mnemonic counts per kernel (sm_100a)')
print(f'{"kernel":44s} ' + ' '.join(f'{k.split()[0]:>9s}' for k in pats))
for n, c in rows.items():
    print(f'{short(n):44s} ' + ' '.join(f'{c[k]:9d}' for k in pats))
