"""top stall instructions of an `ncu --page source --csv` export"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
si = hdr.index('# Samples'); src = hdr.index('Source'); ie = hdr.index('Instructions Executed')
data = [r for r in rows[2:] if len(r) == len(hdr) and r[si].isdigit()]
tot = sum(int(r[si]) for r in data)
print('total samples', tot, 'instrs', len(data))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
top = sorted(range(len(data)), key=lambda i: -int(data[i][si]))[:n]
for i in sorted(top):
    print(i, data[i][si], data[i][ie], data[i][src][:120])
