import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; data=[r for r in rows[2:] if len(r)==len(hdr) and r[hdr.index("# Samples")].isdigit()]
si=hdr.index('# Samples'); src=hdr.index('Source')
tot=sum(int(r[si]) for r in data if r[si].isdigit())
print('total samples',tot, 'instrs', len(data))
n=int(sys.argv[2]) if len(sys.argv)>2 else 40
top=sorted(range(len(data)), key=lambda i:-int(data[i][si] or 0))[:n]
for i in sorted(top):
    print(i, data[i][si], data[i][src][:110])
