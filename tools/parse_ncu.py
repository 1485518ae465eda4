import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=None
for r in rows:
    if len(r)>5 and r[0]=="ID": hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        print(d["ID"], d["Kernel Name"][:34], d["Metric Name"][:28], d["Metric Value"])
