"""Print the SASS instructions that collect the most warp-stall samples in an `ncu --page source --csv --print-source sass` export."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
hdr = rows[1]; data = rows[2:]
ia = hdr.index("Source"); isamp = hdr.index("# Samples"); iex = hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp]) for r in data)
print("total samples", tot)
for k, r in enumerate(data):
    s = int(r[isamp])
    if s >= tot * frac:
        st = {hdr[i][6:]: int(r[i]) for i in stall_cols if int(r[i]) > 0}
        top = sorted(st.items(), key=lambda x: -x[1])[:3]
        print(k, r[ia].strip()[:80], s, r[iex], top)
