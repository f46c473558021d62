"""Split the warp-stall samples of an `ncu --page source --csv --print-source sass` export of conv_tc_kernel by warp
role (producer / MMA issuer / epilogue), using the role-entry branches' executed-instruction counts as separators."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
isamp = hdr.index("# Samples"); isrc = hdr.index("Source"); iex = hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
tot = sum(int(r[isamp]) for r in data)
print("total samples", tot)
# role boundaries: first UTMALDG = producer region, first UTCHMMA = MMA region, first LDTM = epilogue region
def first(pat):
    for k, r in enumerate(data):
        if pat in r[isrc]:
            return k
    return None
marks = sorted((k, n) for k, n in ((first("UTMALDG"), "producer"), (first("UTCHMMA"), "mma"), (first("LDTM"), "epilogue")) if k is not None)
print("first marker rows:", marks)
def summ(lo, hi):
    st = {}
    for r in data[lo:hi]:
        for i in stall:
            st[hdr[i][6:]] = st.get(hdr[i][6:], 0) + int(r[i])
    return sum(int(r[isamp]) for r in data[lo:hi]), sorted(st.items(), key=lambda x: -x[1])[:6]
if len(sys.argv) > 3:
    b = [int(x) for x in sys.argv[3].split(",")]
    for lo, hi in zip(b[:-1], b[1:]):
        print(lo, hi, summ(lo, hi))
for k, r in enumerate(data):
    s = int(r[isamp])
    if s >= tot * thr:
        st = {hdr[i][6:]: int(r[i]) for i in stall if int(r[i]) > 0}
        top = sorted(st.items(), key=lambda x: -x[1])[:3]
        print(k, r[isrc].strip()[:70], s, r[iex], top)
