"""Print the first eager pass of gpurun_out/launches.csv next to a saved baseline."""
import csv, re, sys
def load(p):
    return list(csv.DictReader([l for l in open(p) if not l.startswith('==')]))
rows = load(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/launches.csv')
base = load(sys.argv[2]) if len(sys.argv) > 2 else None
NAMES = ["pack_in","pre_1","pre_2","c1_1","c1_2","c3d_1","c2_1","c2_2","c3d_2","c3_1","c3_2","c4_1","c4_2","warp","gru1","gru2","gru3","c5_1","c5_2","c6_1","c6_2","c7_1","c7_2","c8_1","c8_2","head1","head2"]
tot = btot = 0
for i, row in enumerate(rows[:27]):
    t = float(row['Metric Value']) / 1e3; tot += t
    b = float(base[i]['Metric Value']) / 1e3 if base else 0; btot += b
    print("%2d %-8s %-14s %8.0f us   (base %6.0f)" % (i, NAMES[i], row['Grid Size'], t, b))
print("total %.0f us (base %.0f)" % (tot, btot))
