"""top stall-sample SASS lines of the first kernel in an ncu report: python scripts/ncu_hot.py rep [N]"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
body = []
for r in rows[hdr_i + 1:]:
    if not r or r[0] == "Kernel Name":
        break
    body.append(r)
si = hdr.index("# Samples")
tot = sum(int(r[si] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
for idx, r in sorted(enumerate(body), key=lambda t: -int(t[1][si] or 0))[:n]:
    print(f"{int(r[si]):7d} {100*int(r[si])/max(tot,1):5.1f}%  #{idx:5d}  {r[1].strip()[:110]}")
