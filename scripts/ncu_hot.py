"""Top stall-sample / instruction-count SASS lines of an `ncu --page source --csv` dump."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iN, iI, iT = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
body = [r for r in rows[2:] if len(r) > iT]
tot_s = sum(int(r[iN] or 0) for r in body); tot_i = sum(int(r[iI] or 0) for r in body)
print("instructions", len(body), "samples", tot_s, "warp-inst", tot_i)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print("-- by stall samples")
for k, r in sorted(enumerate(body), key=lambda t: -int(t[1][iN] or 0))[:top]:
    print(f"{k:5d} {int(r[iN]):7d} {100*int(r[iN])/max(tot_s,1):5.1f}%  inst {int(r[iI]):9d}  {r[iS].strip()[:90]}")
if len(sys.argv) > 3:
    # cumulative view: ranges "a-b" of instruction indices
    for rg in sys.argv[3:]:
        a, b = [int(v) for v in rg.split("-")]
        s = sum(int(r[iN] or 0) for r in body[a:b+1]); i = sum(int(r[iI] or 0) for r in body[a:b+1])
        t = sum(int(r[iT] or 0) for r in body[a:b+1])
        print(f"range {rg}: samples {100*s/max(tot_s,1):5.1f}%  warp-inst {100*i/max(tot_i,1):5.1f}%  lanes {t/max(i,1):.1f}")
