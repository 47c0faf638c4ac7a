"""Top SASS lines by stall samples per kernel from `ncu -i X.ncu-rep --page source --csv`: python tools/ncu_hot.py src.csv [top]"""
import csv, sys
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
kern, hdr, rows = None, None, []
def flush():
    if not rows: return
    tot = sum(int(r[2] or 0) for r in rows)
    inst = sum(int(r[5] or 0) for r in rows)
    print("==", kern[:90], "samples", tot, "warp-instructions", inst)
    order = sorted(range(len(rows)), key=lambda i: -int(rows[i][2] or 0))[:top]
    for i in sorted(order):
        r = rows[i]
        stalls = {h: int(v) for h, v in zip(hdr[30:47], r[30:47]) if v and int(v) > 0}
        main = sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
        print("%5d %5.1f%% exec %9s  %-70s %s" % (i, 100.0 * int(r[2] or 0) / max(tot, 1), r[5], r[1].strip()[:70], main))
for r in csv.reader(open(sys.argv[1])):
    if not r: continue
    if r[0] == "Kernel Name":
        flush(); kern, rows = r[1], []
    elif r[0] == "Address": hdr = r
    else: rows.append(r)
flush()
