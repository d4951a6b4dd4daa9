"""Hot spots of one kernel from an .ncu-rep source page: the SASS instructions with the most
stall samples / executed instructions, in program order with a little context.
    python tools/ncu_source_hot.py rep.ncu-rep "<kernel name substring>" [top]"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(raw)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}
        blocks.append(cur)
    elif cur is not None and row:
        cur["rows"].append(row)
for b in blocks:
    if pat not in b["name"]:
        continue
    hdr, rows = b["rows"][0], b["rows"][1:]
    ci = {h: i for i, h in enumerate(hdr)}
    S, E, T = ci["# Samples"], ci["Instructions Executed"], ci["Avg. Threads Executed"]
    tot_s = sum(int(r[S]) for r in rows)
    tot_e = sum(int(r[E]) for r in rows)
    print(f"## {b['name'][:100]}\n   instructions {len(rows)}  samples {tot_s}  warp-instructions executed {tot_e}")
    order = sorted(range(len(rows)), key=lambda i: -int(rows[i][S]))[:top]
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for i in sorted(order):
        r = rows[i]
        stalls = sorted(((int(r[ci[h]]), h[6:]) for h in stall_cols if int(r[ci[h]]) > 0), reverse=True)[:3]
        print(f"{i:5d} {int(r[S]):7d} smp {100.0 * int(r[S]) / max(tot_s, 1):5.1f}%  exec {int(r[E]):9d}  thr {r[T]:>5s}  "
              f"{r[ci['Source']].strip()[:70]:70s} {stalls}")
    break
