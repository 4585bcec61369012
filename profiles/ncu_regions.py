"""Executed warp-instructions of one kernel bucketed by regions of its body: python profiles/ncu_regions.py rep kernel-regex body-file first-body-line b1,b2,...
SASS instructions are walked in address order; helper code inlined from other files / earlier lines is charged to the
body line that precedes it."""
import csv, subprocess, sys, io
rep, kre, body, first = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
bounds = [int(x) for x in sys.argv[5].split(",")]
idx = sys.argv[6] if len(sys.argv) > 6 else "1"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id", "::regex:" + kre + ":" + idx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
sass, cur, hdr, line = [], None, None, None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur, hdr = r[1].split("/")[-1], None; continue
    if r[0] == "Line No": hdr = r; continue
    if not hdr or not cur: continue
    if r[0].isdigit(): line = int(r[0]); continue
    if r[2].startswith("0x"):
        try: sass.append((int(r[2], 16), int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")]), cur, line))
        except Exception: pass
sass.sort()
tot = sum(s[1] for s in sass); ts = sum(s[2] for s in sass)
buckets = {}
cb = first
for addr, inst, samp, f, l in sass:
    if f == body and l >= first: cb = l
    k = max([b for b in bounds if b <= cb] or [first])
    a = buckets.setdefault(k, [0, 0, 0]); a[0] += inst; a[1] += samp; a[2] += 1
print("total executed", tot, "static", len(sass))
for k in sorted(buckets):
    i, s, n = buckets[k]
    print(f"from line {k:4d}: executed {100*i/tot:5.1f}%  samples {100*s/max(ts,1):5.1f}%  static {n}")
