"""Aggregate an ncu report's source page per CUDA source line: python profiles/ncu_lines.py rep.ncu-rep <kernel-regex> [top]"""
import csv, subprocess, sys, io
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", ] + (["--kernel-id", "::regex:" + kre + ":" + (sys.argv[4] if len(sys.argv) > 4 else "1")] if kre != "all" else []),
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg, cur, hdr = {}, None, None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur, hdr = r[1], None; continue
    if r[0] == "Function Name": print(r[1]); continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and cur:
        try:
            line = int(r[0]); inst = int(r[hdr.index("Instructions Executed")]); samp = int(r[hdr.index("# Samples")])
            conf = int(r[hdr.index("L1 Conflicts Shared N-Way")]) if "L1 Conflicts Shared N-Way" in hdr else 0
            exc = int(r[hdr.index("L1 Wavefronts Shared Excessive")]) if "L1 Wavefronts Shared Excessive" in hdr else 0
        except Exception:
            continue
        a = agg.setdefault((cur.split("/")[-1], line), [0, 0, 0, r[1]])
        a[0] += inst; a[1] += samp; a[2] += exc
ti = sum(a[0] for a in agg.values()) or 1; ts = sum(a[1] for a in agg.values()) or 1; te = sum(a[2] for a in agg.values()) or 1
print("total inst", ti, "samples", ts, "excess smem wavefronts", te)
for (f, l), (i, s, e, src) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{f}:{l:4d} inst {100*i/ti:5.1f}% samp {100*s/ts:5.1f}% exc {100*e/te:5.1f}%  {src.strip()[:100]}")
