"""Executed warp instructions per SASS opcode: python profiles/ncu_opcodes.py rep.ncu-rep <kernel-regex> [launch-index]"""
import csv, subprocess, sys, io, collections
rep, kre = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-id", "::regex:" + kre + ":" + (sys.argv[3] if len(sys.argv) > 3 else "1")],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
agg = collections.Counter(); stat = collections.Counter(); samp = collections.Counter()
for r in rows:
    if not r: continue
    if "Instructions Executed" in r and "Source" in r: hdr = r; continue
    if not hdr or len(r) != len(hdr): continue
    try:
        n = int(r[hdr.index("Instructions Executed")]); s = int(r[hdr.index("# Samples")])
    except ValueError:
        continue
    src = r[hdr.index("Source")].strip()
    toks = src.split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = ".".join(op.split(".")[:2]) if op.startswith(("LD", "ST", "ATOM", "RED")) else op.split(".")[0]
    agg[op] += n; stat[op] += 1; samp[op] += s
tot = sum(agg.values()) or 1; ts = sum(samp.values()) or 1
print("total executed", tot, "static", sum(stat.values()))
for op, n in agg.most_common(40):
    print(f"{op:12s} exec {100*n/tot:5.1f}%  samples {100*samp[op]/ts:5.1f}%  static {stat[op]}")
