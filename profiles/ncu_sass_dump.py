"""SASS of one kernel in address order with executed counts and the CUDA line each instruction belongs to:
python profiles/ncu_sass_dump.py rep kernel-regex [launch-index] > file"""
import csv, subprocess, sys, io
rep, kre = sys.argv[1], sys.argv[2]
idx = sys.argv[3] if len(sys.argv) > 3 else "1"
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
        try: sass.append((int(r[2], 16), int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")]), cur, line, r[3]))
        except Exception: pass
sass.sort()
for addr, inst, samp, f, l, src in sass:
    print(f"{addr:6x} {inst:10d} {samp:5d} {f}:{l:<5d} {src.strip()}")
