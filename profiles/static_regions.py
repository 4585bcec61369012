"""Static SASS instruction count of one kernel bucketed by body regions: python profiles/static_regions.py all.sass kernel-substr body-file first b1,b2,..."""
import re, sys, collections
path, kname, body, first = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
bounds = [int(x) for x in sys.argv[5].split(",")]
cnt = collections.Counter(); infn = False; cb = first; tot = 0
for l in open(path):
    if re.match(r'\s*\.section\s+\.text\.', l) or l.startswith('.text.'):
        infn = kname in l
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        f, ln = m.group(1).split('/')[-1], int(m.group(2))
        if f == body and ln >= first: cb = ln
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        cnt[max([b for b in bounds if b <= cb] or [first])] += 1; tot += 1
print("static total", tot)
for k in sorted(cnt): print(f"from line {k:4d}: {cnt[k]}")
