"""Per-source-line stall samples of one kernel: joins `ncu --page source --csv` (SASS rows with sampling columns) with the line table
nvdisasm prints for the same cubin.  usage: ncu_lines.py report.ncu-rep libmcba.so kernel_mangled_substring [top]"""
import csv, re, subprocess, sys, tempfile, os, collections
rep, lib, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
addr2line, cur, inside = {}, None, False
for l in sass:
  if l.startswith(".text."):
    inside = kern in l; continue
  if not inside: continue
  m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
  if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
  m = re.match(r'\s*/\*([0-9a-f]+)\*/', l)
  if m: addr2line[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ia, isamp, iinst = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = None
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot = 0
for r in rows[hi + 1:]:
  if len(r) != len(hdr): continue
  a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
  if base is None: base = a
  ln = addr2line.get(a - base)
  s = int(r[isamp] or 0); tot += s
  e = agg[ln]; e[0] += s; e[1] += int(r[iinst] or 0)
  for i in stall_cols:
    v = int(r[i] or 0)
    if v: e[2][hdr[i][6:]] += v
print("total samples", tot)
for ln, (s, n, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
  print("%6.2f%% %9d inst  %-28s %s" % (100.0 * s / tot, n, "%s:%d" % ln if ln else "?", " ".join("%s=%d" % kv for kv in st.most_common(4))))
if len(sys.argv) > 5:      # regions: "name:file:lo-hi,..."
  regs = []
  for spec in sys.argv[5].split(","):
    nm, fl, rng = spec.split(":"); lo, hi = rng.split("-"); regs.append((nm, fl, int(lo), int(hi)))
  ragg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
  for ln, (s, n, st) in agg.items():
    nm = "other"
    if ln:
      for r in regs:
        if ln[0] == r[1] and r[2] <= ln[1] <= r[3]: nm = r[0]; break
      else: nm = "other:" + ln[0]
    e = ragg[nm]; e[0] += s; e[1] += n; e[2].update(st)
  print("---- regions")
  for nm, (s, n, st) in sorted(ragg.items(), key=lambda kv: -kv[1][0]):
    print("%6.2f%% %10d inst  %-28s %s" % (100.0 * s / tot, n, nm, " ".join("%s=%d" % kv for kv in st.most_common(5))))
