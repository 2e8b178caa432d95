"""Top stalled SASS instructions of one kernel from an .ncu-rep (source page, sass view)."""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = [i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r]
h = rows[hi[0]]
body = rows[hi[0] + 1: hi[1] - 1 if len(hi) > 1 else None]
si, ni, ii = h.index('# Samples'), h.index('Source'), h.index('Instructions Executed')
lsb, ssb, bar, wait = (h.index(k) for k in ('stall_long_sb', 'stall_short_sb', 'stall_barrier', 'stall_wait'))
def f(v):
    try: return float(v)
    except: return 0.0
tot = sum(f(r[si]) for r in body if len(r) > si)
print("total samples", tot, "instructions", len(body))
print("sum long_sb", sum(f(r[lsb]) for r in body if len(r) > lsb), "short_sb", sum(f(r[ssb]) for r in body if len(r) > ssb),
      "barrier", sum(f(r[bar]) for r in body if len(r) > bar), "wait", sum(f(r[wait]) for r in body if len(r) > wait))
idx = sorted(range(len(body)), key=lambda i: -f(body[i][si]) if len(body[i]) > si else 0)[:topn]
for i in sorted(idx):
    r = body[i]
    prev = body[i - 1][ni][:60] if i > 0 else ''
    print(f"{i:5d} samp={f(r[si]):7.0f} exec={f(r[ii]):9.0f} lsb={f(r[lsb]):6.0f} ssb={f(r[ssb]):6.0f} bar={f(r[bar]):5.0f} wait={f(r[wait]):5.0f}  {r[ni][:90]}")
