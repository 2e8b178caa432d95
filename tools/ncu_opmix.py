"""Dynamic SASS opcode mix of one kernel from an .ncu-rep (needs the source page)."""
import csv, io, subprocess, sys, collections, re
rep, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = [i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r]
h = rows[hi[0]]
body = rows[hi[0] + 1: hi[1] - 1 if len(hi) > 1 else None]
ni, ii = h.index('Source'), h.index('Instructions Executed')
mix = collections.Counter()
for r in body:
    if len(r) <= ii: continue
    try: n = float(r[ii])
    except: continue
    m = re.match(r'\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', r[ni])
    if m: mix[m.group(1)] += n
tot = sum(mix.values())
print(f"total warp-instructions {tot:.3e}")
groups = {'fp': ('FFMA', 'FADD', 'FMUL'), 'int': ('IADD3', 'IMAD', 'LEA', 'LOP3', 'SHF', 'ISETP', 'SEL', 'MOV', 'IABS', 'VIMNMX', 'UIADD3', 'UIMAD', 'UMOV', 'ULEA', 'USHF', 'UISETP', 'PLOP3', 'HFMA2', 'CS2R', 'IMNMX'),
          'smem': ('LDS', 'STS'), 'gmem': ('LDG', 'STG', 'LDGSTS', 'LDGDEPBAR', 'DEPBAR'), 'local': ('LDL', 'STL'),
          'ctrl': ('BRA', 'BSSY', 'BSYNC', 'BAR', 'WARPSYNC', 'NOP', 'EXIT', 'CALL', 'RET'), 'const': ('LDC', 'LDCU', 'S2R', 'S2UR')}
for g, ops in groups.items():
    n = sum(mix[o] for o in ops)
    print(f"  {g:6s} {n / tot * 100:5.1f}%   " + ", ".join(f"{o}:{mix[o] / tot * 100:.1f}" for o in ops if mix[o] / tot > 0.004))
other = {o: n for o, n in mix.items() if not any(o in ops for ops in groups.values())}
print("  other  " + ", ".join(f"{o}:{n / tot * 100:.1f}" for o, n in sorted(other.items(), key=lambda kv: -kv[1])[:12]))
