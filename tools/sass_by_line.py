"""Attribute the SASS instructions of one kernel to source lines (static counts): python tools/sass_by_line.py lib.so 'kernel substring'."""
import collections, re, subprocess, sys, tempfile, os, glob
so, key = sys.argv[1], sys.argv[2]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
    txt = subprocess.run(["nvdisasm", "--print-line-info", "-c", cubin], capture_output=True, text=True).stdout
    for f in re.split(r"\n//-+ \.text\.", txt)[1:]:
        name = f.split(" ", 1)[0]
        if key not in name:
            continue
        cur, cnt, ops = None, collections.Counter(), collections.Counter()
        for line in f.splitlines():
            m = re.search(r'//## File "([^"]+)", line (\d+)', line)
            if m:
                cur = (m.group(1).split("/")[-1], int(m.group(2)))
                continue
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
            if m and cur:
                cnt[cur] += 1
                ops[m.group(1)] += 1
        print(name[:70], sum(cnt.values()), "instructions")
        byfile = collections.Counter()
        for (fn, l), c in cnt.items():
            byfile[fn] += c
        print(" by file:", byfile.most_common())
        for (fn, l), c in sorted(cnt.items(), key=lambda kv: -kv[1])[:30]:
            print(f"   {c:5d}  {fn}:{l}")
        print(" opcodes:", ops.most_common(14))
