"""Per-kernel SASS mnemonic counts of libfeddb200.so (evidence for the instruction claims in DESIGN.md):
    python tools/sass_grep.py > profiles/r02_sass_grep.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "feddlib_b200", "libfeddb200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
want = ["UBLKCP", "UBLKPF", "LDGSTS", "STG.E.ENL2.256", "STG.E.256", "STG.E.128", "STG.E.64", "LDG.E.ENL2.256", "LDG.E.256", "LDG.E.128", "LDG.E.64", "LDS.128", "LDS.64", "STS.128", "STS.64",
        "DFMA", "DMUL", "DADD", "SHFL", "RED.E.ADD.F64", "ATOM", "CCTL", "BAR.SYNC", "WARPSYNC"]
name, counts, total = None, None, 0
rows = []
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        if name: rows.append((name, total, counts))
        name, counts, total = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        total += 1
        op = m.group(1)
        for w in want:
            if op.startswith(w):
                counts[w] += 1
                break
if name: rows.append((name, total, counts))
dem = subprocess.run(["cu++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}: static instruction counts per kernel (sm_100a)")
for (n, t, c), d in zip(rows, dem):
    short = re.sub(r"\([^()]*\)$", "", d.strip()).replace("void ", "").replace("fb::", "").replace("(int)", "").replace("(OpX)", "")
    if t < 40: continue
    print(f"{short:46s} total {t:6d}  " + "  ".join(f"{k}={v}" for k, v in sorted(c.items()) if v))
