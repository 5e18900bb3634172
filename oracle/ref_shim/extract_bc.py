"""Slices the Dirichlet boundary-condition members of the reference's BCBuilder (core/General/BCBuilder_def.hpp, read where
it lies, never copied into the repository) into oracle/_ref/bc_subset.inc, which bc_driver.cpp includes: what SURVEY.md
section 8(f) rank 1 names -- setSystem / setDirichletBC / setLocalRowOne / setLocalRowZero and setRHS -- with the helpers they
call."""
import re
import sys

src_path, out_path = sys.argv[1], sys.argv[2]
src = open(src_path, encoding="utf-8", errors="replace").read()
WANT = ["BCBuilder", "addBC", "setRHS", "setDirichletBoundaryFromExternal", "blockHasDirichletBC", "findFlag", "setSystem",
        "setDirichletBC", "setLocalRowOne", "setLocalRowZero"]
out, count = [], {}
for m in re.finditer(r"template\s*<class SC,\s*class LO,\s*class GO,\s*class NO>\s*\n[^\n;{]*?BCBuilder<SC,LO,GO,NO>::(\w+)\s*\(", src):
    name = m.group(1)
    if name not in WANT:
        continue
    i = src.index("(", m.end() - 1)
    depth = 0
    while True:
        depth += {"(": 1, ")": -1}.get(src[i], 0)
        if depth == 0:
            break
        i += 1
    j = src.index("{", i)
    depth, k = 0, j
    while True:
        depth += {"{": 1, "}": -1}.get(src[k], 0)
        if depth == 0:
            break
        k += 1
    line = src.count("\n", 0, m.start()) + 1
    out.append(f"// ---- BCBuilder_def.hpp:{line} {name}\n#line {line} \"{src_path}\"\n" + src[m.start():k + 1] + "\n")
    count[name] = count.get(name, 0) + 1
missing = [w for w in WANT if w not in count]
if missing:
    sys.exit(f"extract_bc.py: not found in {src_path}: {missing}")
open(out_path, "w").write("namespace FEDD {\n" + "\n".join(out) + "\n} // namespace FEDD\n")
print("extracted", sum(count.values()), "definitions:", count)
