"""Slices BlockMatrix::merge and what it calls out of the reference's BlockMatrix_def.hpp / BlockMap_def.hpp (read where they lie,
never copied into the repository) into oracle/_ref/bm_subset.inc, which bm_driver.cpp includes."""
import re
import sys

out_path = sys.argv[-1]
jobs = [(sys.argv[1], "BlockMap", r"BlockMap<LO,GO,NO>::", r"template\s*<\s*class LO,\s*class GO,\s*class NO>",
         ["BlockMap", "~BlockMap", "addBlock", "merge", "getMergedMap"]),
        (sys.argv[2], "BlockMatrix", r"BlockMatrix<SC,LO,GO,NO>::", r"template\s*<class SC,\s*class LO,\s*class GO,\s*class NO>",
         ["BlockMatrix", "~BlockMatrix", "size", "blockExists", "addBlock", "merge", "determineLocalOffsets", "determineGlobalOffsets", "mergeBlockNew"])]
out = []
for src_path, cls, qual, tmpl, want in jobs:
    src = open(src_path, encoding="utf-8", errors="replace").read()
    count = {}
    for m in re.finditer(tmpl + r"\s*\n[^\n;{]*?" + qual + r"(~?\w+)\s*\(", src):
        name = m.group(1)
        if name not in want:
            continue
        head = src[m.start():src.index("{", m.end())]
        if cls == "BlockMatrix" and name == "BlockMatrix" and "BlockMatrixPtr_Type" in head:
            continue                                      # the copy constructor needs Matrix(MatrixPtr)
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
        out.append(f"// ---- {cls}_def.hpp:{line} {name}\n#line {line} \"{src_path}\"\n" + src[m.start():k + 1] + "\n")
        count[name] = count.get(name, 0) + 1
    missing = [w for w in want if w not in count]
    if missing:
        sys.exit(f"extract_bm.py: not found in {src_path}: {missing}")
    print(cls, "extracted", count)
open(out_path, "w").write("namespace FEDD {\nusing std::max;\n" + "\n".join(out) + "\n} // namespace FEDD\n")
