"""Slices the hot-path member functions out of the reference's FE_def.hpp (read where it lies, never
copied into the repository) into oracle/_ref/fe_subset.inc, which ref_driver.cpp includes.  Only the
routines SURVEY.md section 8(a) puts on the hot path are taken; everything else in the 10 kLoC file needs
Trilinos pieces that have no place in a checker."""
import re
import sys

src_path, out_path = sys.argv[1], sys.argv[2]
src = open(src_path, encoding="utf-8", errors="replace").read()
WANT = ["FE", "addFE", "applyBTinv", "assemblyMass", "assemblyBDStabilization", "assemblyStress", "assemblyRHS", "assemblyLaplace", "assemblyLaplaceVecField", "assemblyLinElasXDim",
        "assemblyAdvectionVecField", "assemblyAdvectionInUVecField", "assemblyDivAndDivT", "assemblyDivAndDivTFast",
        "epsilonTensor", "phi", "gradPhi", "buildTransformation", "determineDegree", "getQuadratureValues",
        "getPhi", "getPhiGlobal", "getDPhi", "checkFE"]
out = []
count = {}
for m in re.finditer(r"template\s*<class SC,\s*class LO,\s*class GO,\s*class NO>\s*\n[^\n;{]*?FE<SC,LO,GO,NO>::(\w+)\s*\(", src):
    name = m.group(1)
    if name not in WANT:
        continue
    # find the body: first '{' after the parameter list, then match braces
    i = src.index("(", m.end() - 1)
    depth = 0
    while True:
        if src[i] == "(":
            depth += 1
        elif src[i] == ")":
            depth -= 1
            if depth == 0:
                break
        i += 1
    j = src.index("{", i)
    depth, k = 0, j
    while True:
        if src[k] == "{":
            depth += 1
        elif src[k] == "}":
            depth -= 1
            if depth == 0:
                break
        k += 1
    line = src.count("\n", 0, m.start()) + 1
    out.append(f"// ---- FE_def.hpp:{line} {name}\n#line {line} \"{src_path}\"\n" + src[m.start():k + 1] + "\n")
    count[name] = count.get(name, 0) + 1
missing = [w for w in WANT if w not in count]
if missing:
    sys.exit(f"extract.py: not found in {src_path}: {missing}")
open(out_path, "w").write("namespace FEDD {\n" + "\n".join(out) + "\n} // namespace FEDD\n")
print("extracted", sum(count.values()), "definitions:", count)
