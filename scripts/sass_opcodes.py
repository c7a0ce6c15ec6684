"""cuobjdump -sass opcode histogram per kernel of deep-fusion_b200/lib/libdfcuda.so -> profiles/r02_sass_opcodes.txt
(no GPU needed).  Lists, per kernel, the tensor-core / TMEM / TMA / cluster opcodes that prove the Blackwell path
(UTCIMMA = tcgen05.mma kind::i8, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA loads / stores, UTCBAR = tcgen05.commit)
and the ten most frequent opcodes."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "deep-fusion_b200", "lib", "libdfcuda.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kern, hist = None, {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
KEY = ("UTCIMMA", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCATOMSWS", "SYNCS", "F2IP", "I2FP", "FADD2", "FMUL2", "LDL", "STL")
out = [f"# opcode histogram of {os.path.relpath(so, ROOT)} (cuobjdump -sass; scripts/sass_opcodes.py)", ""]
tot = collections.Counter()
for k in sorted(hist):
    h = hist[k]
    if not h:
        continue
    fam = collections.Counter()
    for op, c in h.items():
        base = op.split(".")[0]
        if base in KEY:
            fam[op if base in ("UTCIMMA", "UTMALDG", "UTMASTG", "UTCBAR", "LDTM", "STTM") else base] += c
            tot[op if base in ("UTCIMMA", "UTMALDG", "UTMASTG", "UTCBAR") else base] += c
    out.append(k)
    out.append(f"  instructions: {sum(h.values())}")
    out.append("  blackwell / epilogue opcodes: " + ", ".join(f"{op} {c}" for op, c in sorted(fam.items())))
    base = collections.Counter()
    for op, c in h.items():
        base[op.split(".")[0]] += c
    out.append("  top: " + ", ".join(f"{op} {c}" for op, c in base.most_common(10)))
    out.append("")
out.append("TOTAL over all kernels: " + ", ".join(f"{op} {c}" for op, c in sorted(tot.items())))
open(os.path.join(ROOT, "profiles", "r02_sass_opcodes.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[-1:]))
print(len(hist), "kernels")
