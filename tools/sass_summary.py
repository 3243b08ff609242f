"""Per-kernel SASS evidence of the built library (no GPU needed): which machine instructions the hand-written
kernels compile to.  usage: sass_summary.py [lib] > profiles/rNN_sass_summary.md"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(
    ROOT, "stm-multifrontal-qr-factorization-empowered-by-gcn_b200", "lib", "libstmqr_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
regs = {}
for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)", res):
    regs[m.group(1)] = (int(m.group(2)), int(m.group(3)), int(m.group(4)))
pats = [("DMMA", r"\bDMMA\b"), ("DFMA", r"\bDFMA\b"), ("DADD/DMUL", r"\bD(ADD|MUL)\b"), ("LDGSTS (cp.async)", r"\bLDGSTS"),
        ("LDS", r"\bLDS\b|\bLDS\."), ("STS", r"\bSTS\b|\bSTS\."), ("LDG", r"\bLDG\."), ("STG", r"\bSTG\."),
        ("SHFL", r"\bSHFL\."), ("BAR", r"\bBAR\."), ("cluster barrier (UCGABAR)", r"UCGABAR"),
        ("MUFU.RSQ64H", r"MUFU\.RSQ64H"), ("MUFU.RCP64H", r"MUFU\.RCP64H"), ("ATOM/RED", r"\b(ATOM|RED|ATOMG|REDG)\b|\bATOMG\.|\bRED\."),
        ("UTMALDG (TMA)", r"UTMALDG"), ("local (LDL/STL)", r"\b(LDL|STL)\b")]
parts = re.split(r"\n\s*Function : ", txt)
print("# SASS summary of libstmqr_b200.so (cuobjdump -sass, sm_100a)\n")
print("`c++filt` names; counts are static instruction counts in the kernel's SASS.\n")
print("| kernel | regs | stack | instrs | " + " | ".join(p[0] for p in pats) + " |")
print("|---|---|---|---|" + "---|" * len(pats))
rows = []
for p in parts[1:]:
    mangled = p.split("\n", 1)[0].strip()
    body = p
    name = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
    name = name.replace("(anonymous namespace)::", "").replace("stmqr::", "").replace("void ", "")
    name = re.sub(r"\(.*$", "", name).replace("(int)", "").replace("(bool)", "")
    n = len(re.findall(r"/\*[0-9a-f]{4,}\*/\s+[A-Z@]", body))
    r = regs.get(mangled, ("?", "?", "?"))
    rows.append((name, r[0], r[1], n, [len(re.findall(rx, body)) for _, rx in pats]))
for name, rg, stk, n, c in sorted(rows):
    print(f"| `{name}` | {rg} | {stk} | {n} | " + " | ".join(str(x) for x in c) + " |")
print("\nNo `UTMALDG`: the engine moves tiles with `cp.async` (`LDGSTS`, 16-byte where the front's `ld` allows) and keeps "
      "panel slabs resident in shared memory; FP64 has no `tcgen05.mma` kind, so the tensor path is warp-level `DMMA` "
      "(`mma.sync.m8n8k4.f64`).")
