"""Per-kernel SASS evidence of libinsider_b200.so: counts of the mnemonics that prove the sm_100a features each kernel claims
(DMMA = FP64 tensor-core MMA, UBLKCP = cp.async.bulk / TMA 1-D bulk copy, UTMALDG = cp.async.bulk.tensor (tensor-map TMA), SYNCS = mbarrier, LDS/STS, DFMA/DADD/DMUL, SHFL, BAR,
UCGABAR = cluster barrier) plus registers / spills from `ptxas -v`. Usage: python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "insider_b200", "lib", "libinsider_b200.so")
MNEMONICS = ["DMMA", "UBLKCP", "UTMALDG", "SYNCS", "LDS", "STS", "LDGSTS", "DFMA", "DADD", "DMUL", "SHFL", "BAR", "UCGABAR", "ATOM", "RED", "LDG", "STG", "BRA"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(name):
    name = re.sub(r"ib::\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*$", "", name)


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            if op.startswith("UCGABAR"):
                op = "UCGABAR"
            counts[cur]["_total"] += 1
            if op in MNEMONICS:
                counts[cur][op] += 1
    regs = {}
    for log in glob.glob(os.path.join(ROOT, "insider_b200", "csrc", "*.ptxas.log")):
        txt = open(log).read()
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers", txt):
            regs[m.group(1)] = (int(m.group(5)), int(m.group(3)), int(m.group(4)))
    dm = demangle(list(counts))
    ver = subprocess.run(["cuobjdump", "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-1]
    print(f"cuobjdump -sass insider_b200/lib/libinsider_b200.so  ({ver}); arch sm_100a only; registers / spills from nvcc -Xptxas -v (insider_b200/csrc/*.ptxas.log)")
    print("columns: instructions | regs | spill st/ld B | " + " ".join(MNEMONICS))
    rows = []
    for f, c in counts.items():
        r = regs.get(f, (0, 0, 0))
        rows.append((short(dm.get(f, f)), c["_total"], r, [c[m] for m in MNEMONICS]))
    for name, tot, r, cs in sorted(rows):
        print(f"{name[:86]:86s} {tot:6d} | {r[0]:3d} | {r[1]:4d}/{r[2]:<4d} | " + " ".join(f"{v:5d}" for v in cs))
    fams = collections.Counter()
    for name, tot, r, cs in rows:
        fam = re.sub(r"<.*$", "", name)
        for m, v in zip(MNEMONICS, cs):
            if v:
                fams[(fam, m)] = max(fams[(fam, m)], v)
    print("\nfeature evidence per kernel family (max over template instances):")
    for fam in sorted({f for f, _ in fams}):
        print(f"  {fam:28s} " + ", ".join(f"{m} {fams[(fam, m)]}" for m in MNEMONICS if fams[(fam, m)] and m in ("DMMA", "UBLKCP", "SYNCS", "UCGABAR", "DFMA", "LDGSTS")))


if __name__ == "__main__":
    main()
