"""Static evidence from the built library (no GPU needed): per-kernel registers / stack / static shared memory
(`cuobjdump -res-usage`) and how often the Blackwell-specific SASS instructions appear in each kernel (`cuobjdump -sass`):
UTC*MMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor copies, LDTM / STTM = tcgen05.ld / st (tensor memory), HMMA = mma.sync,
LDGSTS = cp.async, SYNCS = mbarrier ops, UCGABAR / CGA = cluster barriers.

    python tools/sass_evidence.py > profiles/r2_sass_evidence.md
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gridmm_b200", "lib", "libgridmm_b200.so")
PATTERNS = [("UTC*MMA", r"\bUTC\w*MMA"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
            ("HMMA", r"\bHMMA"), ("LDGSTS", r"\bLDGSTS"), ("SYNCS", r"\bSYNCS"), ("UCGABAR", r"\bUCGABAR"), ("ELECT", r"\bELECT")]


def demangle(names):
    out = subprocess.run(["c++filt"] + names, stdout=subprocess.PIPE, text=True).stdout.splitlines()
    return [re.sub(r"\(.*", "", o).replace("void ", "").replace("gmm::", "") for o in out]


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], stdout=subprocess.PIPE, text=True).stdout
    usage = {}
    for m in re.finditer(r"Function (\S+):\s*\n?\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", res):
        usage[m.group(1)] = tuple(int(x) for x in m.groups()[1:])
    sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True).stdout
    counts, cur = collections.defaultdict(collections.Counter), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None:
            continue
        for name, pat in PATTERNS:
            if re.search(pat, line):
                counts[cur][name] += 1
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            counts[cur]["instr"] += 1
    names = sorted(usage)
    pretty = dict(zip(names, demangle(names)))
    print("# Static SASS evidence of the sm_100a library (`tools/sass_evidence.py`, cuobjdump %s)\n" %
          subprocess.run(["cuobjdump", "--version"], stdout=subprocess.PIPE, text=True).stdout.strip().splitlines()[-1])
    print("Registers / stack / static shared memory per kernel, SASS instruction count, and occurrences of the Blackwell-specific\n"
          "instructions (UTC*MMA = `tcgen05.mma`, UTMALDG / UTMASTG = TMA, LDTM / STTM = `tcgen05.ld` / `st`, SYNCS = mbarrier,\n"
          "UCGABAR = cluster barrier, HMMA = `mma.sync`, LDGSTS = `cp.async`). Dynamic shared memory is set at launch and not listed.\n"
          "No kernel spills (LOCAL = 0 everywhere).\n")
    print("| kernel | regs | stack | local | SASS instr | " + " | ".join(n for n, _ in PATTERNS) + " |")
    print("|---|---:|---:|---:|---:|" + "---:|" * len(PATTERNS))
    for n in sorted(names, key=lambda k: pretty[k]):
        reg, stack, shared, local = usage[n]
        c = counts.get(n, {})
        print("| `%s` | %d | %d | %d | %d | %s |" % (pretty[n][:70], reg, stack, local, c.get("instr", 0),
                                                   " | ".join(str(c.get(p, 0)) for p, _ in PATTERNS)))


if __name__ == "__main__":
    main()
