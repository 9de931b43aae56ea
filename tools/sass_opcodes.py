"""Per-kernel counts of the SASS opcodes that prove the hardware paths (B200_PROFILING.md): FP64 tensor-core MMA
(DMMA), TMA tensor loads (UTMALDG), bulk async copies (UBLKCP), mbarrier traffic (SYNCS), register re-allocation
(USETMAXREG), and -- for the record -- the tcgen05 family (UTCMMA / LDTM / STTM: absent, tcgen05 has no f64 kind).

    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt          (needs cuobjdump; no GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "auto_oo_b200", "_lib", "liboo_b200.so")
OPS = ["DMMA", "UTMALDG", "UBLKCP", "SYNCS", "USETMAXREG", "UTCMMA", "LDTM", "STTM", "HMMA", "LDS", "STS", "LDG", "STG"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            if op in OPS:
                counts[cur][op] += 1
    names = demangle(list(counts))
    total = collections.Counter()
    print(f"# SASS opcode counts per kernel of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a)")
    print("# " + "  ".join(f"{o:>10s}" for o in OPS) + "  kernel")
    for k, c in counts.items():
        total.update(c)
        name = names[k].replace("oo::", "").replace("(anonymous namespace)::", "")
        name = re.sub(r"\((?:const |double|int|long|unsigned|void|CUtensorMap|oo|TnArgs|TriArgs|RdmView|TView|FusedExpmArgs).*", "", name) or k
        print("  " + "  ".join(f"{c.get(o, 0):10d}" for o in OPS) + "  " + name)
    print("# " + "  ".join(f"{total.get(o, 0):10d}" for o in OPS) + "  TOTAL")


if __name__ == "__main__":
    sys.exit(main())
