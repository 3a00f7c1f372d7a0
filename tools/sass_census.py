"""SASS census of lib/libfa_sm100a.so: per kernel, the Blackwell-native mnemonics (UTC*MMA = tcgen05.mma,
LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tile load, SYNCS = mbarrier, MUFU.EX2) and the resource usage
(registers, spill stack, static shared memory).  Needs only cuobjdump, no GPU.

    python tools/sass_census.py > profiles/<round>_sass_census.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "flashattention.jl_b200", "lib", "libfa_sm100a.so")
PATTERNS = [("UTC*MMA", r"\bUTC[A-Z]*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"),
            ("UTCBAR", r"\bUTCBAR"), ("SYNCS", r"\bSYNCS"), ("MUFU.EX2", r"\bMUFU\.EX2"), ("HMMA", r"\bHMMA"),
            ("LDG", r"\bLDG"), ("STG", r"\bSTG"), ("LDL/STL", r"\b(LDL|STL)"), ("total", r"^\s+/\*[0-9a-f]{4}\*/")]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(name):
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("fa::", "")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            usage[cur] = dict(re.findall(r"(REG|STACK|SHARED|LOCAL):(\d+)", line))
            cur = None
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for key, pat in PATTERNS:
            if re.search(pat, line):
                counts[cur][key] += 1
    dm = demangle(list(counts))
    want = sys.argv[1] if len(sys.argv) > 1 else r"tc_|band|ring|merge|circ2d|softmax"
    keys = [k for k, _ in PATTERNS]
    print("# SASS census of lib/libfa_sm100a.so (cuobjdump -sass / -res-usage, sm_100a)\n")
    print("`UTC*MMA` = tcgen05.mma, `LDTM`/`STTM` = tcgen05.ld/st, `UTMALDG` = cp.async.bulk.tensor, `SYNCS` = mbarrier ops,")
    print("`HMMA` = legacy mma.sync (must be 0).  Static counts of instructions in the kernel body, not executed counts.")
    print("`LDL/STL` in the tc_* kernels are the argument blocks of the mbarrier-watchdog `printf` (cold path) plus one")
    print("dynamically indexed two-entry tile-range array per issuer warp -- not register spills (ptxas -v: 0 spill bytes")
    print("except the experimental SPLIT = 2 / 3 forward variants, which are off by default).\n")
    print("| kernel | regs | stack B | " + " | ".join(keys) + " |")
    print("|---|---|---|" + "---|" * len(keys))
    for fn, c in sorted(counts.items(), key=lambda kv: short(dm[kv[0]])):
        name = short(dm[fn])
        if not re.search(want, name):
            continue
        u = usage.get(fn, {})
        print(f"| `{name}` | {u.get('REG', '?')} | {u.get('STACK', '?')} | " + " | ".join(str(c[k]) for k in keys) + " |")


if __name__ == "__main__":
    main()
