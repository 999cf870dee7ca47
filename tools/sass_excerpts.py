"""profiles/r02_sass_excerpts.md: per kernel of libb200gs.so, how often the SASS mnemonics that carry the design occur
(bulk-copy engine, mbarrier, multimem, reductions to global memory, MUFU, warp shuffles / votes / match), plus the first
occurrence of each as an excerpt.  cuobjdump -sass on the built library; no GPU needed.

    python tools/sass_excerpts.py > profiles/r02_sass_excerpts.md
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "3d-gaussian-splatting-for-novel-view-synthesis_b200", "b200gs", "libb200gs.so")
PATTERNS = collections.OrderedDict([
    ("UBLKCP (cp.async.bulk: TMA 1-D bulk copy)", r"\bUBLKCP"),
    ("SYNCS (mbarrier arrive / try_wait)", r"\bSYNCS"),
    ("MULTIMEM (NVLS ld_reduce / st)", r"LDGMC|REDGMC|STGMC"),
    ("REDG / RED (reduction to global)", r"\bREDG|\bRED\."),
    ("ATOMG (global atomics with return)", r"\bATOMG"),
    ("ATOMS (shared atomics)", r"\bATOMS"),
    ("MUFU.EX2", r"MUFU\.EX2"),
    ("MUFU (other)", r"MUFU\.(?!EX2)"),
    ("SHFL", r"\bSHFL"),
    ("VOTE / VOTEU", r"\bVOTEU?\b"),
    ("MATCH", r"\bMATCH"),
    ("REDUX", r"\bREDUX"),
    ("LDS.128", r"LDS(\.U)?\.128"),
    ("LDG.E.128 / .CONSTANT", r"LDG\.E(\.\w+)*\.128"),
    ("STG.E.128", r"STG\.E(\.\w+)*\.128"),
    ("BAR.SYNC", r"\bBAR\."),
    ("FFMA", r"\bFFMA"),
])


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    arch = set(re.findall(r"arch = (sm_\w+)", out))
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = []
        elif cur and re.match(r"\s*/\*[0-9a-f]{4}\*/", line):
            kernels[cur].append(line.strip())
    dem = subprocess.run(["c++filt"] + list(kernels), stdout=subprocess.PIPE, text=True).stdout.splitlines()
    print("# SASS evidence, libb200gs.so (round 2)\n")
    print(f"`cuobjdump -sass` of the built library; cubin architectures found: {sorted(arch)}.  Counts are static instruction "
          "counts per kernel.\n")
    names = list(PATTERNS)
    print("| kernel | instr | " + " | ".join(n.split(" ")[0] for n in names) + " |")
    print("|---|---|" + "---|" * len(names))
    first = {}
    for (mangled, lines), name in zip(kernels.items(), dem):
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("gs::", "").replace("(anonymous namespace)::", "")
        counts = []
        for label, pat in PATTERNS.items():
            hits = [ln for ln in lines if re.search(pat, ln)]
            counts.append(len(hits))
            if hits and (label, short) not in first and len([k for k in first if k[0] == label]) < 2:
                first[(label, short)] = hits[0]
        print(f"| `{short[:60]}` | {len(lines)} | " + " | ".join(str(c) if c else "" for c in counts) + " |")
    print("\n## First occurrences\n")
    for (label, short), ln in first.items():
        print(f"* **{label}** in `{short[:60]}`: `{re.sub(r'/\\*[0-9a-f]+\\*/', '', ln).strip()[:110]}`")


if __name__ == "__main__":
    main()
