"""profiles/r02_ptxas_resources.md: registers, static shared memory, spills and barriers of every kernel in libb200gs.so as
ptxas reports them (`nvcc -Xptxas -v`, the flags of build.py), plus the resident warps per SM those numbers allow on
sm_100a (64 K registers, 64 warps, 32 CTAs, 228 KB of shared memory per SM; registers are allocated per warp in units of
256).  Block sizes come from the kernels' __launch_bounds__; dynamic shared memory is not known to ptxas, the launch sites'
sizes are listed by hand below.  No GPU needed.

    python tools/ptxas_table.py > profiles/r02_ptxas_resources.md
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "3d-gaussian-splatting-for-novel-view-synthesis_b200", "build.py")

CSRC = os.path.join(ROOT, "3d-gaussian-splatting-for-novel-view-synthesis_b200", "csrc")
# dynamic shared memory of the launch sites (ptxas cannot know it): bytes, or a note
DYNAMIC = {
    "preprocess_fwd_tma_kernel": 2 * 30 * 1024,        # 2 x sizeof(PreStage), preprocess.cu
    "preprocess_bwd_tma_kernel": "2 x sizeof(PreStageBwd)",
    "route_write_kernel": 768 * 64,                    # kRouteStageBytes, route.cu
    "onesweep_pass2_kernel": "pass2_smem_bytes<B, T, I>()",
}


def block_sizes():
    """kernel name -> threads per CTA, from the __launch_bounds__ of csrc/*.cu and the constants they name."""
    consts, bounds = {}, {}
    for fn in sorted(os.listdir(CSRC)):
        if not fn.endswith((".cu", ".cuh")):
            continue
        src = open(os.path.join(CSRC, fn)).read()
        for m in re.finditer(r"constexpr int (k\w+) = (\d+);", src):
            consts[m.group(1)] = int(m.group(2))
        for m in re.finditer(r"__launch_bounds__\(([^,)]+)[^)]*\)+\s*(\w+)\(", src):
            bounds[m.group(2)] = m.group(1).strip()
    return {k: consts.get(v.split("/")[0].strip(), v) for k, v in bounds.items()}


def demangle(names):
    try:
        out = subprocess.run(["cu++filt"], input="\n".join(names), stdout=subprocess.PIPE, text=True, check=True).stdout
        return out.splitlines()
    except (OSError, subprocess.CalledProcessError):
        return names


def short(name):
    name = re.sub(r"^void ", "", name)
    depth = 0
    for i, ch in enumerate(name):                  # drop the argument list: the first "(" outside the template arguments
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            name = name[:i]
            break
    return name.replace("gs::", "").replace("(bool)1", "true").replace("(bool)0", "false").replace("(int)", "")


def warps_per_sm(regs, threads, smem):
    warps_cta = (threads + 31) // 32
    regs_warp = ((regs * 32 + 255) // 256) * 256
    by_regs = (65536 // regs_warp) // warps_cta
    by_warps = 64 // warps_cta
    by_smem = (228 * 1024) // (smem + 1024) if smem is not None else by_warps     # 1 KB reserved per CTA
    ctas = max(0, min(by_regs, by_warps, by_smem, 32))
    return ctas, ctas * warps_cta


def main():
    log = subprocess.run([sys.executable, BUILD, "--force", "--verbose"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         text=True, check=True).stdout
    rows = []
    it = iter(log.splitlines())
    cur = None
    for line in it:
        m = re.search(r"Compiling entry function '(\S+)' for '(sm_\w+)'", line)
        if m:
            cur = {"mangled": m.group(1), "arch": m.group(2), "stack": 0, "spill_st": 0, "spill_ld": 0}
            continue
        if cur is None:
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            cur["stack"], cur["spill_st"], cur["spill_ld"] = map(int, m.groups())
            continue
        m = re.search(r"Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes smem)?", line)
        if m:
            cur["regs"] = int(m.group(1))
            cur["bars"] = int(m.group(2) or 0)
            cur["smem"] = int(m.group(3) or 0)
            rows.append(cur)
            cur = None
    names = demangle([r["mangled"] for r in rows])
    for r, n in zip(rows, names):
        r["name"] = short(n)
    rows.sort(key=lambda r: r["name"])
    archs = sorted({r["arch"] for r in rows})
    print("# ptxas resource usage of libb200gs.so (`nvcc -Xptxas -v`, flags of build.py)\n")
    print(f"{len(rows)} kernels, target(s): {', '.join(archs)}.  `resident` = CTAs x warps per SM that registers, shared")
    print("memory (static + the dynamic size of the launch site where it is listed in tools/ptxas_table.py) and the 64-warp")
    print("limit allow, for the block size in the kernel's `__launch_bounds__`.\n")
    spills = [r for r in rows if r["spill_st"] or r["spill_ld"]]
    print(f"Kernels with register spills: {len(spills)}" + ("" if not spills else
          " (" + ", ".join(f"`{r['name']}` {r['spill_st']}+{r['spill_ld']} B" for r in spills) + ")") + ".\n")
    print("| kernel | registers | static smem (B) | dynamic smem (B) | barriers | stack (B) | spill st/ld (B) | threads | resident CTAs x warps |")
    print("|---|---|---|---|---|---|---|---|---|")
    sizes = block_sizes()
    for r in rows:
        base = re.sub(r"<.*$", "", r["name"]).split("::")[-1]
        targs = [a.strip() for a in re.sub(r"^[^<]*<|>$", "", r["name"]).split(",")] if "<" in r["name"] else []
        threads = sizes.get(base, 256)
        if base == "onesweep_pass2_kernel":
            threads = int(targs[1])
        elif base == "split_super_kernel":
            threads = int(targs[1])
        elif base == "blend_bwd_ppl_kernel":
            threads = 256 // int(targs[0])
        dyn = DYNAMIC.get(base, 0)
        ctas, warps = warps_per_sm(r["regs"], threads, r["smem"] + dyn if isinstance(dyn, int) else None)
        note = "" if isinstance(dyn, int) else " (registers / warps only)"
        print(f"| `{r['name']}` | {r['regs']} | {r['smem']} | {dyn} | {r['bars']} | "
              f"{r['stack']} | {r['spill_st']}/{r['spill_ld']} | {threads} | {ctas} x {warps // max(ctas, 1)} = {warps}{note} |")

if __name__ == "__main__":
    main()
