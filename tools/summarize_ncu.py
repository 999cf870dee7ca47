"""Turn ncu outputs into the small tracked summaries under profiles/.
  python tools/summarize_ncu.py launches <launches.csv> <out.md>
  python tools/summarize_ncu.py full <raw.csv> <out.csv> [traffic.json]
"""
import collections
import csv
import json
import sys


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = [r for r in rows if "Kernel Name" in r][0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[rows.index(hdr) + 1:]:
        name = r[ki].split("(")[0]
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{k[:70]}` | {c} | {t:.1f} | {t / c:.1f} | {t / tot:.3f} |\n")


KEEP = ["ID", "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed_op_shared_atom.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum"]


# single-kernel regions of bench.py's per-kernel table (multi-kernel regions keep the figures of an earlier capture)
REGION_OF = {"blend_fwd_kernel": "blend_fwd", "blend_bwd_kernel": "blend_bwd", "blend_bwd_ppl_kernel": "blend_bwd", "preprocess_fwd_tma_kernel": "preprocess_fwd",
             "preprocess_bwd_tma_kernel": "preprocess_bwd", "l1_ssim_fwd_kernel": "l1_ssim_fwd",
             "l1_ssim_bwd_kernel": "l1_ssim_bwd", "adam_step_kernel": "adam_step", "scan_emit_super_kernel": "emit_super"}


def full(src, dst, traffic=None):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    tr = {}
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i][:100] for i in idx])
            name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("gs::", "")
            def val(k):
                v, u = float(r[hdr.index(k)]), units[hdr.index(k)]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            region = REGION_OF.get(name.split("<")[0])
            if region and region not in tr:
                tr[region] = {"dram_bytes": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                              "inst_executed": float(r[hdr.index("smsp__inst_executed.sum")])}
    if traffic:
        old = json.load(open(traffic)) if __import__("os").path.exists(traffic) else {}
        old.update(tr)
        old["_source"] = dst
        json.dump(old, open(traffic, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(*sys.argv[2:])
