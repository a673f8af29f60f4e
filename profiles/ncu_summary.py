"""Summarise `ncu --page raw --csv` exports: python profiles/ncu_summary.py file_raw.csv [...]  -> markdown table rows."""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__inst_executed.sum", "warp inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]
STALL = "smsp__average_warps_issue_stalled_"


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        out.append({h: (v, u) for h, v, u in zip(hdr, r, units)})
    return out


def fmt(v, u):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    if u in ("ns", "nsecond"):
        return "%.3f ms" % (x / 1e6)
    if u in ("usecond", "us"):
        return "%.3f ms" % (x / 1e3)
    if u in ("msecond", "ms"):
        return "%.3f ms" % x
    if u in ("byte", "Kbyte", "Mbyte", "Gbyte"):
        x *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        return "%.3g MB" % (x / 1e6)
    if abs(x) >= 1e6:
        return "%.3g" % x
    return ("%.1f" % x) if x != int(x) else str(int(x))


for path in sys.argv[1:]:
    for k in load(path):
        name = k.get("Kernel Name", ("?", ""))[0]
        print("### %s  (%s)" % (path.split("/")[-1], name[:80]))
        cells = []
        for key, label in KEYS:
            if key in k:
                cells.append("%s %s" % (label, fmt(*k[key])))
        print(" | ".join(cells))
        st = []
        for h, (v, u) in k.items():
            if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                try:
                    st.append((float(v), h[len(STALL):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        st.sort(reverse=True)
        print("stalls per issue: " + ", ".join("%s %.2f" % (n, x) for x, n in st[:6]))
        print()
