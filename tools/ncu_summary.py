"""Summarise an .ncu-rep (ncu --set full) as one markdown table row per captured launch.
usage: python tools/ncu_summary.py gpurun_out/prof_gemm.ncu-rep [more.ncu-rep ...]"""
import csv, io, subprocess, sys

COLS = [
    ("grid", "Grid Size"), ("block", "Block Size"), ("us", "gpu__time_duration.sum"),
    ("dram_rd_MB", "dram__bytes_read.sum"), ("dram_wr_MB", "dram__bytes_write.sum"),
    ("dram_%", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l1_%", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("sm_%", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor_%", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
    ("occ_%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"), ("smem_dyn", "launch__shared_mem_per_block_dynamic"),
    ("sm_active_cyc", "sm__cycles_active.avg"), ("elapsed_cyc", "sm__cycles_elapsed.max"),
]

def to_mb(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0)

for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    def find(name):
        exact = [i for i, h in enumerate(hdr) if h == name]
        if exact:
            return exact[0]
        # section-prefixed duplicates (e.g. "FBSP.TriageCompute.<metric>"): take the first one that holds data
        cands = [i for i, h in enumerate(hdr) if h.endswith("." + name)]
        for i in cands:
            if len(rows) > 2 and rows[2][i] not in ("", "no data"):
                return i
        return cands[0] if cands else None
    ix = [(lab, find(name)) for lab, name in COLS]
    kname = find("Kernel Name")
    print(f"### {path}\n")
    print("| kernel | " + " | ".join(l for l, _ in ix) + " |")
    print("|---|" + "---|" * len(ix))
    for r in rows[2:]:
        cells = []
        for lab, i in ix:
            if i is None:
                cells.append("n/a"); continue
            v = r[i]
            if lab.endswith("_MB"):
                v = f"{to_mb(v, units[i]):.2f}"
            elif lab == "us":
                v = f"{float(v.replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(units[i], 1.0):.2f}"
            cells.append(v)
        print("| " + r[kname].split("(")[0].replace("void ", "")[:40] + " | " + " | ".join(cells) + " |")
    print()
