"""Summarise .ncu-rep files (ncu --set full) as one markdown table row per captured launch, and optionally record the
measured DRAM traffic per launch of each kernel in a JSON file that bench.py reads for `roofline.traffic`.

usage: python tools/ncu_summary.py [--traffic-json profiles/r02_ncu_traffic.json] gpurun_out/prof_x.ncu-rep [more.ncu-rep ...]

Tensor column: tcgen05.mma (SASS UTCHMMA) is counted by sm__ops_path_tensor_op_utchmma_*; the classic
sm__pipe_tensor_cycles_active / sm__inst_executed_pipe_tensor counters only see the legacy HMMA path, so both are printed."""
import csv
import io
import json
import os
import subprocess
import sys

COLS = [
    ("grid", "Grid Size"), ("block", "Block Size"), ("us", "gpu__time_duration.sum"),
    ("dram_rd_MB", "dram__bytes_read.sum"), ("dram_wr_MB", "dram__bytes_write.sum"),
    ("dram_%", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l1_%", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("sm_%", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("utchmma_f16_%", "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed"),
    ("utchmma_bf16_%", "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed"),
    ("hmma_%", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active"),
    ("occ_%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"), ("smem_dyn", "launch__shared_mem_per_block_dynamic"),
    ("sm_active_cyc", "sm__cycles_active.avg"), ("elapsed_cyc", "sm__cycles_elapsed.max"),
]


def to_mb(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0)


def main(argv):
    traffic_path = None
    if argv and argv[0] == "--traffic-json":
        traffic_path, argv = argv[1], argv[2:]
    traffic = {}
    for path in argv:
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
        print("| kernel | " + " | ".join(lab for lab, _ in ix) + " |")
        print("|---|" + "---|" * len(ix))
        for r in rows[2:]:
            cells, vals = [], {}
            for lab, i in ix:
                if i is None:
                    cells.append("n/a")
                    continue
                v = r[i]
                if lab.endswith("_MB"):
                    v = f"{to_mb(v, units[i]):.2f}"
                elif lab == "us":
                    v = f"{float(v.replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(units[i], 1.0):.2f}"
                cells.append(v)
                vals[lab] = v
            short = r[kname].split("(")[0].replace("void ", "").replace("smb::", "")
            print("| " + short[:48] + " | " + " | ".join(cells) + " |")
            try:
                key = short.split("<")[0]
                e = traffic.setdefault(key, {"launches": 0, "dram_bytes": 0.0, "us": 0.0, "grids": []})
                e["launches"] += 1
                e["dram_bytes"] += (float(vals["dram_rd_MB"]) + float(vals["dram_wr_MB"])) * 1e6
                e["us"] += float(vals["us"])
                if vals["grid"] not in e["grids"]:
                    e["grids"].append(vals["grid"])
            except (KeyError, ValueError):
                pass
        print()
    if traffic_path:
        old = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
        for k, e in traffic.items():
            old[k] = {"dram_bytes_per_launch": e["dram_bytes"] / e["launches"], "launches_captured": e["launches"],
                      "us_per_launch_under_ncu": e["us"] / e["launches"], "grids": e["grids"],
                      "source": "ncu --set full --clock-control none: " + ", ".join(os.path.basename(p) for p in argv)}
        json.dump(old, open(traffic_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main(sys.argv[1:])
