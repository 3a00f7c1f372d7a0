"""Summarise .ncu-rep captures (tools/ncu_capture.sh) as one markdown table per report.
    python tools/ncu_summary.py gpurun_out/r1s_*.ncu-rep > profiles/r1s_ncu_summary.md   (needs ncu, no GPU)"""
import csv
import io
import re
import subprocess
import sys

METRICS = [
    ("duration", "gpu__time_duration.sum"),
    ("grid x block", None),
    ("registers / thread", "launch__registers_per_thread"),
    ("dynamic smem / block", "launch__shared_mem_per_block_dynamic"),
    ("dram read", "dram__bytes_read.sum"),
    ("dram write", "dram__bytes_write.sum"),
    ("dram throughput % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 hit rate %", "lts__t_sector_hit_rate.pct"),
    ("tensor pipe active % (of active cycles)", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("XU (MUFU) pipe %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    ("FMA pipe %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("ALU pipe %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("LSU pipe %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("issue slots busy %", "sm__inst_issued.avg.pct_of_peak_sustained_active"),
    ("warps active % of max", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("SM busy %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("shared-memory bank conflicts (ld / st)", None),
    ("local loads / stores (inst)", None),
]


def rows(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    hdr, units = r[0], r[1]
    return [dict(zip(hdr, zip(row, units))) for row in r[2:]]


def val(d, key):
    if key in d:
        v, u = d[key]
        try:
            v = f"{float(v):.4g}"
        except ValueError:
            pass
        return f"{v} {u}".strip()
    return "n/a"


def main():
    print("# ncu summaries (`ncu --set full --clock-control none --import-source on`, one B200, each command had exited 0")
    print("without ncu first; read with `ncu -i ... --page raw --csv`).  Cold-cache, serialised launches: compare")
    print("shares and pipe fractions, not absolute times.  The .ncu-rep files stay in gpurun_out/ (scratch).\n")
    for path in sys.argv[1:]:
        for d in rows(path):
            name = re.sub(r"\(.*", "", d["Kernel Name"][0]).replace("void unnamed>::", "").replace("void fa::", "")
            print(f"## `{name}` — {path.split('/')[-1]}\n")
            print("| metric | value |\n|---|---|")
            for label, key in METRICS:
                if label == "grid x block":
                    v = f"{val(d, 'launch__grid_size')} x {val(d, 'launch__block_size')}"
                elif label.startswith("shared-memory bank"):
                    v = f"{val(d, 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum')} / {val(d, 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum')}"
                elif label.startswith("local loads"):
                    v = f"{val(d, 'sass__inst_executed_local_loads')} / {val(d, 'sass__inst_executed_local_stores')}"
                else:
                    v = val(d, key)
                print(f"| {label} | {v} |")
            print()


if __name__ == "__main__":
    main()
