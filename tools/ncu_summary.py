"""Summarise an .ncu-rep (read here with `ncu -i ... --page raw --csv`) into the text kept under profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rN_ncu_x.txt
"""
import csv, io, subprocess, sys

KEYS = [
    "gpu__time_duration.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    for d in data:
        print(f"==== {d[col['Kernel Name']][:110]}")
        tot = 0.0
        for k in KEYS:
            if k in col and d[col[k]] not in ("", "n/a"):
                print(f"  {k} [{units[col[k]]}] = {d[col[k]]}")
                if k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    tot += float(d[col[k]].replace(",", "")) * UNIT.get(units[col[k]], 1.0)
        stalls = []
        for h, i in col.items():
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(d[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("  top stalls (warps per issue-active cycle): " + ", ".join(f"{n}={v:.2f}" for v, n in stalls[:5]))
        print(f"  traffic (dram read+write) = {tot / 1e6:.1f} MB")


if __name__ == "__main__":
    main(sys.argv[1])
