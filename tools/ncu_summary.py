"""Summarise an .ncu-rep (read on the CPU box) into the few numbers the roofline needs.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.md]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "kernel duration"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__cycles_active.avg", "SM active cycles"),
    ("sm__cycles_elapsed.avg", "SM elapsed cycles"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe % (TPC triage)"),
    ("sm__inst_executed_pipe_fma.sum", "FMA-pipe instructions"),
    ("sm__inst_executed_pipe_fma_realtime.avg.pct_of_peak_sustained_elapsed", "FMA pipe % (triage)"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts (LSU)"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts (LSU)"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        name = d.get("Kernel Name", ("?", ""))[0]
        out.append(f"### {name}\n")
        out.append("| metric | value | unit |\n|---|---|---|")
        for key, label in KEYS:
            hit = [h for h in hdr if h.endswith(key)]
            for h in hit[:1]:
                v, u = d[h]
                if v != "":
                    out.append(f"| {label} (`{key}`) | {v} | {u} |")
        out.append("")
    text = "\n".join(out)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
