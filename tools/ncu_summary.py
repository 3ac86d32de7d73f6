#!/usr/bin/env python
"""Summarise an .ncu-rep (read here with `ncu -i`, no GPU needed) into a small CSV/markdown for profiles/.
   usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rNN_name"""
import csv, io, subprocess, sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("dram__bytes.sum.per_second", "dram bytes per second"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (elapsed)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (active)"),
    ("sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.sum", "UTCHMMA tf32 ops (flop)"),
    ("sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.sum.per_second", "UTCHMMA tf32 rate"),
    ("sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.sum", "UTCHMMA bf16 ops (flop)"),
    ("sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.sum.per_second", "UTCHMMA bf16 rate"),
    ("sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.sum", "UTCHMMA fp16 ops (flop)"),
    ("sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.sum.per_second", "UTCHMMA fp16 rate"),
    ("sm__inst_executed_pipe_tmem.sum", "TMEM instructions"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread (launch)"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem LSU wavefronts % of peak"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global load sectors (LSU)"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "global store sectors (LSU)"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "global store requests (LSU)"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock"),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    names = [d[col["Kernel Name"]] for d in data]
    with open(out + ".csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(data))])
        w.writerow(["Kernel Name", ""] + names)
        for k, _ in KEYS:
            if k in col:
                w.writerow([k, units[col[k]]] + [d[col[k]] for d in data])
    with open(out + ".md", "w") as f:
        f.write(f"# ncu summary of `{rep.split('/')[-1]}` (`ncu --set full --clock-control none --import-source on`)\n\n")
        f.write("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |\n|---|---|" + "---|" * len(data) + "\n")
        f.write("| kernel | | " + " | ".join("`" + n.replace("void ", "").split("(")[0] + "`" for n in names) + " |\n")
        for k, label in KEYS:
            if k in col:
                f.write(f"| {label} (`{k}`) | {units[col[k]]} | " + " | ".join(d[col[k]] for d in data) + " |\n")
    print(open(out + ".md").read())


if __name__ == "__main__":
    main()
