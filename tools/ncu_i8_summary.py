"""Markdown summary of an `ncu --set full` capture for the kernels of the integer metric route (run here, no GPU needed):
  ncu -i gpurun_out/s5_prof.ncu-rep --page raw --csv > /tmp/raw.csv && python tools/ncu_i8_summary.py /tmp/raw.csv"""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2 -> SM bytes"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (of SM active cycles)"),
    ("sm__ops_path_tensor_op_utcimma_src_int8_sparsity_off.avg.pct_of_peak_sustained_elapsed", "UTCIMMA int8 ops % of peak"),
    ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "TMEM pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__cycles_active.avg", "SM active cycles (avg)"),
    ("sm__cycles_elapsed.max", "elapsed cycles"),
]


def main(path):
    rows = list(csv.reader(open(path)))
    H, U = rows[0], rows[1]
    seen = set()
    for r in rows[2:]:
        name = r[H.index("Kernel Name")].split("(")[0]
        if name in seen:
            continue
        seen.add(name)
        print(f"### `{name}`\n")
        print("| metric | value |\n|---|---|")
        for key, label in KEYS:
            if key in H and r[H.index(key)] not in ("", "n/a"):
                print(f"| {label} | {r[H.index(key)]} {U[H.index(key)]} |")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
