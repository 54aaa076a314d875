"""Key metrics per launch from an .ncu-rep (ncu --set full): python tools/ncu_summary.py rep.ncu-rep"""
import csv
import subprocess
import sys

WANT = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_utc_cycles_active.avg.pct_of_peak_sustained_active" if False else "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    tensor_cols = ["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
                   "sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
                   "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
                   "lts__t_sectors_srcunit_tex.sum", "lts__t_bytes.sum.per_second", "dram__bytes.sum.per_second",
                   "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
                   # operand feed of the tcgen05 kernels: bytes TMA brought into the SMs, and the tensor core's
                   # shared-memory read wavefronts
                   "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
                   "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum.per_second",
                   "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum.pct_of_peak_sustained_elapsed",
                   "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]
    print(f"# {rep}: {len(data)} launches")
    for name in WANT + [c for c in tensor_cols if c not in WANT]:
        if name not in hdr:
            continue
        i = hdr.index(name)
        print(f"{name} [{units[i]}]: " + " | ".join(r[i] for r in data))


if __name__ == "__main__":
    main()
