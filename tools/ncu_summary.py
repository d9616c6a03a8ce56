"""Condense `ncu --set full` reports into the small CSVs kept under profiles/ (the .ncu-rep files stay in gpurun_out/).

    python tools/ncu_summary.py gpurun_out/r2_fwd.ncu-rep [more.ncu-rep ...] > profiles/r2_xxx_ncu_full_summary.csv
"""
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.per_cycle_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__cluster_size",
]


def main():
    w = csv.writer(sys.stdout)
    first = True
    for path in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        ix = {h: i for i, h in enumerate(hdr)}
        cols = [k for k in KEEP if k in ix]
        if first:
            w.writerow(["report", "Kernel Name", "Grid Size", "Block Size"] + cols)
            w.writerow(["", "", "", ""] + [units[ix[k]] for k in cols])
            first = False
        for r in data:
            w.writerow([path.split("/")[-1], r[ix["Kernel Name"]][:60], r[ix["Grid Size"]], r[ix["Block Size"]]] +
                       [r[ix[k]] for k in cols])


if __name__ == "__main__":
    main()
