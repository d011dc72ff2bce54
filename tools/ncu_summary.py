"""Dumps the headline metrics of every kernel in an .ncu-rep (read with `ncu -i ... --page raw --csv`)."""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "smsp__inst_executed.sum", "sm__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu.sum",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# {path}")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"## {d.get('Kernel Name', '?')}  (launch id {d.get('ID', '?')})")
        for k in hdr:
            if any(k == key or k.startswith(key) for key in KEYS):
                u = units[hdr.index(k)]
                print(f"{k} [{u}] = {d[k]}")


if __name__ == "__main__":
    main(sys.argv[1])
