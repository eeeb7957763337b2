"""Print the key metrics of an `ncu --set full` report exported with `ncu -i X.ncu-rep --page raw --csv`."""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active", "sm__inst_executed_pipe_tensor",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit", "sm__cycles_elapsed.max",
        "smsp__inst_executed.sum", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared",
        "smsp__pcsamp_warps_issue_stalled", "sm__mio", "mufu", "xu"]


def main(path):
    rows = list(csv.reader(open(path, newline="")))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("=" * 100)
        print(d.get("Kernel Name", "?")[:90], "| grid", d.get("Grid Size"), "block", d.get("Block Size"))
        for h, u in zip(hdr, units):
            if any(k in h for k in KEYS):
                v = d[h]
                if v not in ("", "n/a"):
                    print(f"  {h} = {v} {u}")


if __name__ == "__main__":
    main(sys.argv[1])
