"""Key counters of an `ncu --set full` report as text (reads with `ncu -i ... --page raw --csv`, no GPU needed).
Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [...]"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_uniform.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct"]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader([ln for ln in out.splitlines() if ln.startswith('"')]))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            name = vals[hdr.index("Kernel Name")]
            print(f"== {path}\n   kernel: {name[:140]}")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    print(f"   {w:75s} {vals[i]:>16s} {units[i]}")
            try:
                rd = float(vals[hdr.index("dram__bytes_read.sum")]); wr = float(vals[hdr.index("dram__bytes_write.sum")])
                u = units[hdr.index("dram__bytes_read.sum")]
                print(f"   traffic (dram read + write)                                                 {rd + wr:16.3f} {u} per launch")
            except (ValueError, IndexError):
                pass


if __name__ == "__main__":
    main()
