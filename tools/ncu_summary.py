"""Text summaries for profiles/ from what a gpurun call brought back.

    python tools/ncu_summary.py full  report.ncu-rep  warp_units  "header line"   > profiles/rNN_x_ncu_full_K.txt
    python tools/ncu_summary.py list  launches.csv    "header line"               > profiles/rNN_x_launches_summary.txt

full: selected raw metrics of the first kernel in the report + the dynamic opcode mix per warp unit
      (tools/ncu_opmix.py).  list: per-kernel count / total / average / share of an ncu launch list
      (--metrics gpu__time_duration.sum --csv).
"""
import collections
import csv
import pathlib
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.avg.per_cycle_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.avg",
    "launch__shared_mem_per_block_dynamic",
    # global-store efficiency (logging kernels): requests, sectors, useful bytes per sector, L2 write hits
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "smsp__sass_average_data_bytes_per_sector_mem_global_op_st.ratio", "lts__t_sectors_srcunit_tex_op_write.sum",
    "lts__t_sector_op_write_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_write.avg.pct_of_peak_sustained_elapsed",
]


def full(rep, units, header):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u, r = rows[0], rows[1], rows[2]
    print(f"# {header}")
    print(f"# kernel: {r[h.index('Kernel Name')]}")
    for k in KEYS:
        if k in h:
            i = h.index(k)
            print(f"{k:80s} {u[i]:16s} {r[i]}")
    print()
    print("# dynamic opcode mix (tools/ncu_opmix.py)")
    mix = subprocess.run([sys.executable, str(pathlib.Path(__file__).with_name("ncu_opmix.py")), rep, units],
                         capture_output=True, text=True).stdout
    print(mix.rstrip())


def launches(path, header):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    h = rows[0]
    ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
        tot[r[ik]] += v
        cnt[r[ik]] += 1
    s = sum(tot.values())
    print(f"# {header}")
    for k in sorted(tot, key=lambda k: -tot[k]):
        print(f"{k[:70]:70s} n={cnt[k]:4d} total={tot[k]:10.3f} ms  avg={1e3 * tot[k] / cnt[k]:10.1f} us  share={100 * tot[k] / s:5.1f}%")


if __name__ == "__main__":
    if sys.argv[1] == "full":
        full(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        launches(sys.argv[2], sys.argv[3])
