"""Turn gpurun_out ncu artefacts into the text summaries kept under profiles/.
  python profiles/summarize.py launches <launches.csv> <title>       -> per-kernel device-time shares
  python profiles/summarize.py full <prof.ncu-rep> <title>           -> selected `ncu --set full` metrics (raw page)
"""
import collections
import csv
import io
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__cluster_size", "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]


def launches(path, title):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr = rows[0]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    t, n = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        unit = r[hdr.index("Metric Unit")]
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        t[r[kn][:110]] += ms
        n[r[kn][:110]] += 1
    tot = sum(t.values())
    print(f"# {title}\n# total device time {tot:.3f} ms over {sum(n.values())} launches")
    for k, v in t.most_common(12):
        print(f"{v:12.4f} ms {n[k]:6d} {100 * v / tot:6.2f}%  {k}")


def full(path, title):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# {title}")
    for row in rows[2:]:
        print(f"# kernel: {row[hdr.index('Kernel Name')]}")
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:100s} {units[i]:15s} {row[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
