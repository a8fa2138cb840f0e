"""Summarise `ncu -i X.ncu-rep --page raw --csv` exports: one block per captured launch with the metrics the
design notes quote.  python tools/ncu_summary.py raw.csv [raw2.csv ...]"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.sum.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
STALL = "smsp__average_warps_issue_stalled_"
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path, errors="replace")))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for n, r in enumerate(rows[2:]):
        print("launch %d  %s" % (n, r[ix["Kernel Name"]][:110]))
        for w in WANT:
            if w in ix:
                print("  %-62s %s %s" % (w, r[ix[w]], units[ix[w]]))
        st = []
        for h, i in ix.items():
            if h.startswith(STALL) and h.endswith("_per_warp_active.pct"):
                try:
                    st.append((float(r[i]), h[len(STALL):-len("_per_warp_active.pct")]))
                except ValueError:
                    pass
        st.sort(reverse=True)
        if st:
            print("  top stalls (%% of warp-active cycles): " + ", ".join("%s=%.1f" % (nm, v) for v, nm in st[:6]))
        print()
