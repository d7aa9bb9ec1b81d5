"""Summarise .ncu-rep files (read here with `ncu -i ... --page raw --csv`) as markdown tables of the metrics the
profiles/ notes quote.  Usage: python tools/ncu_md.py title rep1 [rep2 ...] > profiles/x.md"""
import csv, io, subprocess, sys
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg.per_second", "dram__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.max",
        # shared-memory data pipe (the limiter of both trunk kernels, profiles/r02_summary.md part 3): wavefronts of LDS / STS /
        # mbarrier polls, of the tensor core's operand fetch, and how busy the pipe was
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.per_cycle_active", "smsp__inst_executed.sum"]
print("# %s\n" % sys.argv[1])
for rep in sys.argv[2:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    name = vals[hdr.index("Kernel Name")]
    print("## %s\n\n`%s`\n\n| metric | value |\n|---|---|" % (rep.split("/")[-1].replace(".ncu-rep", ""), name))
    for k in KEEP:
        if k in hdr:
            i = hdr.index(k)
            print("| %s | %s %s |" % (k, vals[i], units[i]))
    print()
