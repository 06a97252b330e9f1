"""JSON digest of one kernel from an ncu report (the tracked evidence under profiles/).
Usage: python tools/ncu_digest.py report.ncu-rep "launch description" "ncu flags" > profiles/rNx_ncu_step_kernel.json"""
import csv, json, subprocess, sys
rep, launch, flags = sys.argv[1], sys.argv[2], sys.argv[3]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(txt.splitlines()))
h, u, v = r[0], r[1], r[2]
keep = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
        "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers", "launch__registers_per_thread",
        "sm__cycles_elapsed.max", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__sass_average_branch_targets_threads_uniform.pct",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "sm__sass_l1tex_t_requests_pipe_lsu_mem_global_op_ldgsts.sum")
m = {k: [x, un] for k, un, x in zip(h, u, v)
     if k in keep or (k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"))}
name = dict(zip(h, v)).get("Kernel Name", "samsim_step_kernel")
print(json.dumps({"kernel": name, "launch": launch, "ncu": flags, "metrics": dict(sorted(m.items()))}, indent=1))
