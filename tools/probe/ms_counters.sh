# Shared-memory pipe counters of one multiple-scattering (or other, $1 = kernel regex) launch of the bench workload.
mkdir -p gpurun_out
M=gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum,l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum,smsp__sass_inst_executed_op_shared_ld.sum,smsp__sass_inst_executed_op_shared_st.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
ncu --metrics $M --clock-control none -k regex:"${1:-multiple_scattering_rows}" -s ${2:-4} -c ${3:-1} --csv --log-file gpurun_out/${TAG:-ms}_counters.csv python tools/ncu_target.py > gpurun_out/${TAG:-ms}_counters.log 2>&1
python - <<PY
import csv
rows = list(csv.reader(open("gpurun_out/${TAG:-ms}_counters.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"]
if hdr:
    h = rows[hdr[0]]
    for r in rows[hdr[0] + 1:]:
        d = dict(zip(h, r))
        print(d.get("Kernel Name", "")[:40], d["Metric Name"], d["Metric Value"])
PY
