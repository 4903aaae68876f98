"""Prints the metrics that matter for this path from an `ncu --page raw --csv` export and writes them
as a compact csv (the file committed under profiles/; tools/ncu_to_json.py reads it too).
Usage: ncu -i rep.ncu-rep --page raw --csv --print-metric-instances details > raw.csv
       python tools/ncu_summary.py raw.csv [out.csv]"""
import csv
import sys

csv.field_size_limit(1 << 30)

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum',
        'sm__cycles_active.avg', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum',
        'l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum',
        'smsp__sass_inst_executed_op_shared_ld.sum', 'smsp__sass_inst_executed_op_shared_st.sum',
        'smsp__sass_inst_executed_op_global_ld.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        # executed-work census: thread instructions per SASS opcode (needs --print-metric-instances details)
        'sass__thread_inst_executed_true_per_opcode',
        'l1tex__data_pipe_lsu_wavefronts.avg', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.avg',
        'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
for w in WANT:  # some raw exports prefix a metric with its section ("SM_A.TriageCompute.<metric>")
    if w not in idx:
        full = [h for h in hdr if h.endswith("." + w)]
        if full:
            idx[w] = idx[full[0]]
keep = [w for w in WANT if w in idx]
for r in rows[2:]:
    print('-----')
    for w in keep:
        print(f"{w:86s} {r[idx[w]][:70]} {units[idx[w]]}")
if len(sys.argv) > 2:
    with open(sys.argv[2], 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(keep)
        w.writerow([units[idx[k]] for k in keep])
        for r in rows[2:]:
            # instance details are kept for the opcode census only
            w.writerow([r[idx[k]] if k.endswith("per_opcode") else r[idx[k]].split(" (")[0] for k in keep])
