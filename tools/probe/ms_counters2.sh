# Breakdown of the L1 data-pipe wavefronts of one launch ($1 = kernel regex) of the bench workload.
mkdir -p gpurun_out
M=gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_cmd_read.sum,l1tex__data_pipe_lsu_wavefronts_cmd_write.sum,l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum,l1tex__data_pipe_lsu_wavefronts_mem_lgds_cmd_read.sum,l1tex__data_pipe_lsu_wavefronts_mem_lgds_cmd_write.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_cmd_read.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_cmd_write.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_misc.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_shared_st.sum,smsp__inst_executed_op_global_ld.sum,smsp__inst_executed_op_global_st.sum,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum
ncu --metrics $M --clock-control none -k regex:"${1:-multiple_scattering_rows}" -s ${2:-4} -c ${3:-1} --csv --log-file gpurun_out/${TAG:-ms2}_counters.csv python tools/ncu_target.py > gpurun_out/${TAG:-ms2}_counters.log 2>&1
python - <<PY
import csv
rows = list(csv.reader(open("gpurun_out/${TAG:-ms2}_counters.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"]
if hdr:
    h = rows[hdr[0]]
    for r in rows[hdr[0] + 1:]:
        d = dict(zip(h, r))
        print(d.get("Kernel Name", "")[:30], d["Metric Name"], d["Metric Value"])
PY
