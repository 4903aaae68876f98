mkdir -p gpurun_out
for X in symm peer; do
for TOOL in memcheck racecheck synccheck; do
  PAS_EXCHANGE=$X timeout 900 compute-sanitizer --tool $TOOL --target-processes all --print-limit 20 python tools/sanitize_multi.py > gpurun_out/r2_sanitizer_2ranks_${X}_$TOOL.log 2>&1
  echo "== $X $TOOL rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize multi ok|Error|error" gpurun_out/r2_sanitizer_2ranks_${X}_$TOOL.log | sort | uniq -c | head -8
done
done
