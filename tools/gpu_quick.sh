# quick GPU check: parity suite + per-phase timings of the bench workload for a few tuning values
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | grep -E "AssertionError|passed|failed|Error" | cut -c1-300
for v in "PAS_MS_BLOCKS=2 PAS_SS_BLOCKS=2" "PAS_MS_BLOCKS=3 PAS_SS_BLOCKS=3" "PAS_MS_BLOCKS=4 PAS_SS_BLOCKS=4"; do
  echo "== $v"; env $v python tools/ncu_target.py; env $v python tools/ncu_target.py --rgb
done
