#!/bin/bash
# DRAM traffic of the dominant kernels on the state bench.py times (the QCGD loop at 1e7 parents, fourth pass):
#   gpurun -- bash scripts/ncu_traffic.sh          -> gpurun_out/r2_traffic.csv
#   python scripts/ncu_traffic.py gpurun_out/r2_traffic.csv      (here) -> profiles/traffic.json, read by bench.py (roofline.traffic)
# Two metrics only, so that ncu replays each kernel once or twice (the tables are tens of GB: a --set full capture would save and
# restore them ~40 times); launches: per pass symbolic_kernel<split_merge> (children to bins), bin_dedup_kernel,
# symbolic_items_kernel<erase_create>, table_compact_kernel.  --launch-skip 12 = the first three passes.
set -e
PARENTS=${PARENTS:-10000000}
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:"symbolic_items|symbolic_kernel|bin_dedup_kernel|table_compact_kernel" --launch-skip 12 --launch-count 4 \
    --csv --log-file gpurun_out/r2_traffic.csv python scripts/loop_probe.py --parents $PARENTS --passes 4 --skip 4
tail -8 gpurun_out/r2_traffic.csv
