#!/bin/bash
# Clock-independent comparison of the fused-kernel variants: elapsed SM cycles and tensor-pipe activity per launch.
#   gpurun -- 'bash tools/ncu_cycles.sh [rows]'   -> gpurun_out/fuse_cycles.csv
rows=${1:-65536}
mkdir -p gpurun_out
timeout 900 ncu --metrics sm__cycles_elapsed.max,gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:laff_fuse_kernel --csv --log-file gpurun_out/fuse_cycles.csv \
  python tools/bench_fuse.py "$rows" --quick > gpurun_out/fuse_cycles.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/fuse_cycles.csv")) if len(r) > 10]
hdr = rows[0]
iname, imet, ival, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
per = {}
for r in rows[1:]:
    per.setdefault(int(r[iid]), {"k": r[iname]})[r[imet]] = float(r[ival].replace(",", ""))
for i in sorted(per):
    d = per[i]
    print(i, d["k"][:40], "cycles %.0f" % d.get("sm__cycles_elapsed.max", 0), "ms %.3f" % (d.get("gpu__time_duration.sum", 0) / 1e6),
          "tensor %.1f%%" % d.get("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", 0))
PY
