#!/bin/bash
# DRAM traffic and throughput of the HBM-bound kernels (pooling, frame pooling, normalise / cast / split, list extraction),
# one ncu pass over the same launches tools/bench_kernels.py and tools/bench_lists.py time with CUDA events.
#   gpurun -- 'bash tools/ncu_hbm_kernels.sh'   -> gpurun_out/hbm_kernels.csv + a per-kernel summary on stdout
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second
K='regex:attention_pool_kernel|frame_pool_kernel|l2norm_quantize_kernel|cast_pad_kernel|split3_kernel'
timeout 900 ncu --metrics $M --clock-control none -k "$K" -c 150 --csv --log-file gpurun_out/hbm_kernels.csv \
  python tools/bench_kernels.py > gpurun_out/hbm_kernels.log 2>&1
K2='regex:topk_sort_kernel|topk_select_kernel|EpiCollect|collect_init_kernel'
timeout 900 ncu --metrics $M --clock-control none -k "$K2" -c 60 --csv --log-file gpurun_out/hbm_kernels_lists.csv \
  python tools/bench_lists.py 2048 1000000 2000 > gpurun_out/hbm_kernels_lists.log 2>&1
python - <<'PY'
import csv, collections, statistics
for path in ("gpurun_out/hbm_kernels.csv", "gpurun_out/hbm_kernels_lists.csv"):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    per = collections.defaultdict(dict)
    for r in rows[1:]:
        per[int(r[ix["ID"]])]["k"] = r[ix["Kernel Name"]]
        per[int(r[ix["ID"]])]["grid"] = r[ix["Grid Size"]]
        per[int(r[ix["ID"]])][r[ix["Metric Name"]]] = (float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]])
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3}
    groups = collections.defaultdict(list)
    for i, d in per.items():
        t = d["gpu__time_duration.sum"][0] * scale[d["gpu__time_duration.sum"][1]]
        rd = d["dram__bytes_read.sum"][0] * scale[d["dram__bytes_read.sum"][1]]
        wr = d["dram__bytes_write.sum"][0] * scale[d["dram__bytes_write.sum"][1]]
        groups[(d["k"][:70], d["grid"])].append((t, rd, wr, d["dram__throughput.avg.pct_of_peak_sustained_elapsed"][0]))
    for (k, grid), v in groups.items():
        t = statistics.median(x[0] for x in v); rd = statistics.median(x[1] for x in v); wr = statistics.median(x[2] for x in v)
        pct = statistics.median(x[3] for x in v)
        print("%-72s grid %-14s n=%3d  %8.3f ms  read %7.3f GB  write %7.3f GB  %7.0f GB/s  dram %5.1f %% of peak" % (
            k, grid, len(v), t * 1e3, rd / 1e9, wr / 1e9, (rd + wr) / t / 1e9, pct))
PY
