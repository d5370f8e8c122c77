M="gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second"
for cfg in 1:5 1:8 1:10 1:13; do
  c=${cfg%%:*}; g=${cfg##*:}
  python tools/gpu_check.py --section perfx:rank:2:10000:1000000:$c:$g:0:0 2>&1 | tail -1
  ncu --metrics $M --clock-control none -k gemm_kernel -s 3 -c 1 python tools/gpu_check.py --section perfx:rank:2:10000:1000000:$c:$g:0:0 2>&1 | grep -E "dram__bytes_read|gpu__time|hit_rate|tensor|cycles_elapsed"
done
python tools/gpu_check.py --section perfx:rank:2:10000:1000000:1:10:0:2 2>&1 | tail -1
python tools/gpu_check.py --section perfx:rank:2:2990:2990:1:10:0:0 2>&1 | tail -1
python tools/gpu_check.py --section perfx:dense:2:2990:2990:1:10:0:0 2>&1 | tail -1
