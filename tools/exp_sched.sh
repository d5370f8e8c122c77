M="gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second,lts__t_sectors_srcunit_tex_op_read_evict_normal_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read_evict_normal_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_read_evict_last_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read_evict_last_lookup_miss.sum"
for cfg in 1:10 1:20 1:40 2:10 16:10; do
  c=${cfg%%:*}; g=${cfg##*:}
  python tools/gpu_check.py --section perfx:rank:2:10000:1000000:$c:$g:0:0 2>&1 | tail -1
  ncu --metrics $M --clock-control none -k gemm_kernel -s 3 -c 1 python tools/gpu_check.py --section perfx:rank:2:10000:1000000:$c:$g:0:0 2>&1 | grep -E "dram__bytes_read|gpu__time|hit_rate|tensor|evict|cycles_elapsed"
done
