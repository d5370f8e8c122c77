#!/bin/bash
# Round-2 ncu evidence (1 GPU, under gpurun):  bash tools/ncu_r02.sh   -> gpurun_out/r02_*.csv / .ncu-rep
# (1) launch list of the bench command (shares), (2) per-launch tensor / DRAM metrics of the sweep launches of one step
# (traffic), (3) one --set full capture of a sweep launch, (4) --set full of the sparse BoW projection kernel.
mkdir -p gpurun_out
PARTS=${PARTS:-"1 2 3 4 5"}
has() { [[ " $PARTS " == *" $1 "* ]]; }
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --parity-queries 0"
has 1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_bench_launches.csv \
  $B --mode-b-steps 1 > gpurun_out/r02_bench_launches.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second,lts__t_bytes.sum
has 2 && timeout 900 ncu --metrics $M --clock-control none --kernel-name-base demangled -k regex:EpiRank -s 96 -c 32 --csv --log-file gpurun_out/r02_sweep_launches.csv \
  $B --mode-b-steps 0 --pipeline 0 > gpurun_out/r02_sweep_launches.log 2>&1
has 3 && timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:EpiRank -s 100 -c 1 -f -o gpurun_out/r02_sweep_full \
  $B --mode-b-steps 0 --pipeline 0 > gpurun_out/r02_sweep_full.log 2>&1
has 4 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:bow_project -s 2 -c 1 -f -o gpurun_out/r02_bow_project_full \
  $B --mode-b-steps 0 --pipeline 0 > gpurun_out/r02_bow_full.log 2>&1
has 5 && timeout 900 ncu --set full --clock-control none -k regex:laff_fuse_kernel -s 2 -c 1 -f -o gpurun_out/r02_fuse_txt_full \
  $B --mode-b-steps 0 --pipeline 0 > gpurun_out/r02_fuse_txt_full.log 2>&1
ls -la gpurun_out/r02_*
