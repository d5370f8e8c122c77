#!/bin/bash
# Pipeline variants of the bench at N GPUs (run under gpurun --gpus N):  tools/exp_scale.sh N out.jsonl "cfg1" "cfg2" ...
N=${1:-8}; OUT=${2:-gpurun_out/r02_scale_exp.jsonl}; shift 2
run() {
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus $N --steps 20 --warmup 5 --mode-b-steps 0 --parity-queries 0 --ab-steps 0 $@ 2>>${OUT%.jsonl}.err | tail -1 | tee -a $OUT | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*', '| value %.0f ms/step %.3f sweep %.3f e2e %.0f' % (d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['e2e']['value']))"
}
for cfg in "$@"; do run $cfg; done
