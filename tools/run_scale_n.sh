#!/bin/bash
# usage: tools/run_scale_n.sh N [tag]   (inside a gpurun --gpus N call): bench line + sampler tick stamps of the gene-block sharded path
n=$1; tag=${2:-r3}
export MASTER_ADDR=127.0.0.1
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/${tag}_bench$n.json 2> gpurun_out/${tag}_bench$n.err
BGH_TRACE_TICK=1500 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n tools/time_tick2.py --tune 12 --cta 0 > gpurun_out/${tag}_tick$n.log 2>&1
python tools/bench_summary.py gpurun_out/${tag}_bench$n.json
tail -16 gpurun_out/${tag}_tick$n.log
