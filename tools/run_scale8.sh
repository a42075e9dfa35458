export MASTER_ADDR=127.0.0.1
for n in 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/r2e_bench$n.json 2> gpurun_out/r2e_bench$n.err
  BGH_TRACE_TICK=1500 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n tools/time_tick2.py --tune 12 --cta 0 > gpurun_out/r2e_tick$n.log 2>&1
done
python tools/bench_summary.py gpurun_out/r2e_bench4.json gpurun_out/r2e_bench8.json
tail -14 gpurun_out/r2e_tick4.log; tail -14 gpurun_out/r2e_tick8.log
M=60000 G=300 C=8 TUNE=12 DRAWS=6 DRIVER=persistent timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29568 tools/sharded_sampler_check.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -4
