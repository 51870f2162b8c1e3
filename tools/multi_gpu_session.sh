cd $GRAFT_REPO_ROOT
python tools/numa_probe.py > gpurun_out/r2_numa_probe.txt 2>&1
for n in 1 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n tools/copy_probe.py 2>/dev/null | grep n_gpus >> gpurun_out/r2_copy_probe.jsonl
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/tiling_multigpu.py 2>&1 | grep n_gpus >> gpurun_out/r2_tiling_multigpu.jsonl
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 tools/tiling_multigpu.py 2>&1 | grep n_gpus >> gpurun_out/r2_tiling_multigpu.jsonl
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 10 --warmup 3 --no-configs --mstpp-batch 0 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
cat gpurun_out/r2_copy_probe.jsonl gpurun_out/r2_tiling_multigpu.jsonl; tail -c 600 gpurun_out/r2_bench_n8.json
