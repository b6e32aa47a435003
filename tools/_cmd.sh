python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench rc=$?" >> gpurun_out/multi.log
python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/bench_n1.json 2>/dev/null; echo "bench1 rc=$?" >> gpurun_out/multi.log
cat gpurun_out/multi.log; tail -c 600 gpurun_out/bench_n2.json | head -c 300
