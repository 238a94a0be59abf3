mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python tools/numa_probe.py > gpurun_out/r2i_numa8.json 2>&1
for n in 8 4 2; do $TR --nproc-per-node $n --master-port 2960$n tools/h2d_bw.py 2>/dev/null | grep '^{' > gpurun_out/r2i_h2d_${n}ranks.jsonl; done
python tools/h2d_bw.py 2>/dev/null | grep '^{' > gpurun_out/r2i_h2d_1ranks.jsonl
python -m pytest tests/test_gpu_distributed_nccl.py -q > gpurun_out/r2i_pytest_nccl.log 2>&1; tail -3 gpurun_out/r2i_pytest_nccl.log
for n in 8 4; do $TR --nproc-per-node $n --master-port 2970$n bench.py --gpus $n --steps 10 --warmup 3 2> gpurun_out/r2i_bench_${n}gpu.err | grep '^{' > gpurun_out/r2i_bench_${n}gpu.json; done
for c in 2 3 4; do $TR --nproc-per-node 8 --master-port 2980$c bench.py --gpus 8 --config $c --steps 10 --warmup 3 2> gpurun_out/r2i_bench_config${c}_8gpu.err | grep '^{' > gpurun_out/r2i_bench_config${c}_8gpu.json; done
cat gpurun_out/r2i_h2d_8ranks.jsonl
ls -la gpurun_out/ | grep r2i
