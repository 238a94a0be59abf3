mkdir -p gpurun_out
# launch list of the bench command (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2m_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_bench_under_ncu.log 2>&1
# full capture of the two kernels of a step
ncu --set full --clock-control none --import-source on -k regex:'k_inside_any32|k_prep_reg' --launch-skip 6 -c 2 -o gpurun_out/r2m_pipeline python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_ncu.log 2>&1
ls -la gpurun_out/r2m*
# the driver-style numbers, clean (not under a profiler)
python bench.py --steps 10 --warmup 3 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2m_bench_reference.json 2> gpurun_out/r2m_bench_reference.err
tail -c 600 gpurun_out/r2m_bench_reference.json
python tools/latency_check.py > gpurun_out/r2m_latency_check.txt 2>&1; head -6 gpurun_out/r2m_latency_check.txt
