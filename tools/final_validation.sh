# round-end validation on the GPU box; results -> gpurun_out/
set -x
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
WTB200_DISABLE_FUSED=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; tail -c 600 gpurun_out/bench_final_n1.json
timeout 600 python bench.py --impl reference > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; tail -c 400 gpurun_out/bench_final_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-incumbent > /dev/null 2>&1
timeout 300 python tools/perf_other_configs.py 1d 3d mat mat2 db8 2>&1 | grep -v Warn > gpurun_out/other_configs_final.txt; cat gpurun_out/other_configs_final.txt
