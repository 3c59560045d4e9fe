# Everything profiles/ holds for the round, from ONE build in ONE gpurun call:
#   gpurun --timeout 2400 -- 'bash benchmarks/final_measurements.sh'   then   python profiles/make_r2.py (see its header)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2f_gputests.log 2>&1; echo "gpu tests exit $?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke exit $?"
python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench exit $?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; echo "ref exit $?"
( python benchmarks/microbench.py env; python benchmarks/microbench.py sweep; python benchmarks/microbench.py frontend; python benchmarks/microbench.py cost_volume; python benchmarks/microbench.py tower; python benchmarks/microbench.py sample ) 2>/dev/null | grep '^{' > gpurun_out/r2f_microbench.txt; echo "micro exit $?"
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r2f_b2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 180 -c 200 --csv --log-file gpurun_out/r2f_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r2f_ncu1.log 2>&1; echo "launch list exit $?"
python benchmarks/microbench.py env > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_project|k_tile_gather" -s 20 -c 4 -o gpurun_out/r2f_env -f python benchmarks/microbench.py env > gpurun_out/r2f_ncu2.log 2>&1; echo "env ncu exit $?"
python benchmarks/microbench.py tower > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_tower -s 10 -c 5 -o gpurun_out/r2f_tower -f python benchmarks/microbench.py tower > gpurun_out/r2f_ncu3.log 2>&1; echo "tower ncu exit $?"
python benchmarks/microbench.py cost_volume > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_cost_volume" -s 4 -c 2 -o gpurun_out/r2f_cv -f python benchmarks/microbench.py cost_volume > gpurun_out/r2f_ncu4.log 2>&1; echo "cv ncu exit $?"
python benchmarks/microbench.py sample > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_bilinear_sample -s 3 -c 1 -o gpurun_out/r2f_sample -f python benchmarks/microbench.py sample > gpurun_out/r2f_ncu5.log 2>&1; echo "sample ncu exit $?"
bash benchmarks/agent_loop_profile.sh; echo "agent loop profile exit $?"   # then python profiles/make_agent_loop.py
python benchmarks/debug/to_channels_last.py > gpurun_out/r2f_to_channels_last.txt 2>&1; echo "transposer exit $?"
