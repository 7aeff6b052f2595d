# What the driver does at round end, plus the profile captures, on one B200.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r01.json 2> gpurun_out/bench_ref_r01.err; tail -1 gpurun_out/bench_ref_r01.json | cut -c1-400
python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; tail -2 gpurun_out/bench_r01.err; cut -c1-300 gpurun_out/bench_r01.json
python tools/run_configs.py > gpurun_out/configs_1gpu.jsonl 2> gpurun_out/configs_1gpu.err; cut -c1-260 gpurun_out/configs_1gpu.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_v6.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-200
ncu --set full --import-source on --clock-control none -k regex:presync_kernel -s 1 -c 1 -o gpurun_out/prof_presync_v6 -f python tools/prof_presync.py C2 2 > gpurun_out/ncu_v6.log 2>&1; tail -1 gpurun_out/ncu_v6.log
ncu --set full --import-source on --clock-control none -k regex:sync_motion_fgrad -s 20 -c 1 -o gpurun_out/prof_sync_v6 -f python tools/prof_sync.py C2 > gpurun_out/ncu_sync_v6.log 2>&1; tail -1 gpurun_out/ncu_sync_v6.log
