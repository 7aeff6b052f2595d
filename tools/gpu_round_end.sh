# What the driver does at round end, plus the profile captures, on one B200 (r02).
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r02.json 2> gpurun_out/bench_ref_r02.err; cut -c1-300 gpurun_out/bench_ref_r02.json
python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; tail -2 gpurun_out/bench_r02.err; cut -c1-300 gpurun_out/bench_r02.json
python tools/run_configs.py C1 C5 > gpurun_out/configs_1gpu_r02.jsonl 2> gpurun_out/configs_1gpu_r02.err; cut -c1-260 gpurun_out/configs_1gpu_r02.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-scale > gpurun_out/ncu_bench_r02.log 2>&1; tail -1 gpurun_out/ncu_bench_r02.log | cut -c1-200
ncu --set full --import-source on --clock-control none -k regex:presync_kernel -s 1 -c 1 -o gpurun_out/prof_presync_r02_final -f python tools/prof_presync.py C2 2 > gpurun_out/ncu_r02_final.log 2>&1; tail -1 gpurun_out/ncu_r02_final.log
