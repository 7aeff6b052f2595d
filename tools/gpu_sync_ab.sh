# one B200: GPU parity tests, Sync variants, bench
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for L in 8 16; do echo "== C2 lanes $L"; RSSYNC_SYNC_LANES=$L timeout 300 python tools/prof_sync.py C2 2>&1 | tail -1; done
echo "== C2 minb2"; RSSYNC_B200_LIB=$PWD/rs-sync_b200/lib_minb2/librssync_b200.so timeout 300 python tools/prof_sync.py C2 2>&1 | tail -1
echo "== C4 minb1"; timeout 300 python tools/prof_sync.py C4 2>&1 | tail -1
echo "== C4 minb2"; RSSYNC_B200_LIB=$PWD/rs-sync_b200/lib_minb2/librssync_b200.so timeout 300 python tools/prof_sync.py C4 2>&1 | tail -1
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01b.json 2> gpurun_out/bench_r01b.err; tail -3 gpurun_out/bench_r01b.err; cat gpurun_out/bench_r01b.json
