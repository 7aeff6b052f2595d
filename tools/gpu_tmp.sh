set -x
for v in lib lib_w9 lib_w10; do echo "== $v"; RSSYNC_B200_LIB=$PWD/rs-sync_b200/$v/librssync_b200.so python bench.py --steps 10 --warmup 3 --no-cpu --no-sync 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['ms_per_step'])"; done
ncu --set full --import-source on --clock-control none -k regex:presync_kernel -s 1 -c 1 -o gpurun_out/prof_presync_v6 -f python tools/prof_presync.py C2 2 > gpurun_out/ncu_v6.log 2>&1; tail -1 gpurun_out/ncu_v6.log
