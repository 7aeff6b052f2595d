# N GPUs of one box: the bench and the sharded configurations under torchrun
N=${1:-2}
set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu_r01.json 2> gpurun_out/bench_${N}gpu_r01.err; tail -2 gpurun_out/bench_${N}gpu_r01.err; cut -c1-400 gpurun_out/bench_${N}gpu_r01.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/run_configs.py C3 C4 C5 > gpurun_out/configs_${N}gpu.jsonl 2> gpurun_out/configs_${N}gpu.err; tail -2 gpurun_out/configs_${N}gpu.err; cut -c1-300 gpurun_out/configs_${N}gpu.jsonl
