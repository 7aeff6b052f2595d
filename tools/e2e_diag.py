import os, sys, time, importlib
sys.path.insert(0, os.getcwd())
import numpy as np, torch, torch.distributed as dist
rank=int(os.environ["RANK"]); world=int(os.environ["WORLD_SIZE"]); local=int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev=torch.device("cuda",local)
if rank != 0 and os.environ.get("BLOCKING"):
    import ctypes
    torch.cuda.init()
    print("blocking sync:", ctypes.CDLL("libcudart.so.12").cudaSetDeviceFlags(ctypes.c_uint(4)), flush=True)
dist.init_process_group("nccl", device_id=dev)
pkg=importlib.import_module("rs-sync_b200"); sharded=importlib.import_module("rs-sync_b200.sharded"); synth=importlib.import_module("rs-sync_b200.synth")
w=synth.make_workload("C2")
prob=pkg.SyncProblem(seed=100); prob.set_stream(torch.cuda.current_stream().cuda_stream)
prob.load(w,bulk=True); prob.flush()
counts=np.full(w.n_frames,w.n_rays)
fb,fe=int(w.frame_ids[0]),int(w.frame_ids[-1])+1
delays=np.linspace(-0.2,0.2,201)
for i in range(8):
    dist.barrier(); torch.cuda.synchronize()
    t=[time.perf_counter()]
    if rank==0:
        prob.SetGyroQuaternions(w.quats,w.quats.shape[0],w.gyro_rate,w.gyro_t0)
        prob.set_track_batch(w.frame_ids,counts,w.ts_a,w.ts_b,w.rays_a,w.rays_b)
    t.append(time.perf_counter())
    sharded.replicate_state(prob,rank=rank,world=world,device=dev)
    t.append(time.perf_counter())
    c=prob.presync_grid(fb,fe,delays,stream=2,call_no=i)
    t.append(time.perf_counter())
    print(f"rank {rank} step {i}: ingest {1e3*(t[1]-t[0]):.2f} replicate {1e3*(t[2]-t[1]):.2f} grid {1e3*(t[3]-t[2]):.2f} ms", flush=True)
dist.destroy_process_group()
