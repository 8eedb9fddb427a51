import os, time, torch, torch.distributed as dist
rank=int(os.environ["RANK"]); world=int(os.environ["WORLD_SIZE"]); torch.cuda.set_device(rank)
dev=torch.device("cuda",rank); dist.init_process_group("nccl", device_id=dev)
n=100_000_000
def t(f, reps=5):
    for _ in range(2): f()
    torch.cuda.synchronize(); dist.barrier(); t0=time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); dist.barrier(); return 1e3*(time.perf_counter()-t0)/reps
src=torch.empty(n,dtype=torch.float64,device=dev); dst=torch.empty(n,dtype=torch.float64,device=dev)
ms=t(lambda: dist.all_to_all_single(dst,src))
if rank==0: print("all_to_all_single even 800MB: %.2f ms"%ms, flush=True)
half=n//world
splits=[half+1000*(i-(world-1)/2) for i in range(world)]; splits=[int(s) for s in splits]; splits[-1]=n-sum(splits[:-1])
# uneven but symmetric-ish: need matching recv splits -> exchange
sc=torch.tensor(splits,device=dev); rc=torch.empty_like(sc); dist.all_to_all_single(rc,sc); rcl=rc.tolist()
dst2=torch.empty(sum(rcl),dtype=torch.float64,device=dev)
ms=t(lambda: dist.all_to_all_single(dst2,src,output_split_sizes=rcl,input_split_sizes=splits))
if rank==0: print("all_to_all_single uneven 800MB: %.2f ms"%ms, flush=True)
g=torch.empty(n*world//4,dtype=torch.int32,device=dev); l=torch.empty(n//4,dtype=torch.int32,device=dev)
ms=t(lambda: dist.all_gather_into_tensor(g,l))
if rank==0: print("all_gather_into_tensor 100MB/rank: %.2f ms"%ms, flush=True)
try:
    import torch.distributed._symmetric_memory as symm
    buf=symm.empty(n, dtype=torch.float64, device=dev)
    hdl=symm.rendezvous(buf, dist.group.WORLD)
    peer=hdl.get_buffer((rank+1)%world, (n,), torch.float64)
    ms=t(lambda: peer.copy_(src))
    if rank==0: print("symm_mem peer copy_ 800MB: %.2f ms (%.0f GB/s)"%(ms, 0.8/ms*1e3), flush=True)
except Exception as e:
    if rank==0: print("symm_mem failed:", repr(e)[:300], flush=True)
dist.destroy_process_group()
