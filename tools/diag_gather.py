import os, time, torch, torch.distributed as dist
rank=int(os.environ["RANK"]); world=int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev=torch.device("cuda",rank)
dist.init_process_group("nccl", device_id=dev)
rec=torch.zeros((50,66),dtype=torch.int32,device=dev); out=torch.zeros((world*50,66),dtype=torch.int32,device=dev)
big=torch.zeros(256<<20,dtype=torch.uint8,device=dev)
for _ in range(5): dist.all_gather_into_tensor(out,rec)
torch.cuda.synchronize(); dist.barrier()
def timeit(fn,n=50):
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); 
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
t1=timeit(lambda: dist.all_gather_into_tensor(out,rec))
t2=timeit(lambda: (big.fill_(1), dist.all_gather_into_tensor(out,rec)))
t3=timeit(lambda: big.fill_(1))
if rank==0: print(f"all_gather only {t1*1e3:.1f} us; fill+gather {t2*1e3:.1f} us; fill {t3*1e3:.1f} us")
dist.destroy_process_group()
