"""Times the pieces of one bench step at N ranks (diagnostic, not a benchmark)."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scanner_b200 as S
rank=int(os.environ.get("RANK","0")); world=int(os.environ.get("WORLD_SIZE","1"))
torch.cuda.set_device(rank); dev=torch.device("cuda",rank)
if world>1: dist.init_process_group("nccl", device_id=dev)
n=2048; n_steps=50; per=4096*world
plan=S.plan_shard(n_steps, per, rank, world); ns=plan.n_units
raw=torch.randint(-20,20,(ns,n,2),dtype=torch.int8,device=dev)
w=S.window_build(5,n)
ctx=S.SpectrumSense(n,20_000_000,8,20.0,w,sample_kind=1,correct_dc_offset=True,max_spectra=1024,max_hits_per_spectrum=16,device=rank)
words=ctx.words
d_spec=torch.empty((ns,n),dtype=torch.float32,device=dev); d_mask=torch.empty((ns,words),dtype=torch.int32,device=dev)
d_cnt=torch.empty((ns,),dtype=torch.int32,device=dev); d_rec=torch.empty((n_steps,words+2),dtype=torch.int32,device=dev)
d_g=torch.empty((world,n_steps,words+2),dtype=torch.int32,device=dev); d_m=torch.empty_like(d_rec)
st=torch.cuda.current_stream(); sh=st.cuda_stream
def ev():
    e=torch.cuda.Event(enable_timing=True); e.record(st); return e
for it in range(6):
    torch.cuda.synchronize()
    if world>1: dist.barrier()
    c0=time.perf_counter(); e0=ev()
    ctx.launch_device(raw.data_ptr(),ns,d_spec.data_ptr(),d_mask.data_ptr(),d_cnt.data_ptr(),0,0,sh); c1=time.perf_counter(); e1=ev()
    ctx.summarize_steps(d_mask.data_ptr(),d_cnt.data_ptr(),ns,plan.first_unit,per,n_steps,d_rec.data_ptr(),sh); c2=time.perf_counter(); e2=ev()
    if world>1: S.gather_step_records(d_rec,world,out=d_g)
    c3=time.perf_counter(); e3=ev()
    if world>1: ctx.merge_step_records(d_g.data_ptr(),world,n_steps,d_m.data_ptr(),sh)
    c4=time.perf_counter(); e4=ev()
    torch.cuda.synchronize()
    if it>=3:
        print(f"rank{rank} it{it} GPU ms: fused {e0.elapsed_time(e1):.3f} summarize {e1.elapsed_time(e2):.3f} gather {e2.elapsed_time(e3):.3f} merge {e3.elapsed_time(e4):.3f} | CPU ms: {1e3*(c1-c0):.3f} {1e3*(c2-c1):.3f} {1e3*(c3-c2):.3f} {1e3*(c4-c3):.3f}", flush=True)
if world>1: dist.destroy_process_group()
