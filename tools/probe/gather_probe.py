import os, sys, time, torch, torch.distributed as dist
rank=int(os.environ["RANK"]); world=int(os.environ["WORLD_SIZE"]); lr=int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
for mb in (1, 2.3, 9.5):
    n=int(mb*1e6)
    rec=torch.zeros(n,dtype=torch.uint8).pin_memory()
    for it in range(3):
        torch.cuda.synchronize(); dist.barrier(); t0=time.perf_counter()
        d=rec.cuda(non_blocking=True); torch.cuda.synchronize(); t1=time.perf_counter()
        parts=[torch.empty_like(d) for _ in range(world)] if rank==0 else None
        dist.gather(d, parts, dst=0); torch.cuda.synchronize(); t2=time.perf_counter()
        if rank==0:
            h=torch.stack(parts).cpu(); 
        t3=time.perf_counter()
        if rank==0 and it==2: print("MB %.1f: h2d %.2f ms, gather %.2f ms, stack+d2h %.2f ms" % (mb,(t1-t0)*1e3,(t2-t1)*1e3,(t3-t2)*1e3), flush=True)
dist.destroy_process_group()
