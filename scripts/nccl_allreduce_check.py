import os, torch, torch.distributed as dist, time
rank=int(os.environ["RANK"]); torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
for n, dt in ((2097152, torch.float64), (16777216, torch.float64), (16777216, torch.float32), (16777216, torch.int64)):
    x=torch.ones(n, dtype=dt, device="cuda")
    for _ in range(3): dist.all_reduce(x)
    torch.cuda.synchronize(); a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): dist.all_reduce(x)
    b.record(); torch.cuda.synchronize()
    ms=a.elapsed_time(b)/5
    if rank==0: print(n, dt, "%.3f ms  algbw %.1f GB/s" % (ms, n*x.element_size()/ms/1e6), flush=True)
dist.destroy_process_group()
