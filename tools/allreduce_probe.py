"""Time of the data-parallel gradient all-reduce (flat 101 MiB fp32 buffer of the bench workload)."""
import os, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 26_466_000
x = torch.randn(n, device=dev)
for op, name in ((dist.ReduceOp.AVG, "avg"), (dist.ReduceOp.SUM, "sum")):
    for _ in range(5): dist.all_reduce(x, op=op)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): dist.all_reduce(x, op=op)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    if rank == 0:
        w = dist.get_world_size()
        print(f"world {w} {name}: {ms*1e3:.0f} us, algbw {n*4/ms/1e6:.0f} GB/s, busbw {n*4/ms/1e6*2*(w-1)/w:.0f} GB/s  env ALGO={os.environ.get('NCCL_ALGO')} PROTO={os.environ.get('NCCL_PROTO')} NVLS={os.environ.get('NCCL_NVLS_ENABLE')}", flush=True)
    x.normal_()
dist.destroy_process_group()
