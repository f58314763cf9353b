"""Times the exchange-inside-Adam kernel (ncf_adam_p2p over CUDA-IPC peer buffers) against NCCL
reduce-scatter -> ncf_adam_range -> all-gather on buffers of the bench workload's size (ml20m NeuMF
f=32 L=3: 26.5 M fp32 elements), and checks that both leave the same parameters.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node=N --master-addr 127.0.0.1 \
        --master-port 29531 tools/p2p_probe.py [elements]

Rank 0 prints one JSON line.  Device-side times (CUDA events), max over ranks.
"""
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ncf_b200 import ops  # noqa: E402


def timed(fn, iters, dev):
    for _ in range(3):
        fn()
    dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 26_481_217
    per = -(-n // (4 * world)) * 4
    n_pad, lo = per * world, rank * per
    gen = torch.Generator(device=dev).manual_seed(rank)

    gbuf, pbuf = ops.PeerBuffer(n_pad, dev), ops.PeerBuffer(n_pad, dev)
    handles = [None] * world
    dist.all_gather_object(handles, (gbuf.handle(), pbuf.handle()))
    gptrs = [gbuf.address if r == rank else gbuf.open_peer(handles[r][0]) for r in range(world)]
    pptrs = [pbuf.address if r == rank else pbuf.open_peer(handles[r][1]) for r in range(world)]
    dist.barrier()

    grads = torch.randn(n_pad, device=dev, generator=gen) * 1e-3      # this rank's local gradients
    p0 = torch.randn(n_pad, device=dev, generator=torch.Generator(device=dev).manual_seed(99)) * 1e-2
    step = torch.zeros(1, dtype=torch.int64, device=dev)
    flag = torch.zeros(1, device=dev)
    hyper = dict(lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8)

    # --- peer-memory path ---------------------------------------------------------------------------------
    m1, v1 = torch.zeros(per, device=dev), torch.zeros(per, device=dev)

    def p2p_step():
        gbuf.tensor.copy_(grads)          # stands in for the training step that produced the gradients
        dist.all_reduce(flag)
        ops.adam_p2p(gptrs, pptrs, m1, v1, lo, rank, step, **hyper)
        dist.all_reduce(flag)

    pbuf.tensor.copy_(p0)
    dist.barrier()
    p2p_step()
    torch.cuda.synchronize(dev)
    p_after_p2p = pbuf.tensor.clone()

    # --- NCCL path ------------------------------------------------------------------------------------------
    g2, p2 = torch.empty(n_pad, device=dev), p0.clone()
    m2, v2 = torch.zeros(per, device=dev), torch.zeros(per, device=dev)

    def nccl_step():
        g2.copy_(grads)
        mine_g, mine_p = g2[lo:lo + per], p2[lo:lo + per]
        dist.reduce_scatter_tensor(mine_g, g2, op=dist.ReduceOp.AVG)
        ops.adam_range(mine_p, m2, v2, mine_g, step, **hyper)
        dist.all_gather_into_tensor(p2, mine_p)

    nccl_step()
    torch.cuda.synchronize(dev)
    scale = p2.abs().max().item()
    diff = (p_after_p2p - p2).abs().max().item() / scale

    def copy_only():
        g2.copy_(grads)

    iters = 20
    t_copy = timed(copy_only, iters, dev)
    t_p2p = timed(p2p_step, iters, dev) - t_copy
    t_nccl = timed(nccl_step, iters, dev) - t_copy
    if rank == 0:
        mib = n_pad * 4 / 2**20
        print(json.dumps({"world": world, "elements": n_pad, "buffer_mib": round(mib, 1),
                          "p2p_ms": round(t_p2p, 4), "nccl_rs_adam_ag_ms": round(t_nccl, 4),
                          "max_rel_diff_after_one_step": diff}))
    gbuf.close_peers()
    pbuf.close_peers()
    torch.cuda.synchronize(dev)
    dist.barrier()          # nobody frees a buffer a peer still has mapped
    gbuf.free()
    pbuf.free()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
