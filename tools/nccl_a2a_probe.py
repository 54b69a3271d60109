"""NCCL send/recv all-to-all timing at the sizes the sharded build exchanges (diagnosis; launched with torch.distributed.run)."""
import os
import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for mb_total in (50, 200, 400, 1200):
    n = mb_total * (1 << 20) // world // 8 * 8
    send = torch.empty(n * world, dtype=torch.uint8, device="cuda")
    recv = torch.empty_like(send)
    for it in range(6):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.all_to_all_single(recv, send)
        e1.record(); torch.cuda.synchronize()
        if it == 5 and rank == 0:
            ms = e0.elapsed_time(e1)
            print(f"all_to_all {mb_total} MB per rank ({world} ranks): {ms:.3f} ms, {mb_total * (world - 1) / world / ms:.1f} GB/s sent per rank", flush=True)
dist.destroy_process_group()
