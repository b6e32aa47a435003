"""all_gather_into_tensor latency on this box for the message sizes of the detection exchange (tools only)."""
import os, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for nbytes in (4096, 65536, 131072, 524288, 4 << 20):
    src = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    dst = torch.zeros(nbytes * world, dtype=torch.uint8, device="cuda")
    for _ in range(10):
        dist.all_gather_into_tensor(dst, src)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        dist.all_gather_into_tensor(dst, src)
    b.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"ctas={os.environ.get('NCCL_MAX_CTAS','default')} bytes/rank={nbytes}: {a.elapsed_time(b) / 50 * 1000:.1f} us per all-gather", flush=True)
dist.destroy_process_group()
