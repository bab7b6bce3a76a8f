"""Host <-> device copy bandwidth with all ranks copying at once (torchrun): the bound on end-to-end
throughput when every GPU returns its own CSR to pinned host memory."""
import os, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 768 << 20
h_in, h_out = torch.empty(n // 2, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
d_in, d_out = torch.empty(n // 2, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
for name, a, b in (("H2D alone", True, False), ("D2H alone", False, True), ("H2D + D2H together", True, True)):
    run(a, b)
    dt = run(a, b)
    if rank == 0:
        gb = ((n // 2 if a else 0) + (n if b else 0)) / 1e9
        print("%d GPU(s), %-20s: %.1f ms per round -> %.1f GB/s per GPU, %.1f GB/s aggregate" % (world, name, dt * 1e3, gb / dt, gb / dt * world), flush=True)
if world > 1:
    dist.destroy_process_group()
