"""GPU box, under torchrun: what the host side gives each rank's 128 MiB pinned upload - one rank at a time, then all ranks at once.
Names the limiter of the N-GPU e2e number (VERDICT r01 task 8).
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 tools/gpu/h2d_probe.py"""
import os, time, torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nbytes = 128 << 20
h = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); h.fill_(rank)
d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
def rate(reps=20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
def barrier():
    if world > 1: dist.barrier()
rate(3)
alone = 0.0
for r in range(world):
    barrier()
    if r == rank: alone = rate()
barrier()
together = rate()
barrier()
try:
    numa = open(f"/sys/bus/pci/devices/{torch.cuda.get_device_properties(local).pci_bus_id:02x}".replace("devices/", "devices/0000:") + ":00.0/numa_node").read().strip()
except Exception:
    numa = "?"
res = torch.tensor([alone, together], device="cuda")
if world > 1:
    allr = [torch.zeros_like(res) for _ in range(world)]
    dist.all_gather(allr, res)
else:
    allr = [res]
if rank == 0:
    print(f"128 MiB pinned host -> device, {world} rank(s); cpu_count {os.cpu_count()}")
    for r, t in enumerate(allr):
        print(f"  rank {r}: alone {t[0].item():6.1f} GB/s   all {world} at once {t[1].item():6.1f} GB/s")
    tot = sum(t[1].item() for t in allr)
    print(f"  aggregate with all ranks copying: {tot:.1f} GB/s; a 128 MiB upload then takes {nbytes / (tot / world * 1e9) * 1e3:.2f} ms per rank against {nbytes / (allr[0][0].item() * 1e9) * 1e3:.2f} ms alone")
if world > 1: dist.destroy_process_group()
