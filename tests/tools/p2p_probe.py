"""2-rank microbenchmark of frame transport options (run under torchrun --nproc-per-node 2): symmetric-memory peer
view copy_ (copy engine), an elementwise kernel writing the peer view (SM stores over NVLink), NCCL broadcast."""
import os, sys, json
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 18817920
sym = symm_mem.empty(n, dtype=torch.uint8, device=dev)
hdl = symm_mem.rendezvous(sym, dist.group.WORLD)
src = torch.randint(0, 255, (n,), dtype=torch.uint8, device=dev)
peer = hdl.get_buffer(1 - rank, (n,), torch.uint8, 0)
peer32 = hdl.get_buffer(1 - rank, (n // 4,), torch.int32, 0)
src32 = src.view(torch.int32)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
res = {}
if rank == 0:
    res["ce_copy_peer_view_us"] = timeit(lambda: peer.copy_(src, non_blocking=True))
    res["sm_kernel_store_peer_view_us"] = timeit(lambda: torch.add(src32, 1, out=peer32))
    res["ce_copy_local_us"] = timeit(lambda: sym.copy_(src, non_blocking=True))
else:
    for _ in range(3): dist.barrier()
dist.barrier()
res["nccl_broadcast_us"] = timeit(lambda: dist.broadcast(src, src=0))
if rank == 0:
    res["GBs"] = {k: round(n / v / 1e3, 1) for k, v in res.items()}
    print(json.dumps(res))
dist.destroy_process_group()
