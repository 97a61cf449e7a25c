"""Probe (N >= 2, torchrun): does torch symmetric memory work on this box - peer mappings, device-side barrier, and both
inside a CUDA graph?  Prints one line per check.  Decides whether the sharded step can exchange rows with plain loads /
stores over NVLink instead of NCCL all-to-alls."""
import os

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import torch.distributed._symmetric_memory as symm
    n = 1 << 20
    t = symm.empty(n, dtype=torch.float32, device=dev)
    hdl = symm.rendezvous(t, dist.group.WORLD)
    peer = (rank + 1) % world
    if rank == 0:
        print("rendezvous ok; multicast:", hdl.has_multicast_support, "ptrs:", [hex(p) for p in hdl.buffer_ptrs], flush=True)
    t.fill_(float(rank + 1))
    hdl.barrier(channel=0)
    remote = hdl.get_buffer(peer, (n,), torch.float32)
    ok = bool((remote == float(peer + 1)).all())
    print(f"rank {rank}: eager peer read ok={ok}", flush=True)
    hdl.barrier(channel=0)
    # graph: write own buffer, barrier, read the peer's, barrier
    val = torch.zeros(1, device=dev)
    out = torch.zeros(n, device=dev)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        t.copy_(val.expand(n) + float(rank))
        hdl.barrier(channel=1)
        out.copy_(remote)
        hdl.barrier(channel=1)
    good = True
    for it in range(5):
        val.fill_(float(10 * it))
        g.replay()
        torch.cuda.synchronize()
        good &= bool((out == float(10 * it + peer)).all())
    print(f"rank {rank}: graph replay peer read ok={good}", flush=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(50):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"rank {rank}: graph (4 MB write + barrier + 4 MB peer read + barrier) {e0.elapsed_time(e1) / 50 * 1e3:.1f} us", flush=True)
    e0.record()
    for _ in range(200):
        hdl.barrier(channel=2)
    e1.record(); torch.cuda.synchronize()
    print(f"rank {rank}: eager barrier {e0.elapsed_time(e1) / 200 * 1e3:.1f} us", flush=True)
    del g
    dist.barrier(); torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()
