"""(torchrun, >= 2 GPUs) Can the ranks of one node map each other's device memory?  Tries torch's symmetric memory
(`torch.distributed._symmetric_memory`) and the CUDA-IPC path (`UntypedStorage._share_cuda_`)."""
import os, sys, traceback
import torch, torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(1024, dtype=torch.float32, device=dev)
    hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
    t.fill_(rank + 1)
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (1024,), torch.float32)
    print(f"[{rank}] symmetric memory ok: peers {[hex(p) for p in hdl.buffer_ptrs]} peer value {peer[:2].tolist()}", flush=True)
    hdl.barrier()
except Exception:
    print(f"[{rank}] symmetric memory FAILED:\n{traceback.format_exc()[-1500:]}", flush=True)
try:
    t = torch.full((1 << 20,), float(rank + 1), device=dev)
    h = t.untyped_storage()._share_cuda_()
    hs = [None] * world
    dist.all_gather_object(hs, h)
    torch.cuda.synchronize(); dist.barrier()
    views = []
    for r, hh in enumerate(hs):
        if r == rank:
            views.append(t); continue
        st = torch.UntypedStorage._new_shared_cuda(*hh)
        views.append(torch.empty(0, dtype=torch.float32, device=st.device).set_(st))
    print(f"[{rank}] cuda ipc ok: values {[float(v[0]) for v in views]} devices {[str(v.device) for v in views]} ptrs {[hex(v.data_ptr()) for v in views]}", flush=True)
    # write into the peer
    views[(rank + 1) % world][1] = 100.0 + rank
    torch.cuda.synchronize(); dist.barrier()
    print(f"[{rank}] my [1] after the peer's store: {float(t[1])}", flush=True)
except Exception:
    print(f"[{rank}] cuda ipc FAILED:\n{traceback.format_exc()[-1500:]}", flush=True)
dist.barrier()
dist.destroy_process_group()
