"""Worker of tests/test_gpu_nccl.py::test_peer_memory_allreduce_kernel (launched by torch.distributed.run).

pcrl_p2p_allreduce on ranges of a symmetric buffer -- aligned and unaligned offsets / lengths, several calls per channel,
two channels in flight on two streams, graph replay -- against NCCL's all_reduce of the same data."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


class FakeEngine:
    """The three things dist.PeerAllReduce needs from an engine."""

    def __init__(self, n, dev):
        from pointcloud_rl_b200._lib import lib

        self.L, self.device, self._graphs = lib(), torch.device(dev), {}
        self.grads = torch.zeros(n, device=dev)

    def rebind_grads(self, t):
        self.grads = t


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    from pointcloud_rl_b200.dist import PeerAllReduce

    n = 3_000_003
    eng = FakeEngine(n, dev)
    ar = PeerAllReduce(eng, dist.group.WORLD)
    buf = eng.grads
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    ok = True
    cases = [(0, n), (0, 4), (1, 1), (3, 70_001), (4, 75_000), (5, 2_621_443), (1024, 1 << 20), (n - 5, 5), (7, 2)]
    for rep in range(3):
        for off, cnt in cases:
            buf.copy_(torch.randn(n, device=dev, generator=gen))
            before = buf.clone()
            want = before[off:off + cnt].clone()
            dist.all_reduce(want)
            ar(buf[off:off + cnt])
            torch.cuda.synchronize()
            got = buf[off:off + cnt]
            same_as_peers = got.clone()
            dist.broadcast(same_as_peers, src=0)
            good = torch.equal(got, same_as_peers) and torch.allclose(got, want, rtol=1e-6, atol=1e-6)
            if world == 2:
                good = good and torch.equal(got, want)  # a two-term sum has one rounding whatever the order
            untouched = torch.equal(buf[:off], before[:off]) and torch.equal(buf[off + cnt:], before[off + cnt:])
            if not (good and untouched):
                print(f"[rank {rank}] rep {rep} off {off} n {cnt}: match={good} outside untouched={untouched}", flush=True)
            ok = ok and good and untouched
    # two channels in flight on two streams + CUDA-graph replay
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    a_lo, a_n, b_lo, b_n = 0, 1_000_000, 1_500_000, 77_777

    def both():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur)
        s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            h1 = ar(buf[a_lo:a_lo + a_n], async_op=True)
        with torch.cuda.stream(s2):
            h2 = ar(buf[b_lo:b_lo + b_n], async_op=True)
        h1.wait()
        h2.wait()

    both()  # warm-up (assigns the channels)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    cap = torch.cuda.Stream()
    with torch.cuda.graph(g, stream=cap):
        both()
    for rep in range(4):
        buf.copy_(torch.randn(n, device=dev, generator=gen))
        wa, wb = buf[a_lo:a_lo + a_n].clone(), buf[b_lo:b_lo + b_n].clone()
        dist.all_reduce(wa)
        dist.all_reduce(wb)
        torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
        good = torch.allclose(buf[a_lo:a_lo + a_n], wa, rtol=1e-6, atol=1e-6) and torch.allclose(buf[b_lo:b_lo + b_n], wb, rtol=1e-6, atol=1e-6)
        if not good:
            print(f"[rank {rank}] graph replay {rep}: mismatch", flush=True)
        ok = ok and good
    ar.check()
    flag = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("P2P_ALLREDUCE", "PASS" if flag.item() == 1 else "FAIL", f"world {world}", flush=True)
    del g
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
