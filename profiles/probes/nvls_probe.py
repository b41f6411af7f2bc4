"""Probe / self-check (torchrun, N >= 2) of the multicast all-reduce kernel (csrc/nvls.cu): values against the exact
expected sums, then time per call against ncclAllReduce at the bucket sizes of the cfg2 step.  Not part of the product."""
import ctypes, importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
import __graft_entry__ as ge

def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    ge.build()
    L = importlib.import_module("disentanglement-vae_b200._lib")
    lib = L.load()
    grp = dist.group.WORLD
    nmax = 36 * 1024 * 1024 // 4
    buf = symm.empty(nmax, dtype=torch.float32, device=dev)
    hb = symm.rendezvous(buf, group=grp)
    bar = symm.empty(int(lib.dvae_nvls_barrier_words()), dtype=torch.int32, device=dev)
    bar.zero_()
    hbar = symm.rendezvous(bar, group=grp)
    torch.cuda.synchronize(); dist.barrier()
    bar_ptrs = (ctypes.c_void_p * world)(*[int(p) for p in hbar.buffer_ptrs])
    ctr = torch.zeros(1, dtype=torch.int32, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    out = {"world": world, "multicast_ptr": hex(hb.multicast_ptr)}
    epoch = [0]

    peers = (ctypes.c_void_p * world)(*[int(p) for p in hb.buffer_ptrs])
    p2p = [False]

    def ar(n, ctas=0, soft_ns=0):
        epoch[0] += 1
        ctr.fill_(epoch[0])
        if p2p[0]:
            L.check(lib.dvae_p2p_all_reduce(peers, n, bar_ptrs, rank, world, L.ptr(ctr), 1, 0, ctas, soft_ns,
                                            L.ptr(err) if soft_ns else None, st), "dvae_p2p_all_reduce")
        else:
            L.check(lib.dvae_nvls_all_reduce(hb.multicast_ptr, n, bar_ptrs, rank, world, L.ptr(ctr), 1, 0, ctas, soft_ns,
                                             L.ptr(err) if soft_ns else None, st), "dvae_nvls_all_reduce")
    # values: element i of rank r holds (r + 1) * (i % 97 + 1); the sum is world (world + 1) / 2 * (i % 97 + 1), exact in fp32
    for n in (4, 1024 + 4, 1 << 20, 3670016 + 8, -4, -(1 << 20) - 12):
        p2p[0] = n < 0
        n = abs(n)
        base = (torch.arange(n, device=dev) % 97 + 1).float()
        buf[:n] = base * (rank + 1)
        buf[n:n + 16] = -5.0
        torch.cuda.synchronize(); dist.barrier()
        ar(n, soft_ns=5_000_000_000)
        torch.cuda.synchronize()
        ok = bool(torch.equal(buf[:n], base * (world * (world + 1) // 2))) and bool((buf[n:n + 16] == -5.0).all()) and int(err.item()) == 0
        out[f"values_ok_{'p2p' if p2p[0] else 'nvls'}_n{n}"] = ok
        dist.barrier()
    def timeit(fn, n=30):
        for _ in range(5): fn()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return round(e0.elapsed_time(e1) / n * 1e3, 1)
    buf.zero_()
    for mb in (1, 4, 10, 14, 36):
        n = mb * 1024 * 1024 // 4
        x = torch.zeros(n, device=dev)
        out[f"nccl_{mb}MB_us"] = timeit(lambda: dist.all_reduce(x))
        for ctas in (16, 64, 128):
            p2p[0] = False
            out[f"nvls_{mb}MB_ctas{ctas}_us"] = timeit(lambda: ar(n, ctas))
            p2p[0] = True
            out[f"p2p_{mb}MB_ctas{ctas}_us"] = timeit(lambda: ar(n, ctas))
    if rank == 0:
        print(json.dumps(out, indent=1))
    dist.barrier()
    dist.destroy_process_group()

if __name__ == "__main__":
    main()
