"""Worker of tests/test_gpu_dp.py::test_all_reduce_kernel_gives_exact_sums_on_every_rank (torchrun, one rank per GPU):
the repo's gradient-exchange kernel (csrc/nvls.cu) on a symmetric buffer, both variants, against exact integer-valued sums."""
import ctypes
import importlib
import os
import sys

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    import __graft_entry__ as ge
    ge.build()
    L = importlib.import_module("disentanglement-vae_b200._lib")
    lib = L.load()
    nmax = 14 * 1024 * 1024 // 4 + 64
    buf = symm.empty(nmax, dtype=torch.float32, device=dev)
    hb = symm.rendezvous(buf, group=dist.group.WORLD)
    bar = symm.empty(int(lib.dvae_nvls_barrier_words()), dtype=torch.int32, device=dev)
    bar.zero_()
    hbar = symm.rendezvous(bar, group=dist.group.WORLD)
    torch.cuda.synchronize()
    dist.barrier()
    assert hb.multicast_ptr, "no multicast address"
    bar_ptrs = (ctypes.c_void_p * world)(*[int(p) for p in hbar.buffer_ptrs])
    ctr = torch.zeros(1, dtype=torch.int32, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    epoch, bad = 0, []
    # the peer-to-peer variant is what the engine uses on 2 GPUs (on more it must pass the engine's start-up self-test first)
    for mode in (("nvls", "p2p") if world == 2 else ("nvls",)):
        for n, off, ctas in ((4, 0, 0), (8, 4, 0), (1024 + 4, 0, 3), (1 << 20, 12, 0), (14 * 1024 * 1024 // 4, 0, 64), (100003 * 4, 8, 16)):
            base = (torch.arange(n, device=dev) % 97 + 1).float()
            buf.fill_(-5.0)
            buf[off:off + n] = base * (rank + 1)
            torch.cuda.synchronize()
            dist.barrier()
            for rep in range(2):          # twice on the same data: the second call sums the sums (world x the first result)
                epoch += 1
                ctr.fill_(epoch)
                if mode == "p2p":
                    peers = (ctypes.c_void_p * world)(*[int(p) + 4 * off for p in hb.buffer_ptrs])
                    L.check(lib.dvae_p2p_all_reduce(peers, n, bar_ptrs, rank, world, L.ptr(ctr), 1, 0, ctas, 10_000_000_000, L.ptr(err), st), "p2p")
                else:
                    L.check(lib.dvae_nvls_all_reduce(hb.multicast_ptr + 4 * off, n, bar_ptrs, rank, world, L.ptr(ctr), 1, 0, ctas,
                                                     10_000_000_000, L.ptr(err), st), "nvls")
            torch.cuda.synchronize()
            want = base * (world * (world + 1) // 2) * world
            ok = torch.equal(buf[off:off + n], want) and bool((buf[:off] == -5.0).all()) and bool((buf[off + n:] == -5.0).all()) \
                and int(err.item()) == 0
            if not ok:
                bad.append((mode, n, off, ctas))
            dist.barrier()
    flag = torch.tensor([len(bad)], device=dev)
    dist.all_reduce(flag)
    if rank == 0:
        print("XCHG_OK" if int(flag.item()) == 0 else f"XCHG_BAD {bad}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 0 else 1)


if __name__ == "__main__":
    main()
