"""Probe (torchrun, N >= 2): is CUDA symmetric memory / NVSwitch multicast usable on this box, and how do the library
all-reduces compare at the gradient sizes of the cfg2 step (36 MB total, 10 MB last bucket)?  Not part of the product."""
import os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm

def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    grp = dist.group.WORLD
    out = {"NCCL_ALGO": os.environ.get("NCCL_ALGO"), "NCCL_PROTO": os.environ.get("NCCL_PROTO"), "world": world}
    try:
        if os.environ.get("PROBE_NCCL_ONLY") == "1":
            raise RuntimeError("skipped")
        t = symm.empty(9 * 1024 * 1024, dtype=torch.float32, device="cuda")
        h = symm.rendezvous(t, group=grp)
        out["buffer_ptrs"] = [hex(p) for p in h.buffer_ptrs]
        out["signal_pad_ptrs"] = [hex(p) for p in h.signal_pad_ptrs]
        out["multicast_ptr"] = hex(h.multicast_ptr)
        out["signal_pad_size"] = h.signal_pad_size
        out["buffer_size"] = h.buffer_size
    except Exception as e:  # noqa
        out["symm_error"] = repr(e)[:300]
        t, h = None, None
    def timeit(fn, n=20):
        for _ in range(5): fn()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    for mb in (1, 4, 10, 14, 36):
        n = mb * 1024 * 1024 // 4
        x = torch.ones(n, device="cuda")
        out[f"nccl_allreduce_{mb}MB_us"] = round(timeit(lambda: dist.all_reduce(x)), 1)
        if t is not None:
            v = t[:n]
            for name in ("multimem_all_reduce_", "two_shot_all_reduce_", "one_shot_all_reduce"):
                try:
                    op = getattr(torch.ops.symm_mem, name)
                    out[f"{name}{mb}MB_us"] = round(timeit(lambda: op(v, "sum", grp.group_name)), 1)
                except Exception as e:  # noqa
                    out[f"{name}{mb}MB_err"] = repr(e)[:160]
    if rank == 0:
        import json
        print(json.dumps(out, indent=None if os.environ.get('PROBE_NCCL_ONLY') == '1' else 1))
    dist.destroy_process_group()

if __name__ == "__main__":
    main()
