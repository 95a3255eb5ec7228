"""PCIe probe: the end-to-end ceiling of bench.py's `e2e` leg, measured without any of this repo's code.

One step of the headline workload moves 3 * n_px + 9 * wpf bytes host -> device and the same device -> host (an 8K frame's pixels and
profile words, each crossing once per direction).  This script copies exactly those byte counts between pinned host memory and the
device as plain `copy_` calls, H2D and D2H concurrently on two streams, on every rank at once (barrier first, max over ranks), and
prints one JSON line: the fastest any implementation behind a host-buffer API can run a step on this box.

    python tools/pcie_probe.py                                        # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/pcie_probe.py
"""
import json
import os
import time

import torch
import torch.distributed as dist

N_PX = 7680 * 4320
WPF = 20766726                      # profile words of an 8K frame, RS(26,20) 1D (t3c_profile_words; bench.py's config.profile_words_per_frame)
H2D = D2H = 3 * N_PX + 9 * WPF      # bytes per step and direction


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    p_src = torch.empty(H2D, dtype=torch.uint8).pin_memory()
    p_dst = torch.empty(D2H, dtype=torch.uint8).pin_memory()
    p_src.random_(0, 256)
    d_a = torch.empty(H2D, dtype=torch.uint8, device=dev)
    d_b = torch.zeros(D2H, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt

    def h2d():
        with torch.cuda.stream(s1):
            d_a.copy_(p_src, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            p_dst.copy_(d_b, non_blocking=True)

    def both():
        h2d()
        d2h()

    t_h, t_d, t_b = timed(h2d, 6), timed(d2h, 6), timed(both, 10)
    if rank == 0:
        print(json.dumps({
            "probe": "plain pinned copies of one bench.py e2e step's bytes, all ranks at once, max over ranks",
            "n_gpus": world, "h2d_bytes_per_step": H2D, "d2h_bytes_per_step": D2H,
            "h2d_alone_ms": 1e3 * t_h, "d2h_alone_ms": 1e3 * t_d, "both_ms": 1e3 * t_b,
            "h2d_alone_gbs_per_gpu": H2D / t_h / 1e9, "d2h_alone_gbs_per_gpu": D2H / t_d / 1e9, "both_gbs_per_gpu_total": (H2D + D2H) / t_b / 1e9,
            "e2e_ceiling_mpix_per_s": world * N_PX / t_b / 1e6,
            "gpu": torch.cuda.get_device_name(local)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
