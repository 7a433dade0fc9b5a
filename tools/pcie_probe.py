"""Host<->device copy ceiling of the box, with nothing but cudaMemcpyAsync: what bounds the `e2e` figure of bench.py.

    python tools/pcie_probe.py [--gpus 1,2,4,8] [--d2h-mb 1804] [--h2d-mb 56] [--reps 5]

For every N in --gpus, N processes (one per GPU, started together) each repeat `reps` times: one pinned H2D copy of --h2d-mb and
one pinned D2H copy of --d2h-mb (the bytes of one 1M-row DirectEmulator.predict step: 56 MB of parameters in, 1,804 MB of spectra
out), on two streams so the directions overlap, timed on the host between barriers.  Reports per-GPU and aggregate GB/s.  A second
pass copies from PAGEABLE memory (what a plain numpy caller of the reference API has).  Prints one JSON object.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time


def worker(rank, n, args, barrier, q):
    import torch

    torch.cuda.set_device(rank)
    d2h_elems, h2d_elems = args.d2h_mb * 1_000_000 // 4, args.h2d_mb * 1_000_000 // 4
    dev_out = torch.empty(d2h_elems, dtype=torch.float32, device="cuda")
    dev_in = torch.empty(h2d_elems, dtype=torch.float32, device="cuda")
    res = {}
    for kind in ("pinned", "pageable"):
        host_out = torch.empty(d2h_elems, dtype=torch.float32, pin_memory=(kind == "pinned"))
        host_in = torch.ones(h2d_elems, dtype=torch.float32)
        if kind == "pinned":
            host_in = host_in.pin_memory()
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        for _ in range(2):  # warm-up (page faults of the pageable buffers included)
            with torch.cuda.stream(s1):
                dev_in.copy_(host_in, non_blocking=True)
            with torch.cuda.stream(s2):
                host_out.copy_(dev_out, non_blocking=True)
            torch.cuda.synchronize()
        barrier.wait()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            with torch.cuda.stream(s1):
                dev_in.copy_(host_in, non_blocking=True)
            with torch.cuda.stream(s2):
                host_out.copy_(dev_out, non_blocking=True)
            torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.reps
        barrier.wait()
        res[kind] = {"s_per_step": dt, "gbs": (args.d2h_mb + args.h2d_mb) / 1e3 / dt}
        del host_out, host_in
    q.put((rank, res))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", default="1")
    ap.add_argument("--d2h-mb", type=int, default=1804)
    ap.add_argument("--h2d-mb", type=int, default=56)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch

    have = torch.cuda.device_count()
    out = {"d2h_mb": args.d2h_mb, "h2d_mb": args.h2d_mb, "reps": args.reps, "host_cores": len(os.sched_getaffinity(0)), "runs": []}
    ctx = mp.get_context("spawn")
    for n in [int(v) for v in args.gpus.split(",")]:
        if n > have:
            out["runs"].append({"n_gpus": n, "skipped": f"only {have} GPUs visible"})
            continue
        barrier, q = ctx.Barrier(n), ctx.Queue()
        ps = [ctx.Process(target=worker, args=(r, n, args, barrier, q)) for r in range(n)]
        for p in ps:
            p.start()
        got = dict(q.get(timeout=600) for _ in range(n))
        for p in ps:
            p.join()
        run = {"n_gpus": n}
        for kind in ("pinned", "pageable"):
            per = [got[r][kind]["gbs"] for r in range(n)]
            slowest = max(got[r][kind]["s_per_step"] for r in range(n))
            run[kind] = {"per_gpu_gbs": [round(v, 2) for v in per], "aggregate_gbs": round(n * (args.d2h_mb + args.h2d_mb) / 1e3 / slowest, 2),
                         "signals_per_s_ceiling": round(n * 1e6 * (args.d2h_mb / 1804.0) / slowest, 0)}
        out["runs"].append(run)
    print(json.dumps(out))


if __name__ == "__main__":
    sys.exit(main())
