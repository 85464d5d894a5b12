"""NVLink peer-copy microbenchmark (one process, GPUs 0 and 1; run under gpurun --gpus 2): what a copy-engine push reaches
(a) one copy, one direction, (b) the same bytes split over 2 / 4 streams, (c) both directions at once (the slab exchange pushes
both ways simultaneously), (d) an SM copy kernel (torch index copy through peer access) -- the candidates for the transport of
the global transposes.  Prints GB/s per direction."""
import torch

assert torch.cuda.device_count() >= 2
MB = 1 << 20


def bench(fn, nbytes, reps=5):
    fn()
    for d in (0, 1):
        torch.cuda.synchronize(d)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.device(0):
        e0.record()
    for _ in range(reps):
        fn()
    for d in (0, 1):
        torch.cuda.synchronize(d)
    with torch.cuda.device(0):
        e1.record()
        torch.cuda.synchronize(0)
    return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


for size_mb in (8, 32, 128, 512):
    n = size_mb * MB // 4
    a0 = torch.empty(n, dtype=torch.float32, device="cuda:0").normal_()
    b1 = torch.empty(n, dtype=torch.float32, device="cuda:1")
    a1 = torch.empty(n, dtype=torch.float32, device="cuda:1").normal_()
    b0 = torch.empty(n, dtype=torch.float32, device="cuda:0")
    streams0 = [torch.cuda.Stream(device=0) for _ in range(4)]
    streams1 = [torch.cuda.Stream(device=1) for _ in range(4)]

    def one_way(nsplit):
        def f():
            step = n // nsplit
            for i in range(nsplit):
                with torch.cuda.stream(streams0[i]):
                    b1[i * step:(i + 1) * step].copy_(a0[i * step:(i + 1) * step], non_blocking=True)
            for s_ in streams0[:nsplit]:
                s_.synchronize()
        return f

    def both_ways(nsplit):
        def f():
            step = n // nsplit
            for i in range(nsplit):
                with torch.cuda.stream(streams0[i]):
                    b1[i * step:(i + 1) * step].copy_(a0[i * step:(i + 1) * step], non_blocking=True)
                with torch.cuda.stream(streams1[i]):
                    b0[i * step:(i + 1) * step].copy_(a1[i * step:(i + 1) * step], non_blocking=True)
            for s_ in streams0[:nsplit] + streams1[:nsplit]:
                s_.synchronize()
        return f

    r = {f"1way x{k}": bench(one_way(k), n * 4) for k in (1, 2, 4)}
    r.update({f"2way x{k} (per direction)": bench(both_ways(k), n * 4) for k in (1, 2, 4)})
    print(f"p2p {size_mb} MB: " + "  ".join(f"{k}={v:.0f} GB/s" for k, v in r.items()), flush=True)
