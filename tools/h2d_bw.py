"""Pinned host -> device copy bandwidth of this box (the ceiling of bench.py's e2e number), alone and with N GPUs copying at
the same time:   python tools/h2d_bw.py [n_gpus ...]      e.g.  python tools/h2d_bw.py 1 2 4 8
Every GPU gets its own process (as under torchrun); the processes start their timed copies together (file barrier)."""
import os
import subprocess
import sys
import time

import torch


def worker(dev: int, n: int, tag: str):
    torch.cuda.set_device(dev)
    out = []
    for mb in (15, 77, 372):
        h = torch.empty(mb * 1000 * 1000, dtype=torch.uint8).pin_memory()
        d = torch.empty_like(h, device="cuda")
        for _ in range(3):
            d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        open(f"/tmp/h2d_{tag}_{mb}_{dev}", "w").close()
        while sum(os.path.exists(f"/tmp/h2d_{tag}_{mb}_{k}") for k in range(n)) < n:   # all ranks ready
            time.sleep(0.001)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 40
        e0.record()
        for _ in range(iters):
            d.copy_(h, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        out.append(mb * iters / e0.elapsed_time(e1))
    print(f"  gpu {dev}: " + "  ".join(f"{mb} MB {g:.1f} GB/s" for mb, g in zip((15, 77, 372), out)), flush=True)
    return out


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--worker":
        worker(int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
        sys.exit(0)
    counts = [int(a) for a in sys.argv[1:]] or [1]
    for n in counts:
        n = min(n, torch.cuda.device_count())
        tag = f"{os.getpid()}_{n}"
        print(f"{n} GPU(s) copying concurrently (pinned host memory, one process per GPU):", flush=True)
        procs = [subprocess.Popen([sys.executable, __file__, "--worker", str(d), str(n), tag]) for d in range(n)]
        for p in procs:
            p.wait()
        for f in os.listdir("/tmp"):
            if f.startswith(f"h2d_{tag}_"):
                os.remove(os.path.join("/tmp", f))
