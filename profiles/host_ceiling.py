"""Host-link ceiling: pinned cudaMemcpyAsync H2D + D2H on 1, 2, 4, 8 GPUs AT ONCE (one plain copy per buffer, as
csrc/lib.cu issues them), the sizes of one Waymo frame (45 MB up, 60 MB down).  bench.py reports `e2e` as a
fraction of `aggregate_duplex_gbs[n_gpus]`.

    python profiles/host_ceiling.py > profiles/r02_host_ceiling.json        # on the 8-GPU box
"""
import json
import os
import sys
import threading
import time

import torch

UP, DOWN, REPS = 45_081_628, 60_218_400, 12


def run(n, mode):
    """mode: 'duplex' | 'h2d' | 'd2h'.  One thread per GPU (as one process per GPU would), all started together."""
    bufs = []
    for d in range(n):
        torch.cuda.set_device(d)
        bufs.append(dict(hu=torch.empty(UP, dtype=torch.uint8).pin_memory(), hd=torch.empty(DOWN, dtype=torch.uint8).pin_memory(),
                         du=torch.empty(UP, dtype=torch.uint8, device=f"cuda:{d}"),
                         dd=torch.empty(DOWN, dtype=torch.uint8, device=f"cuda:{d}"),
                         s_in=torch.cuda.Stream(d), s_out=torch.cuda.Stream(d)))
    barrier = threading.Barrier(n)
    times = [0.0] * n

    def work(d):
        torch.cuda.set_device(d)
        b = bufs[d]
        for it in range(2):                       # warm-up, then timed
            torch.cuda.synchronize(d)
            barrier.wait()
            t0 = time.perf_counter()
            for _ in range(REPS):
                if mode in ("duplex", "h2d"):
                    with torch.cuda.stream(b["s_in"]):
                        b["du"].copy_(b["hu"], non_blocking=True)
                if mode in ("duplex", "d2h"):
                    with torch.cuda.stream(b["s_out"]):
                        b["hd"].copy_(b["dd"], non_blocking=True)
            torch.cuda.synchronize(d)
            times[d] = time.perf_counter() - t0
            barrier.wait()

    th = [threading.Thread(target=work, args=(d,)) for d in range(n)]
    [t.start() for t in th]
    [t.join() for t in th]
    nbytes = REPS * ((UP if mode != "d2h" else 0) + (DOWN if mode != "h2d" else 0))
    return n * nbytes / max(times) / 1e9


def main():
    ng = torch.cuda.device_count()
    out = {"gpus_visible": ng, "bytes_up": UP, "bytes_down": DOWN, "reps": REPS, "cpu_count": os.cpu_count(),
           "aggregate_duplex_gbs": {}, "aggregate_h2d_gbs": {}, "aggregate_d2h_gbs": {},
           "how": "pinned torch tensors, one copy_(non_blocking) per buffer and direction on two streams per GPU, "
                  "one host thread per GPU started on a barrier, wall clock over 12 frames' worth per GPU"}
    for n in (1, 2, 4, 8):
        if n > ng:
            break
        out["aggregate_duplex_gbs"][str(n)] = run(n, "duplex")
        out["aggregate_h2d_gbs"][str(n)] = run(n, "h2d")
        out["aggregate_d2h_gbs"][str(n)] = run(n, "d2h")
    try:
        out["numa_nodes"] = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
    except OSError:
        out["numa_nodes"] = None
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
