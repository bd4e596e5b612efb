"""GPU-box check of bench.py's teardown order (one GPU is enough).

A pinned tensor that was copied on the library's stream and is freed AFTER dx_close() destroyed that
stream makes torch's pinned allocator record an event on a dead stream ("CUDA error: context is
destroyed", SIGABRT) -- what every rank of the N>1 bench did after printing its line.  Freed before
the close, the process ends cleanly.  usage: python scripts/gpu_teardown_check.py [bad|good]"""
import gc
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child(order):
    import torch
    import dextractor_b200 as dx
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    ctx = dx.Context(0)
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    state = {}
    with torch.cuda.stream(ext):
        h = torch.arange(1539, dtype=torch.int64).pin_memory()
        d = torch.empty(1539, dtype=torch.int64, device=dev)
        d.copy_(h, non_blocking=True)
        h2 = torch.empty(1539, dtype=torch.int64).pin_memory()
        h2.copy_(d, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        state["xbuf"] = (h, d, h2)
    del h, d, h2
    if order == "good":
        state.clear()
        gc.collect()
        torch.cuda.synchronize()
    ctx.close()
    state.clear()
    print("child done", order, flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1])
    else:
        for order in ("bad", "good"):
            r = subprocess.run([sys.executable, __file__, order], capture_output=True, text=True)
            print(order, "rc =", r.returncode, "| destroyed-context message:",
                  "context is destroyed" in r.stderr, "|", r.stdout.strip())
