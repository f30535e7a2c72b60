"""Development aid: device time per frame of the fast path at the frame sizes of BASELINE configs 2-5, one frame at a
time and with three frames in flight.  python tools/config_probe.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tissue_image_processing_b200 import _native as nat
from tests.test_gpu_movie import _synth_device


def run(name, C, Z, Y, X, shift=0, streams=3, steps=12):
    dev = torch.device("cuda", 0)
    frames = [_synth_device(C, Z, Y, X, 50 + i) for i in range(2)]
    projs = [nat.DeviceProjector(C, Z, Y, X, atoh_shift=shift, mode="fast") for _ in range(streams)]
    strs = [torch.cuda.Stream(device=dev) for _ in range(streams)]

    def serial(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n):
            projs[0].run(frames[i % 2])
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    def piped(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for st in strs:
            st.wait_event(a)
        for i in range(n):
            with torch.cuda.stream(strs[i % streams]):
                projs[i % streams].run(frames[i % 2])
        for st in strs:
            ev = torch.cuda.Event()
            ev.record(st)
            torch.cuda.current_stream().wait_event(ev)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    serial(3); piped(3)
    s, p = serial(steps), piped(steps)
    vox = C * Z * Y * X
    algo = 2 * Z * Y * X * (C + 2) + Y * X * (4 + 4 * C)
    print("%-28s C=%d %dx%dx%d  serial %.3f ms (%.0f Gvox/s, %.2f TB/s algorithmic)  3 in flight %.3f ms (%.0f Gvox/s, %.2f TB/s)"
          % (name, C, X, Y, Z, s, vox / s / 1e6, algo / s / 1e9, p, vox / p / 1e6, algo / p / 1e9), flush=True)
    del frames, projs
    torch.cuda.empty_cache()


if __name__ == "__main__":
    nat.handle(0)
    run("config 2 single stack", 1, 64, 2048, 2048)
    run("config 3 movie frame", 1, 48, 1024, 1024)
    run("config 4 two channels", 2, 64, 2048, 2048)
    run("config 4 two channels, shift", 2, 64, 2048, 2048, shift=2)
    run("config 5 tile 2048", 1, 128, 2048, 2048)
    run("config 5 whole frame", 1, 128, 4096, 4096, steps=6)
