"""Which GPU should rank r of a frame-parallel job use?

The projection path has no collective: every rank streams its own frames from host memory.  On a multi-GPU box the
host links are the shared resource, and they are not symmetric - on the 8 x B200 node of this project four GPUs
hang off one host bridge (23 GB/s each when all eight copy) and four off another (36 GB/s each), while
``nvidia-smi topo`` inside the VM shows nothing (all NV18, one NUMA node).  torchrun's default LOCAL_RANK -> cuda:LOCAL_RANK
puts a 2- or 4-rank job entirely on the first bridge.  So the mapping is measured: all visible GPUs copy from pinned
memory at once for a fraction of a second, GPUs are grouped by the rate they reach, and ranks are dealt over the
groups in turn (fastest group first) - a 2-rank job gets one GPU on each bridge, a 4-rank job two on each.

Only the choice of device is made here; nothing on the data path changes.
"""
from __future__ import annotations

import json
import os
import time


def probe_host_links(devices=None, mbytes=64, repeats=4):
    """Pinned host -> device GB/s reached by every GPU in ``devices`` while ALL of them copy at the same time (one
    process, one stream and one pinned buffer per GPU).  ~0.1 s per GPU."""
    import torch
    if devices is None:
        devices = list(range(torch.cuda.device_count()))
    n = int(mbytes) << 20
    bufs, devs, streams, events = [], [], [], []
    for d in devices:
        with torch.cuda.device(d):
            bufs.append(torch.empty(n, dtype=torch.uint8, pin_memory=True).fill_(1))
            devs.append(torch.empty(n, dtype=torch.uint8, device="cuda:%d" % d))
            streams.append(torch.cuda.Stream(device=d))
            events.append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
    for k, d in enumerate(devices):                      # warm-up copy (page tables, first-touch)
        with torch.cuda.device(d), torch.cuda.stream(streams[k]):
            devs[k].copy_(bufs[k], non_blocking=True)
    for d in devices:
        torch.cuda.synchronize(d)
    for k, d in enumerate(devices):
        with torch.cuda.device(d), torch.cuda.stream(streams[k]):
            events[k][0].record()
            for _ in range(repeats):
                devs[k].copy_(bufs[k], non_blocking=True)
            events[k][1].record()
    rates = []
    for k, d in enumerate(devices):
        torch.cuda.synchronize(d)
        rates.append(repeats * n / (events[k][0].elapsed_time(events[k][1]) * 1e-3) / 1e9)
    return rates


def interleaved_order(rates, tolerance=0.15):
    """Device indices (positions in ``rates``) ordered for dealing ranks: devices are grouped by rate (a new group
    starts where the rate drops by more than ``tolerance`` relative to the group's fastest member) and taken from the
    groups in turn, fastest group first.  Equal rates give the identity."""
    order = sorted(range(len(rates)), key=lambda i: (-rates[i], i))
    groups = []
    for i in order:
        if groups and rates[i] >= rates[groups[-1][0]] * (1.0 - tolerance):
            groups[-1].append(i)
        else:
            groups.append([i])
    for g in groups:
        g.sort()
    if len(groups) == 1:
        return list(range(len(rates)))
    out, k = [], 0
    while len(out) < len(rates):
        for g in groups:
            if k < len(g):
                out.append(g[k])
        k += 1
    return out


def choose_device(local_rank, local_world, timeout_s=120.0):
    """Device index for ``local_rank`` of a ``local_world``-rank job on this node.  Local rank 0 measures the host
    links of all visible GPUs and publishes the order in a file named after the launcher's pid (every rank of one
    torchrun agent shares it); the others wait for it.  Falls back to ``local_rank`` when there is nothing to choose
    (one rank, as many ranks as GPUs with equal links, TSP_NO_TOPOLOGY=1, or no answer in time).  Returns
    (device, info dict)."""
    import torch
    visible = torch.cuda.device_count()
    info = {"policy": "local_rank", "visible_gpus": visible}
    if os.environ.get("TSP_NO_TOPOLOGY") or local_world <= 1 or visible <= 1:
        return local_rank % max(visible, 1), info
    path = os.path.join(os.environ.get("TMPDIR", "/tmp"), "tsp_b200_devices_%d_%s.json"
                        % (os.getppid(), os.environ.get("MASTER_PORT", "0")))
    if local_rank == 0:
        rates = probe_host_links(list(range(visible)))
        order = interleaved_order(rates)
        tmp = path + ".tmp"
        with open(tmp, "w") as f:
            json.dump({"rates": rates, "order": order}, f)
        os.replace(tmp, path)
    else:
        t0 = time.time()
        while not os.path.exists(path):
            if time.time() - t0 > timeout_s:
                return local_rank % visible, info
            time.sleep(0.05)
    with open(path) as f:
        data = json.load(f)
    info = {"policy": "host links probed with all %d GPUs copying; ranks dealt over the link groups" % visible,
            "visible_gpus": visible, "probe_gbs": [round(r, 1) for r in data["rates"]], "order": data["order"]}
    return data["order"][local_rank % visible], info
