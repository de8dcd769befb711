#!/usr/bin/env python
"""A/B of the host-frame pipeline (options host_slices / slice_bands / copy_bands) and of the
sync-free painter tables, in one process per workload: the scene is built once, every setting
gets a fresh renderer, warm-up frames, then timed frames (wall clock around a synchronised
call; the L2 is flushed between frames like in bench.py).

    python tools/ab_slices.py [workload ...]          # default: paris4k cubics100k circles8k

One JSON line per setting on stdout."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import forma_b200  # noqa: E402
from forma_b200.binding import RGBA, Color  # noqa: E402
from workloads import build_scene  # noqa: E402

CLEAR = Color(1.0, 1.0, 1.0, 0.0)


def main():
    quick = "--quick" in sys.argv
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or ["paris4k", "cubics100k", "circles8k"]
    api = forma_b200.load()
    dev = torch.device("cuda", 0)
    flush = torch.empty(384 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    for name in names:
        comp, w, h = build_scene(api, name)
        host = torch.empty(w * h * 4, dtype=torch.uint8).pin_memory()
        host_np = host.numpy()
        fb = torch.zeros(w * h * 4, dtype=torch.uint8, device=dev)
        ref = None
        heavy = name.startswith("circles8k")
        steps, warm = (6, 3) if heavy else (15, 4)

        def run(tag, opts, e2e):
            nonlocal ref
            saved = {k: api.get_option(k) for k in opts}
            for k, v in opts.items():
                api.set_option(k, v)
            r = api.Renderer(0)
            r.set_stream(stream.cuda_stream)

            def frame():
                if e2e:
                    comp.evict()
                    r.render(comp, host_np, w, h, RGBA, CLEAR)
                else:
                    r.render_device(comp, fb.data_ptr(), w, h, RGBA, CLEAR)
            for _ in range(warm):
                frame()
            torch.cuda.synchronize()
            times, stages = [], {}
            for _ in range(steps):
                flush.zero_()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                frame()
                torch.cuda.synchronize()
                times.append((time.perf_counter() - t0) * 1e3)
                for k, v in r.stage_times().items():
                    stages[k] = stages.get(k, 0.0) + v / steps
            ok = None
            if e2e:
                if ref is None:
                    ref = host_np.copy()
                ok = bool(np.array_equal(host_np, ref))
            c = r.counters()
            print(json.dumps({"workload": name, "tag": tag, "opts": opts, "e2e": e2e, "ms_mean": round(sum(times) / len(times), 4),
                              "ms_min": round(min(times), 4), "fps_mean": round(1e3 * len(times) / sum(times), 1),
                              "slices": [round(v, 3) for v in r.host_slices()], "tables_mode": c["tables_mode"],
                              "max_connections": os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"),
                              "slice_stages": [[round(d[k], 3) for k in r.STAGES] for d in r.host_slice_stages()],
                              "stage_ms": {k: round(v, 4) for k, v in stages.items()}, "same_frame_as_first": ok}), flush=True)
            for k, v in saved.items():
                api.set_option(k, v)
            del r

        # end to end: the plain path first (its frame is the reference of the others)
        run("plain", {"host_slices": 1}, True)
        for cb in (() if quick else (4, 16)):
            run("plain", {"host_slices": 1, "copy_bands": cb}, True)
        for hs in ((2, 3, 4, 6) if quick else (2, 3, 4, 5, 6, 8)):
            for sb in ((1, 2) if quick else (1, 2, 4)):
                if heavy and sb == 4:
                    continue
                run("sliced", {"host_slices": hs, "slice_bands": sb}, True)
        if quick:
            continue
        run("sliced_unchained", {"host_slices": 4, "slice_bands": 2, "slice_chain": 0}, True)
        run("sliced_sync", {"host_slices": 4, "slice_bands": 2, "sync_free": 0}, True)
        # frame left in HBM: painter tables with / without count read-backs
        run("device", {"sync_free": 1}, False)
        run("device", {"sync_free": 0}, False)
        run("device", {"sync_free": 1, "sort_scan_log2": 31}, False)
        del comp


if __name__ == "__main__":
    main()
