"""Stage times of half-frame / eighth-frame renders of a workload with the band filter on and off
(one GPU): what a rank of a 2- / 8-GPU run pays per stage."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import forma_b200, workloads
from forma_b200.binding import RGBA, Color, Rect
name = sys.argv[1] if len(sys.argv) > 1 else "paris4k"
api = forma_b200.load()
comp, w, h = workloads.build_scene(api, name)
r = api.Renderer(0)
fb = torch.zeros(w * h * 4, dtype=torch.uint8, device="cuda:0")
clear = Color(1, 1, 1, 0)
flush = torch.empty(384 << 20, dtype=torch.uint8, device="cuda:0")
def run(crop, label):
    for filt in (1, 0):
        api.set_option("band_filter", filt)
        comp.evict()
        acc = None
        for i in range(8):
            flush.zero_(); torch.cuda.synchronize()
            r.render_device(comp, fb.data_ptr(), w, h, RGBA, clear, crop)
            torch.cuda.synchronize()
            st = r.stage_times()
            if i >= 3:
                acc = {k: acc[k] + v for k, v in st.items()} if acc else dict(st)
        c = r.counters()
        print(label, "filter", filt, {k: round(v / 5, 4) for k, v in acc.items()}, "segments", c["segments"], "entries", c["entries"], flush=True)
run(None, "whole")
run(Rect((0, w), (0, h // 2 // 16 * 16)), "top half")
run(Rect((0, w), (h // 2 // 16 * 16, h)), "bottom half")
run(Rect((0, w), (h // 16 // 8 * 3 * 16, h // 16 // 8 * 4 * 16)), "one eighth")
for v in (20, 21, 22, 23):
    api.set_option("sort_scan_log2", v)
    print("sort_scan_log2 =", v)
    run(None, " whole")
    run(Rect((0, w), (0, h // 2 // 16 * 16)), " top half")
    run(Rect((0, w), (h // 16 // 8 * 3 * 16, h // 16 // 8 * 4 * 16)), " one eighth")
