#!/usr/bin/env python
"""Where does a band-mode step spend its time?  torchrun --nproc-per-node N tools/bands_probe.py"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from tests import harness as H
from canvas_ity_b200 import _native, sharding

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = _native.load()
size = 4096
y0, rows = sharding.band(size, rank, world)
frame = H.lower_script(H.tiger_script(size, size), size, size)[0]
cv = C.c_void_p()
assert lib.cb200_canvas_create_band(size, size, y0, rows, local, C.byref(cv)) == 0
assert lib.cb200_frame_upload(cv, C.byref(frame.frame)) == 0
band = torch.empty((rows, size, 4), dtype=torch.uint8, device="cuda")
gathered = torch.empty((size, size, 4), dtype=torch.uint8, device="cuda")
st = _native.Stats()
acc = {"replay": 0.0, "readback": 0.0, "gather": 0.0, "frame_ms": 0.0}
for i in range(30):
    t0 = time.perf_counter()
    lib.cb200_frame_replay(cv, 1); lib.cb200_sync(cv)
    t1 = time.perf_counter()
    lib.cb200_read_rgba8_into(cv, C.c_void_p(band.data_ptr()), size, rows, 0, y0); lib.cb200_sync(cv)
    t2 = time.perf_counter()
    dist.all_gather_into_tensor(gathered.view(-1), band.view(-1)); torch.cuda.synchronize()
    t3 = time.perf_counter()
    lib.cb200_get_stats(cv, C.byref(st))
    if i >= 10:
        acc["replay"] += t1 - t0; acc["readback"] += t2 - t1; acc["gather"] += t3 - t2; acc["frame_ms"] += st.last_frame_ms
print("rank %d band rows [%d,%d): per step ms: replay+sync %.3f (device frame %.3f: geometry %.3f raster %.3f sort %.3f composite %.3f) "
      "readback %.3f all_gather %.3f" % (rank, y0, y0 + rows, acc["replay"] * 50, acc["frame_ms"] / 20, st.geometry_ms, st.raster_ms,
                                          st.sort_ms, st.composite_ms, acc["readback"] * 50, acc["gather"] * 50), flush=True)
dist.destroy_process_group()
