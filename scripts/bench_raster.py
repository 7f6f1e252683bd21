"""Output-sink throughput: device raster of a nodal field on the 4M-triangle bench mesh.
    python scripts/bench_raster.py  >> profiles/r01_raster.jsonl
Timed with CUDA events through the C ABI with device-resident buffers; the PNG encode (host, zlib) is timed separately."""
import ctypes as C, json, os, sys, tempfile, time
sys.path.insert(0, ".")
import numpy as np, torch
import fluidsim_b200 as fb
from fluidsim_b200 import _lib

nt, nr = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2048, 1024)
nodes, markers, tris = fb.square_with_hole(nt, nr)
m = fb.Mesh(nodes, tris, markers)
field = torch.from_numpy(np.hypot(nodes[:, 0] - 0.3, nodes[:, 1] - 0.6)).cuda()
lut = torch.from_numpy(fb.colormap_lut("plasma")).cuda()
bg = np.array([0, 0, 0, 255], dtype=np.uint8)
for W in (512, 1024, 2048):
    img = torch.empty((W, W), dtype=torch.float32, device="cuda")
    rgba = torch.empty((W, W, 4), dtype=torch.uint8, device="cuda")
    def frame():
        _lib.call("fs_raster_field", m._h, C.c_void_p(field.data_ptr()), W, W, 0.0, 1.0, 0.0, 1.0, C.c_void_p(img.data_ptr()))
        _lib.call("fs_raster_colormap", C.c_void_p(img.data_ptr()), W, W, 0.0, 1.0, C.c_void_p(lut.data_ptr()), _lib.ptr(bg),
                  C.c_void_p(rgba.data_ptr()))
    for _ in range(3):
        frame()
    ms = C.c_float(0)
    reps = 20
    _lib.call("fs_timer_start")
    for _ in range(reps):
        frame()
    _lib.call("fs_timer_stop", C.byref(ms))
    host = rgba.cpu().numpy()
    t0 = time.perf_counter()
    with tempfile.TemporaryDirectory() as d:
        fb.write_png(os.path.join(d, "f.png"), host)
    t_png = time.perf_counter() - t0
    print(json.dumps({"what": "fs_raster_field + fs_raster_colormap", "triangles": int(m.T), "pixels": W * W,
                      "ms_per_frame": ms.value / reps, "Mpixel_per_s": W * W / (ms.value / reps) / 1e3,
                      "png_encode_ms_host": 1e3 * t_png}), flush=True)
