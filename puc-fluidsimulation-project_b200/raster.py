"""Output sink: the frames the reference shows with matplotlib, rendered on the GPU and written as PNG.

Replaces ``ax.tripcolor(triang, c, shading="gouraud", cmap="plasma", vmin=0, vmax=1)`` +
``plt.pause`` of ``code/StokesColor.py:508-511,593-598`` and the ``tripcolor`` + ``scatter`` redraw of
``code/StokesFood.py:511-526``.  The field is interpolated per pixel
(barycentric, i.e. Gouraud on the scalar) by ``fs_raster_field``, colour-mapped by
``fs_raster_colormap`` and the tracers are splatted by ``fs_raster_points``; the PNG encoder is the
only host-side step (zlib from the standard library).

matplotlib is not available in this environment, so the colour tables are 5-anchor piecewise-linear
approximations of its "plasma" and "viridis" maps (``colormap_lut``); any (256,3) uint8 table can be
passed instead.
"""
from __future__ import annotations

import os
import struct
import zlib

import numpy as np

from ._lib import call, ptr

_ANCHORS = {
    # value 0, 0.25, 0.5, 0.75, 1
    "viridis": [(68, 1, 84), (59, 82, 139), (33, 145, 140), (94, 201, 98), (253, 231, 37)],
    "plasma": [(13, 8, 135), (126, 3, 168), (204, 71, 120), (248, 149, 64), (240, 249, 33)],
    "gray": [(0, 0, 0), (64, 64, 64), (128, 128, 128), (191, 191, 191), (255, 255, 255)],
}


def colormap_lut(name="viridis"):
    """(256,3) uint8 table, linear between five anchor colours."""
    if name not in _ANCHORS:
        raise ValueError(f"unknown colormap {name!r}; have {sorted(_ANCHORS)}")
    a = np.asarray(_ANCHORS[name], dtype=np.float64)
    t = np.linspace(0.0, 4.0, 256)
    k = np.minimum(t.astype(int), 3)
    f = (t - k)[:, None]
    return np.ascontiguousarray(np.rint(a[k] * (1.0 - f) + a[k + 1] * f).astype(np.uint8))


def raster_field(mesh, field, width, height, extent=(0.0, 1.0, 0.0, 1.0)):
    """(height, width) float32 image of the nodal field; NaN outside the mesh.  Row 0 is the top."""
    field = np.ascontiguousarray(field, dtype=np.float64)
    if field.shape != (mesh.N,):
        raise ValueError(f"field must have shape ({mesh.N},)")
    img = np.empty((height, width), dtype=np.float32)
    x0, x1, y0, y1 = map(float, extent)
    call("fs_raster_field", mesh._h, ptr(field), int(width), int(height), x0, x1, y0, y1, ptr(img))
    return img


def colorize(img, vmin, vmax, cmap="viridis", background=(0, 0, 0, 255)):
    """(H, W, 4) uint8 RGBA picture of a float32 raster."""
    img = np.ascontiguousarray(img, dtype=np.float32)
    lut = colormap_lut(cmap) if isinstance(cmap, str) else np.ascontiguousarray(cmap, dtype=np.uint8)
    if lut.shape != (256, 3):
        raise ValueError("colour table must have shape (256, 3)")
    bg = np.asarray(background, dtype=np.uint8)
    if bg.shape != (4,):
        raise ValueError("background must be an RGBA 4-tuple")
    h, w = img.shape
    rgba = np.empty((h, w, 4), dtype=np.uint8)
    call("fs_raster_colormap", ptr(img), int(w), int(h), float(vmin), float(vmax), ptr(lut), ptr(bg), ptr(rgba))
    return rgba


def splat_points(rgba, points, status=None, colors=((0, 0, 255), (255, 0, 0)), radius_px=2.0,
                 extent=(0.0, 1.0, 0.0, 1.0)):
    """Draw tracers as discs into an RGBA picture (in place); colour = colors[status[i]]
    (``code/StokesFood.py:403,522-523``: blue = uneaten, red = eaten)."""
    if rgba.dtype != np.uint8 or rgba.ndim != 3 or rgba.shape[2] != 4 or not rgba.flags["C_CONTIGUOUS"]:
        raise ValueError("rgba must be a C-contiguous (H, W, 4) uint8 array")
    pts = np.ascontiguousarray(points, dtype=np.float64)
    st = None if status is None else np.ascontiguousarray(status, dtype=np.int32)
    col = np.ascontiguousarray(colors, dtype=np.uint8)
    h, w = rgba.shape[:2]
    x0, x1, y0, y1 = map(float, extent)
    call("fs_raster_points", ptr(rgba), int(w), int(h), x0, x1, y0, y1, ptr(pts), ptr(st), int(len(pts)), ptr(col),
         int(len(col)), float(radius_px))
    return rgba


def draw_quiver(rgba, points, vectors, scale=10.0, half_width_px=0.6, head_frac=0.3, color=(0, 0, 0),
                extent=(0.0, 1.0, 0.0, 1.0)):
    """Velocity arrows into an RGBA picture (in place): ``ax.quiver(x, y, u, v, angles='xy', scale_units='xy', scale=scale)``
    of code/StokesColor.py:514-527 (every 3rd node, scale 10, black) and code/StokesFood.py:517-519 (scale 40, white)."""
    if rgba.dtype != np.uint8 or rgba.ndim != 3 or rgba.shape[2] != 4 or not rgba.flags["C_CONTIGUOUS"]:
        raise ValueError("rgba must be a C-contiguous (H, W, 4) uint8 array")
    pts = np.ascontiguousarray(points, dtype=np.float64)
    vec = np.ascontiguousarray(vectors, dtype=np.float64)
    if pts.shape != vec.shape or pts.ndim != 2 or pts.shape[1] != 2:
        raise ValueError("points and vectors must both have shape (P, 2)")
    col = np.ascontiguousarray(color, dtype=np.uint8)
    h, w = rgba.shape[:2]
    x0, x1, y0, y1 = map(float, extent)
    call("fs_raster_quiver", ptr(rgba), int(w), int(h), x0, x1, y0, y1, ptr(pts), ptr(vec), int(len(pts)), float(scale),
         float(half_width_px), float(head_frac), ptr(col))
    return rgba


def write_png(path, rgba):
    """Minimal PNG writer (8-bit RGBA, no interlace)."""
    rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
    if rgba.ndim != 3 or rgba.shape[2] != 4:
        raise ValueError("rgba must have shape (H, W, 4)")
    h, w = rgba.shape[:2]

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    raw = np.empty((h, 1 + 4 * w), dtype=np.uint8)
    raw[:, 0] = 0                                   # filter type 0 on every scanline
    raw[:, 1:] = rgba.reshape(h, 4 * w)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n")
        f.write(chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0)))
        f.write(chunk(b"IDAT", zlib.compress(raw.tobytes(), 6)))
        f.write(chunk(b"IEND", b""))


def write_apng(path, frames, delay_ms=40, loops=0):
    """Animated PNG (APNG) of equally sized RGBA frames: the movie the reference writes as mp4 through
    matplotlib's ffmpeg writer (`scripts/good_visualization2.py:735-744`); ffmpeg is not available here, APNG needs
    only zlib and plays in browsers.  frames: iterable of (H, W, 4) uint8 arrays."""
    frames = [np.ascontiguousarray(f, dtype=np.uint8) for f in frames]
    if not frames:
        raise ValueError("no frames")
    h, w = frames[0].shape[:2]
    if any(f.shape != (h, w, 4) for f in frames):
        raise ValueError("all frames must have the same (H, W, 4) shape")

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    def deflate(rgba):
        raw = np.empty((h, 1 + 4 * w), dtype=np.uint8)
        raw[:, 0] = 0
        raw[:, 1:] = rgba.reshape(h, 4 * w)
        return zlib.compress(raw.tobytes(), 6)

    seq = 0
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n")
        f.write(chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0)))
        f.write(chunk(b"acTL", struct.pack(">II", len(frames), int(loops))))
        for k, fr in enumerate(frames):
            f.write(chunk(b"fcTL", struct.pack(">IIIIIHHBB", seq, w, h, 0, 0, int(delay_ms), 1000, 0, 0)))
            seq += 1
            if k == 0:
                f.write(chunk(b"IDAT", deflate(fr)))          # the first frame doubles as the still image
            else:
                f.write(chunk(b"fdAT", struct.pack(">I", seq) + deflate(fr)))
                seq += 1
        f.write(chunk(b"IEND", b""))


def read_apng(path):
    """Frames of a file written by write_apng (tests)."""
    data = open(path, "rb").read()
    pos, frames, w, h, n_decl = 8, [], 0, 0, 0
    while pos < len(data):
        n, tag = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        crc, = struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])
        if crc != (zlib.crc32(tag + body) & 0xFFFFFFFF):
            raise ValueError(f"bad CRC in chunk {tag!r}")
        if tag == b"IHDR":
            w, h = struct.unpack(">II", body[:8])
        elif tag == b"acTL":
            n_decl = struct.unpack(">II", body)[0]
        elif tag in (b"IDAT", b"fdAT"):
            raw = np.frombuffer(zlib.decompress(body if tag == b"IDAT" else body[4:]), dtype=np.uint8).reshape(h, 1 + 4 * w)
            frames.append(raw[:, 1:].reshape(h, w, 4).copy())
        pos += 12 + n
    if len(frames) != n_decl:
        raise ValueError("frame count does not match the acTL chunk")
    return frames


def read_png(path):
    """Inverse of write_png for the files it writes (tests)."""
    data = open(path, "rb").read()
    if data[:8] != b"\x89PNG\r\n\x1a\n":
        raise ValueError("not a PNG file")
    pos, idat, w, h = 8, b"", 0, 0
    while pos < len(data):
        n, tag = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        if tag == b"IHDR":
            w, h, depth, ctype = struct.unpack(">IIBB", body[:10])
            if (depth, ctype) != (8, 6):
                raise ValueError("only 8-bit RGBA files are supported")
        elif tag == b"IDAT":
            idat += body
        pos += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, 1 + 4 * w)
    if raw[:, 0].any():
        raise ValueError("only filter type 0 is supported")
    return raw[:, 1:].reshape(h, w, 4).copy()


class FrameSink:
    """Numbered PNG frames in a directory: the replacement for the ``plt.pause`` loop.

        sink = FrameSink("frames", sim.mesh, 512, 512)
        for step in range(STEPS):
            sim.step_all()
            sink.field(sim.c, vmin=0, vmax=1, cmap="plasma")                  # StokesColor
            # sink.field(np.linalg.norm(sim.u, axis=1), 0, vmax, tracers=sim.tracer_points, status=sim.tracer_status)
    """

    def __init__(self, directory, mesh, width=512, height=512, extent=(0.0, 1.0, 0.0, 1.0), prefix="frame"):
        self.directory, self.mesh, self.width, self.height = directory, mesh, int(width), int(height)
        self.extent, self.prefix, self.count = tuple(map(float, extent)), prefix, 0
        os.makedirs(directory, exist_ok=True)

    def render(self, field, vmin, vmax, cmap="viridis", tracers=None, status=None, radius_px=2.0,
               background=(0, 0, 0, 255), colors=((0, 0, 255), (255, 0, 0)), quiver=None):
        """quiver = (points, vectors, scale[, color]): velocity arrows on top of the field, under the tracers."""
        img = raster_field(self.mesh, field, self.width, self.height, self.extent)
        rgba = colorize(img, vmin, vmax, cmap, background)
        if quiver is not None:
            qp, qv, qs = quiver[:3]
            draw_quiver(rgba, qp, qv, qs, color=quiver[3] if len(quiver) > 3 else (0, 0, 0), extent=self.extent)
        if tracers is not None and len(tracers):
            splat_points(rgba, tracers, status, colors, radius_px, self.extent)
        return rgba

    def field(self, field, vmin, vmax, **kw):
        path = os.path.join(self.directory, f"{self.prefix}_{self.count:06d}.png")
        write_png(path, self.render(field, vmin, vmax, **kw))
        self.count += 1
        return path

    def movie(self, path=None, delay_ms=40):
        """Bundle the frames written so far into one animated PNG (the reference's mp4, without ffmpeg)."""
        path = path or os.path.join(self.directory, f"{self.prefix}.apng")
        names = [os.path.join(self.directory, f"{self.prefix}_{k:06d}.png") for k in range(self.count)]
        write_apng(path, (read_png(n) for n in names), delay_ms=delay_ms)
        return path
