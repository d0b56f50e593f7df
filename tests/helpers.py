"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np


def camera_rays(width, height, cam, jitter=0.5):
    """Pixel rays of camera.rs:52-66 with a fixed jitter (0.5 = pixel centres); row-major."""
    fw, fh = (width / height, 1.0) if width > height else (1.0, width / height)
    cols, rows = np.meshgrid(np.arange(width, dtype=np.float64), np.arange(height, dtype=np.float64))
    px = (cols + jitter) * (fw * (1.0 / width)) - fw * 0.5
    py = ((height - (rows + 1)) + jitter) * (fh * (1.0 / height)) - fh * 0.5
    d = np.stack([px, py, np.ones_like(px)], -1).reshape(-1, 3)
    o = np.tile(np.asarray(cam, np.float64), (d.shape[0], 1))
    return o, d


def sphere_rays(n, centre, radius, seed):
    """Origins on a sphere of `radius` around `centre`, aimed at random points inside a sphere of radius/3."""
    rng = np.random.default_rng(seed)
    u = rng.normal(size=(n, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    o = np.asarray(centre) + radius * u
    tgt = np.asarray(centre) + rng.normal(size=(n, 3)) * (radius / 6.0)
    return o, tgt - o


def exclude_degenerate(dirs):
    """Rays whose signed-largest direction component is exactly 0 (triangle.rs:108-122 then divides by it)."""
    d = dirs / np.linalg.norm(dirs, axis=1, keepdims=True)
    return d.max(axis=1) > 0.0


def read_png_rgb8(path):
    """Minimal PNG reader for 8-bit RGB, non-interlaced files (what ImageRgbU8::write_png emits): returns (h, w, 3) uint8."""
    import struct, zlib
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w = 8, b"", None
    while pos < len(data):
        n, typ = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        crc, = struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])
        assert zlib.crc32(typ + body) == crc, "chunk CRC"
        if typ == b"IHDR":
            w, h, depth, ctype, comp, flt, inter = struct.unpack(">IIBBBBB", body)
            assert (depth, ctype, comp, flt, inter) == (8, 2, 0, 0, 0)
        elif typ == b"IDAT":
            idat += body
        pos += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 3 * w + 1)
    assert np.all(raw[:, 0] == 0)  # filter type 0 on every scanline
    return raw[:, 1:].reshape(h, w, 3).copy()
