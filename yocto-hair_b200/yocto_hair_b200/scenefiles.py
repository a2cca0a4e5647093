"""On-disk scene formats of the reference, WRITER side only:
the harness writes the synthetic scenes in the reference's own formats so the
drop-in app (host/yscenetrace_b200) loads them through the reference's
unmodified `yocto_sceneio`:

  <dir>/<name>.json             scene JSON (libs/yocto/yocto_sceneio.cpp:1064-1418)
  <dir>/shapes/<shape>.ply      binary PLY, vertex {x y z nx ny nz radius} float,
                                line {list uchar int vertex_indices}: one polyline
                                record per Bezier (<= 255 vertices, the list size
                                is uint8: libs/yocto/yocto_ply.h:1145-1160)
  <dir>/textures/<tex>.hdr      Radiance RGBE, flat scanlines (stb_image reads them)

and reads back the `.pfm` images the app saves (yocto_image.cpp:1527-1555).
"""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

from .scene import Camera, Scene


def _floats(a):
    """float32 values as the shortest decimal strings that round-trip."""
    return [float(np.format_float_positional(np.float32(x), unique=True, trim="-")) if np.isfinite(x) else float(x)
            for x in np.asarray(a, np.float32).ravel()]


def write_ply_lines(path: Path, shape) -> None:
    """Polyline records: maximal runs of consecutive segments (i, i+1), (i+1, i+2), ..."""
    lines = shape.lines
    breaks = np.nonzero(lines[1:, 0] != lines[:-1, 1])[0] + 1
    starts = np.concatenate([[0], breaks])
    ends = np.concatenate([breaks, [len(lines)]])
    assert (ends - starts).max() + 1 <= 255, "PLY list sizes are uint8"
    nv = len(shape.positions)
    with open(path, "wb") as f:
        f.write(("ply\nformat binary_little_endian 1.0\ncomment written by yocto_hair_b200.scenefiles\n"
                 f"element vertex {nv}\nproperty float x\nproperty float y\nproperty float z\n"
                 "property float nx\nproperty float ny\nproperty float nz\nproperty float radius\n"
                 f"element line {len(starts)}\nproperty list uchar int vertex_indices\nend_header\n").encode())
        v = np.concatenate([shape.positions, shape.normals, shape.radius[:, None]], axis=1).astype("<f4")
        f.write(v.tobytes())
        if np.all(ends - starts == ends[0] - starts[0]):  # uniform records: vectorised
            k = int(ends[0] - starts[0]) + 1
            rec = np.zeros(len(starts), dtype=[("n", "u1"), ("idx", "<i4", (k,))])
            rec["n"] = k
            rec["idx"][:, :-1] = lines[:, 0].reshape(len(starts), k - 1)
            rec["idx"][:, -1] = lines[ends - 1, 1]
            f.write(rec.tobytes())
        else:
            for s, e in zip(starts, ends):
                idx = np.concatenate([lines[s:e, 0], lines[e - 1:e, 1]]).astype("<i4")
                f.write(bytes([len(idx)]) + idx.tobytes())


def write_ply_triangles(path: Path, shape) -> None:
    nv, nt = len(shape.positions), len(shape.triangles)
    has_n = len(shape.normals) == nv
    with open(path, "wb") as f:
        f.write(("ply\nformat binary_little_endian 1.0\n"
                 f"element vertex {nv}\nproperty float x\nproperty float y\nproperty float z\n"
                 + ("property float nx\nproperty float ny\nproperty float nz\n" if has_n else "")
                 + f"element face {nt}\nproperty list uchar int vertex_indices\nend_header\n").encode())
        v = np.concatenate([shape.positions, shape.normals], axis=1) if has_n else shape.positions
        f.write(v.astype("<f4").tobytes())
        rec = np.zeros(nt, dtype=[("n", "u1"), ("idx", "<i4", (3,))])
        rec["n"], rec["idx"] = 3, shape.triangles
        f.write(rec.tobytes())


def write_hdr(path: Path, rgb: np.ndarray) -> None:
    """Radiance .hdr with flat (uncompressed) RGBE scanlines."""
    h, w, _ = rgb.shape
    m = rgb.max(axis=2)
    e = np.where(m > 1e-32, np.floor(np.log2(np.maximum(m, 1e-38))) + 1, 0)
    scale = np.where(m > 1e-32, 256.0 / np.exp2(e), 0)
    out = np.zeros((h, w, 4), np.uint8)
    out[..., :3] = np.clip(rgb * scale[..., None], 0, 255).astype(np.uint8)
    out[..., 3] = np.where(m > 1e-32, e + 128, 0).astype(np.uint8)
    assert not (out[0, 0, 0] == 2 and out[0, 0, 1] == 2), "first pixel would read as an RLE marker"
    with open(path, "wb") as f:
        f.write(f"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y {h} +X {w}\n".encode())
        f.write(out.tobytes())


def write_scene(directory, name: str, scene: Scene, camera: Camera) -> Path:
    d = Path(directory)
    (d / "shapes").mkdir(parents=True, exist_ok=True)
    (d / "textures").mkdir(parents=True, exist_ok=True)
    for i, s in enumerate(scene.shapes):
        (write_ply_lines if len(s.lines) else write_ply_triangles)(d / "shapes" / f"shape{i:03d}.ply", s)
    for i, t in enumerate(scene.textures):
        write_hdr(d / "textures" / f"tex{i:03d}.hdr", t.colorf)
    js = {"asset": {"generator": "yocto_hair_b200.scenefiles"}, "cameras": {}, "environments": {}, "materials": {}, "objects": {}}
    js["cameras"]["default"] = {"frame": _floats(camera.frame), "lens": _floats([camera.lens])[0],
                                "aspect": _floats([camera.aspect])[0], "film": _floats([camera.film_size])[0],
                                "focus": _floats([camera.focus])[0], "aperture": _floats([camera.aperture])[0]}
    for i, e in enumerate(scene.environments):
        js["environments"][f"env{i:03d}"] = {"frame": _floats(e.frame), "emission": _floats(e.emission)}
        if e.emission_tex >= 0:
            js["environments"][f"env{i:03d}"]["emission_tex"] = f"tex{e.emission_tex:03d}"
    for i, m in enumerate(scene.materials):
        js["materials"][f"mat{i:03d}"] = {
            "emission": _floats(m.emission), "color": _floats(m.color), "specular": m.specular, "roughness": m.roughness,
            "metallic": m.metallic, "ior": _floats([m.ior])[0], "transmission": m.transmission, "opacity": _floats([m.opacity])[0],
            "thin": bool(m.thin), "scattering": _floats(m.scattering), "scanisotropy": _floats([m.scanisotropy])[0],
            "trdepth": _floats([m.trdepth])[0], "eumelanin": _floats([m.eumelanin])[0], "pheomelanin": _floats([m.pheomelanin])[0],
            "sigma_a": _floats(m.sigma_a), "beta_m": _floats([m.beta_m])[0], "beta_n": _floats([m.beta_n])[0],
            "alpha": _floats([m.alpha])[0], "eta": _floats([m.eta])[0]}
    for i, o in enumerate(scene.objects):  # JSON objects iterate alphabetically: the names keep the order
        js["objects"][f"obj{i:03d}"] = {"frame": _floats(o.frame), "shape": f"shape{o.shape:03d}", "material": f"mat{o.material:03d}"}
    path = d / f"{name}.json"
    path.write_text(json.dumps(js, indent=1))
    return path


def read_pfm(path) -> np.ndarray:
    """PF / Pf image -> (H, W, C) float32.  The reference's writer stores the rows top to
    bottom, little endian, scale -1 (yocto_image.cpp:1527-1555); no flip is applied."""
    with open(path, "rb") as f:
        kind = f.readline().strip()
        w, h = (int(x) for x in f.readline().split())
        scale = float(f.readline())
        c = 3 if kind == b"PF" else 1
        data = np.frombuffer(f.read(w * h * c * 4), dtype="<f4" if scale < 0 else ">f4").reshape(h, w, c)
    return data.astype(np.float32)
