#!/usr/bin/env python3
"""Generates the committed golden vectors from the UNMODIFIED reference
(oracle/_ref/libyh_ref.so, built by `make -C oracle ref` where /root/reference
exists).  Run from the repo root:  python tests/golden/make_golden.py

  hair_kat.npz   known-answer vectors of the hair BSDF: the three hand-picked
                 cases of SURVEY.md section 8a followed by 4093 seeded tuples.
  rng_kat.npz    PCG32 rand1f streams for a few (seed, seq) pairs.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "yocto-hair_b200"))
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "tests"))

from oracle import RefOracle  # noqa: E402
from tuples import make_tuples  # noqa: E402
from yocto_hair_b200 import hair  # noqa: E402

ref = RefOracle()
n = 4096
t = make_tuples(ref, n, seed=20240229)
m = np.frombuffer(t["materials"], dtype=np.float32).reshape(n, 12).copy()


def nrm(x):
    x = np.asarray(x, np.float32)
    return x / np.float32(np.sqrt(np.float32((x * x).sum())))


# SURVEY.md 8a known-answer cases
cases = [
    dict(eumelanin=1.3, v=0.3),
    dict(sigma_a=(.06, .10, .20), beta_m=.1, beta_n=.9, alpha=0, v=0.8),
    dict(color=(.8, .4, .05), beta_m=.6, beta_n=.6, v=0.05),
]
for i, c in enumerate(cases):
    row = np.array([0, 0, 0, 0.3, 0.3, 2, 1.55, 0, 0, 0, 0, 0], np.float32)
    if "sigma_a" in c: row[0:3] = c["sigma_a"]
    if "beta_m" in c: row[3] = c["beta_m"]
    if "beta_n" in c: row[4] = c["beta_n"]
    if "alpha" in c: row[5] = c["alpha"]
    if "color" in c: row[7:10] = c["color"]
    if "eumelanin" in c: row[10] = c["eumelanin"]
    m[i] = row
    t["v"][i] = c["v"]
    t["normals"][i] = (0, 0, 1)
    t["tangents"][i] = (1, 0, 0)
    t["wo"][i] = nrm([.2, .3, .9])
    t["wi"][i] = nrm([-.5, .4, -.6])
    t["rn"][i] = (.37, .81)
mats = hair.make_materials(n, sigma_a=m[:, 0:3], beta_m=m[:, 3], beta_n=m[:, 4], alpha=m[:, 5],
                           eta=m[:, 6], color=m[:, 7:10], eumelanin=m[:, 10], pheomelanin=m[:, 11])
b = ref.hair_brdf(mats, t["v"], t["normals"], t["tangents"])
ws = ref.hair_sample(b, t["wo"], t["rn"])
np.savez_compressed(
    ROOT / "tests" / "golden" / "hair_kat.npz",
    sigma_a=m[:, 0:3], beta_m=m[:, 3], beta_n=m[:, 4], alpha=m[:, 5], eta=m[:, 6], color=m[:, 7:10],
    eumelanin=m[:, 10], pheomelanin=m[:, 11], v=t["v"], normals=t["normals"], tangents=t["tangents"],
    wo=t["wo"], wi=t["wi"], rn=t["rn"], brdf=hair.brdf_view(b).copy(),
    f=ref.hair_eval(b, t["wo"], t["wi"]), pdf=ref.hair_pdf(b, t["wo"], t["wi"]), ws=ws,
    f_ws=ref.hair_eval(b, t["wo"], ws), pdf_ws=ref.hair_pdf(b, t["wo"], ws))
pairs = [(961748941, 7), (1301081, 1), (0, 0), (2**63 + 5, 2**40 + 3), (199382389514, 1)]
np.savez_compressed(ROOT / "tests" / "golden" / "rng_kat.npz",
                    pairs=np.array(pairs, dtype=np.uint64),
                    floats=np.stack([ref.rng_floats(s, q, 256) for s, q in pairs]))
print("wrote hair_kat.npz, rng_kat.npz")
g = np.load(ROOT / "tests" / "golden" / "hair_kat.npz")
print(g["f"][:3], g["pdf"][:3], g["ws"][:3])
