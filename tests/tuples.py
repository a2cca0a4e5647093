"""Seeded BSDF parity inputs: (material, v, frame, wo, wi, rn)
tuples built from one PCG32 stream `make_rng(20240229)` drawn through the checker."""
import numpy as np

from yocto_hair_b200 import hair


def sample_sphere(ruv):
    """yocto_math.h:4847-4852, float32 numpy."""
    ruv = np.asarray(ruv, np.float32)
    z = np.float32(2) * ruv[:, 1] - np.float32(1)
    r = np.sqrt(np.clip(np.float32(1) - z * z, 0, 1)).astype(np.float32)
    phi = np.float32(2 * np.pi) * ruv[:, 0]
    return np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1).astype(np.float32)


def make_tuples(oracle, n, seed=20240229):
    u = oracle.rng_floats(seed, 1, 24 * n).reshape(n, 24)
    wo = sample_sphere(u[:, 0:2])
    wi = sample_sphere(u[:, 2:4])
    v = u[:, 4].copy()
    beta_m = (0.05 + 0.95 * u[:, 5]).astype(np.float32)
    beta_n = (0.05 + 0.95 * u[:, 6]).astype(np.float32)
    k = np.arange(n)
    alpha = np.where(k % 3 == 0, 0.0, np.where(k % 3 == 1, 2.0, 10.0 * u[:, 7])).astype(np.float32)
    mode = (k // 3) % 3
    sigma_a = np.where((mode == 0)[:, None], 4.0 * u[:, 8:11], 0.0).astype(np.float32)
    color = np.where((mode == 1)[:, None], np.maximum(u[:, 11:14], 1e-3), 0.0).astype(np.float32)
    eum = np.where(mode == 2, 8.0 * u[:, 14], 0.0).astype(np.float32)
    pheo = np.where(mode == 2, 2.0 * u[:, 15], 0.0).astype(np.float32)
    mats = hair.make_materials(n, sigma_a=sigma_a, beta_m=beta_m, beta_n=beta_n, alpha=alpha,
                               color=color, eumelanin=eum, pheomelanin=pheo)
    # frames: fixed (n = z, t = x) for the first half, random (non-orthogonal inputs) after
    normals = np.tile(np.array([0, 0, 1], np.float32), (n, 1))
    tangents = np.tile(np.array([1, 0, 0], np.float32), (n, 1))
    half = n // 2
    normals[half:] = sample_sphere(u[half:, 16:18])
    tangents[half:] = sample_sphere(u[half:, 18:20])
    rn = u[:, 20:22].copy()
    return dict(materials=mats, v=v, normals=normals, tangents=tangents, wo=wo, wi=wi, rn=rn)
