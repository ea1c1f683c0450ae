"""Seeded synthetic spatial-transcriptomics workloads shaped like BASELINE.json's configs.

The reference's own generator (``/root/reference/article/A1_synthetic_data/data_generator/
tools.py:9-49, 228-304``) needs the Tabula Sapiens atlas and a git-ignored module, so it
cannot run; this module produces data of the same *shape*: K smooth tissue-zone fields over
the capture area, ~10 % "spatial" genes whose rate mixes the zone fields, the rest pure
noise genes, per-spot depth ~ Gamma(25, 1/25) (tools.py:277), per-gene detection
~ Gamma(5, 1/15) (tools.py:247), Poisson counts.

Everything is torch so the same code fills host tensors (tests, oracle comparison) or
device tensors chunk by chunk (the 700k x 18k bench workload never exists on the host).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

# name -> (n_spots, n_genes, neighbours, coords kind, n_zones)   [BASELINE.json "configs"]
CONFIGS = {
    "cfg1_visium": dict(n=4035, g=36601, k=6, coords="hex", zones=8),
    "cfg2_slideseq": dict(n=40000, g=20000, k=8, coords="random", zones=8),
    "cfg3_stereoseq": dict(n=120000, g=25000, k=8, coords="square", zones=8),
    "cfg4_visium_hd": dict(n=700000, g=18000, k=8, coords="square", zones=12),
    "cfg5_cohort": dict(n=5000, g=30000, k=6, coords="hex", zones=8, samples=8),
}


def make_coords(kind: str, n: int, seed: int = 42, pitch: float = 100.0, jitter: float = 1e-3) -> np.ndarray:
    """(n, 2) float64 spot coordinates.

    ``hex``: Visium-style staggered lattice, kept region = seeded blob; ``square``: square
    lattice cropped to n (Visium HD / Stereo-seq bins); ``random``: uniform in a disc
    (Slide-seq pucks).  Lattices get N(0, jitter*pitch) noise so kNN has no distance ties
    (SURVEY.md H2); pass ``jitter=0`` for an exact lattice.
    """
    rng = np.random.default_rng(seed)
    if kind == "random":
        r = np.sqrt(rng.uniform(0, 1, n)) * pitch * math.sqrt(n / math.pi)
        t = rng.uniform(0, 2 * math.pi, n)
        return np.stack([r * np.cos(t), r * np.sin(t)], 1)
    if kind == "square":
        side = int(math.ceil(math.sqrt(n)))
        rr, cc = np.divmod(np.arange(side * side), side)
        xy = np.stack([cc * pitch, rr * pitch], 1).astype(np.float64)[:n]
    elif kind == "hex":
        side = int(math.ceil(math.sqrt(n * 1.25))) + 2
        rr, cc = np.divmod(np.arange(side * side), side)
        xy = np.stack([(cc + 0.5 * (rr % 2)) * pitch, rr * pitch * math.sqrt(3) / 2], 1)
        # blob mask: keep the n positions closest to a wobbly disc centre
        ctr = xy.mean(0)
        d = xy - ctr
        ang = np.arctan2(d[:, 1], d[:, 0])
        rad = np.hypot(d[:, 0], d[:, 1]) / (1.0 + 0.15 * np.sin(3 * ang + 1.0) + 0.1 * np.cos(5 * ang))
        xy = xy[np.sort(np.argsort(rad, kind="stable")[:n])]
    else:
        raise ValueError(kind)
    if jitter:
        xy = xy + rng.normal(0, jitter * pitch, xy.shape)
    return xy


@dataclass
class ExpressionModel:
    """Per-gene / per-zone parameters; rates for any block of spots are derived on demand."""
    centres: torch.Tensor     # (K, B, 2) bump centres in unit square
    widths: torch.Tensor      # (K, B)
    mix: torch.Tensor         # (K, G) zone -> gene mixing (noise genes: all-ones column / K)
    rate: torch.Tensor        # (G,) base rate g_g * exp(a_g)
    lo: torch.Tensor          # (2,) coordinate origin
    span: float
    seed: int

    @property
    def n_genes(self) -> int:
        return self.rate.shape[0]

    def zone_fields(self, xy: torch.Tensor) -> torch.Tensor:
        """(n, K) soft zone memberships (rows sum to 1)."""
        u = (xy.to(torch.float32) - self.lo) / self.span
        d2 = ((u[:, None, None, :] - self.centres[None]) ** 2).sum(-1)          # (n, K, B)
        f = torch.exp(-d2 / (2 * self.widths[None] ** 2)).sum(-1)               # (n, K)
        return torch.softmax(6.0 * f, dim=1)

    def counts(self, xy: torch.Tensor, block_seed: int = 0) -> torch.Tensor:
        """(n_block, G) float32 Poisson counts for the given spots."""
        gen = torch.Generator(device=xy.device).manual_seed(self.seed * 1000003 + block_seed)
        f = self.zone_fields(xy)
        depth = torch._standard_gamma(torch.full((xy.shape[0], 1), 25.0, device=xy.device), generator=gen) / 25.0
        lam = (f @ self.mix) * self.rate[None, :] * depth
        return torch.poisson(lam, generator=gen)


def make_expression_model(xy_all: np.ndarray, n_genes: int, n_zones: int = 8, seed: int = 42,
                          spatial_frac: float = 0.10, log_rate_mean: float = -1.2, log_rate_sd: float = 1.5,
                          device="cpu") -> ExpressionModel:
    rng = np.random.default_rng(seed + 7)
    B = 3
    centres = rng.uniform(0.05, 0.95, (n_zones, B, 2))
    widths = rng.uniform(0.06, 0.18, (n_zones, B))
    n_sp = int(round(n_genes * spatial_frac))
    mix = np.ones((n_zones, n_genes), dtype=np.float32)
    is_sp = np.zeros(n_genes, bool)
    is_sp[rng.choice(n_genes, n_sp, replace=False)] = True
    # spatial genes: marker of 1-2 zones, (near) silent elsewhere
    m = np.full((n_zones, n_sp), 0.05, dtype=np.float32)
    z1 = rng.integers(0, n_zones, n_sp)
    z2 = rng.integers(0, n_zones, n_sp)
    m[z1, np.arange(n_sp)] += rng.uniform(1.5, 4.0, n_sp)
    m[z2, np.arange(n_sp)] += rng.uniform(0.0, 2.0, n_sp)
    mix[:, is_sp] = m
    rate = rng.gamma(5.0, 1.0 / 15.0, n_genes) * 15.0 / 5.0 * np.exp(rng.normal(log_rate_mean, log_rate_sd, n_genes))
    lo = xy_all.min(0)
    span = float((xy_all.max(0) - lo).max()) or 1.0
    t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float32), device=device)
    return ExpressionModel(t(centres), t(widths), t(mix), t(rate), t(lo), span, seed)


def make_counts_csr(xy: np.ndarray, n_genes: int, n_zones: int = 8, seed: int = 42, chunk: int = 8192,
                    device="cpu", **model_kw):
    """Host scipy CSR int32 counts (n, G) generated chunk by chunk on ``device``."""
    import scipy.sparse as sp

    model = make_expression_model(xy, n_genes, n_zones, seed, device=device, **model_kw)
    xy_t = torch.as_tensor(xy, dtype=torch.float32, device=device)
    blocks = []
    for b, i0 in enumerate(range(0, xy.shape[0], chunk)):
        c = model.counts(xy_t[i0:i0 + chunk], b).to(torch.int32).cpu().numpy()
        blocks.append(sp.csr_matrix(c))
    return sp.vstack(blocks, format="csr").astype(np.int32), model


def normalize_log1p_dense(counts: torch.Tensor) -> torch.Tensor:
    """Dense float32 log1p(normalize_total(counts)) for a block whose row totals are complete
    (used to fill the device-resident bench matrix; target = 1e4 instead of the median so blocks
    are independent)."""
    tot = counts.sum(1, keepdim=True).clamp_min(1.0)
    return torch.log1p(counts * (1.0e4 / tot))
