"""The two callers of the hot path in ``/root/reference/chrysalis/utils.py``, written against the device
path of this package:

* ``integrate_adatas`` (contract: utils.py:161-250) - multi-sample integration.  Moran's I stays per sample (own W, own
  centring, own gene filter), but every sample that still needs scoring goes through ONE block-diagonal launch
  (``core.detect_svgs_multi`` -> ``chr_moran_blockdiag_f32``) instead of a Python loop of ``detect_svgs`` calls.
* ``estimate_compartments`` (contract: utils.py:115-136) - the archetype-count sweep, run as concurrent device fits
  (``core.aa_sweep``).

The field contract is the reference's: ``.obs[sample_col]`` (categorical), ``.var['spatially_variable']`` = union over the
samples, ``.varm['spatially_variable']`` / ``.varm["Moran's I"]`` with one column per sample (plus gene ids / symbols),
``.obsm['entropy']`` and ``.uns['chr_aa']['RSSs']``.  Colour helpers, the harmony wrapper and plotting are out of scope.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import pandas as pd
import scipy.sparse as sp

from . import core
from .anndata_lite import AnnDataLite

SVG_COL, MORAN_COL = "spatially_variable", "Moran's I"


# ---------------------------------------------------------------------------------------------------------------------
# integrate_adatas
# ---------------------------------------------------------------------------------------------------------------------
def _label_sample(ad, name, sample_col: str) -> None:
    """Per-sample bookkeeping before scoring (utils.py:188-207): the image entry of ``.uns['spatial']`` is re-keyed to the
    sample name, the sample label goes to ``.obs``, gene symbols are kept aside and ENSEMBL ids (when present) become the
    var names."""
    images = ad.uns.get("spatial") if hasattr(ad.uns, "get") else None
    if images is not None:
        if len(images) != 1:
            raise AssertionError("integrate_adatas expects exactly one library in .uns['spatial'] per sample")
        (old_key, entry), = list(images.items())
        if old_key != name:
            del images[old_key]
        images[name] = entry
    if sample_col in ad.obs.columns:
        raise Exception('sample_id_col is already present in adata.obs, specify another column.')   # the reference's message
    ad.obs[sample_col] = name
    if "gene_symbols" not in ad.var.columns:
        ad.var["gene_symbols"] = ad.var_names
    if "gene_ids" in ad.var.columns:
        ad.var.index = pd.Index(ad.var["gene_ids"])


def _stack_samples(samples: dict) -> AnnDataLite:
    """What ``anndata.concat(samples, index_unique='-', uns_merge='unique', merge='first')`` (utils.py:223) yields for the
    duck-typed containers used here: observations stacked in sample order with '-<sample>' appended to their names, the
    genes common to all samples (inner join, order of the first sample), ``.obsm`` entries every sample has, ``.var``
    columns from the first sample that carries them, ``.uns`` entries kept when unambiguous."""
    keys = list(samples)
    genes = pd.Index(samples[keys[0]].var_names)
    for k in keys[1:]:
        genes = genes[genes.isin(pd.Index(samples[k].var_names))]
    X = sp.vstack([sp.csr_matrix(samples[k].X)[:, pd.Index(samples[k].var_names).get_indexer(genes)] for k in keys], format="csr")
    obs_names = [f"{o}-{k}" for k in keys for o in samples[k].obs_names]
    shared_obsm = set.intersection(*(set(samples[k].obsm.keys()) for k in keys))
    merged = AnnDataLite(X, var_names=list(genes), obs_names=obs_names,
                         obsm={m: np.concatenate([np.asarray(samples[k].obsm[m]) for k in keys], axis=0) for m in sorted(shared_obsm)})
    merged.obs = pd.concat([samples[k].obs for k in keys], axis=0).set_axis(pd.Index(obs_names), axis=0)
    for k in keys:                                                       # merge='first'
        table = samples[k].var
        for col in table.columns.difference(merged.var.columns, sort=False):
            merged.var[col] = table[col].reindex(genes).to_numpy()
    for k in keys:                                                       # uns_merge='unique'
        for key, value in samples[k].uns.items():
            merged.uns.setdefault(key, value)
    return merged


def _fold_sample_tables(adata, prefix: str, varm_key: str) -> None:
    """Moves the per-sample ``.var`` columns ``<prefix><sample>`` into one ``.varm[varm_key]`` table (utils.py:234-248),
    next to the gene ids / symbols when the object has them."""
    cols = [c for c in adata.var.columns if prefix in c]
    table = adata.var[cols].copy()
    for extra in ("gene_ids", "gene_symbols"):
        if extra in adata.var.columns:
            table[extra] = adata.var[extra]
    adata.varm[varm_key] = table
    adata.var = adata.var.drop(columns=cols)


def integrate_adatas(adatas: List, sample_names: Sequence = None, calculate_svgs: bool = False, sample_col: str = "sample",
                     **kwargs):
    """
    Integrate multiple samples (drop-in for ``chrysalis.integrate_adatas``, /root/reference/chrysalis/utils.py:161-250).

    If ``.var['gene_ids']`` is present it replaces the gene symbols as var names.  ``.var['spatially_variable']`` of the
    result is the union over the samples; the sample-wise tables go to ``.varm['spatially_variable']`` and
    ``.varm["Moran's I"]``.

    :param adatas: list of AnnData (or duck-typed) objects.
    :param sample_names: one name per sample (default 0, 1, ...).
    :param calculate_svgs: score every sample that has no ``.var['spatially_variable']`` yet (``detect_svgs`` semantics per
        sample; samples sharing a gene list are scored together in one block-diagonal launch).
    :param sample_col: ``.obs`` column that receives the sample labels.
    :param kwargs: keyword arguments of ``detect_svgs``.
    :return: the integrated object.
    """
    names = list(np.arange(len(adatas))) if sample_names is None else list(sample_names)
    assert len(adatas) == len(names)

    for ad, name in zip(adatas, names):
        _label_sample(ad, name, sample_col)

    unscored = [ad for ad in adatas if SVG_COL not in ad.var.columns]
    if unscored:
        if not calculate_svgs:
            raise Exception('spatially_variable column is not found in adata.var. Run `chrysalis.detect_svgs` '
                            'first or set the calculate_svgs argument to True.')                     # the reference's message
        core.detect_svgs_multi(unscored, **kwargs)                       # per-sample W / centring / filter, one launch

    samples = {}
    for ad, name in zip(adatas, names):
        ad.var[f"{SVG_COL}_{name}"] = ad.var[SVG_COL]
        ad.var[f"{MORAN_COL}_{name}"] = ad.var[MORAN_COL]
        samples[name] = ad

    adata = _stack_samples(samples)
    adata.obs[sample_col] = adata.obs[sample_col].astype("category")

    # a gene is spatially variable in the cohort when any sample (or a pre-existing column) says so
    flags = adata.var[[c for c in adata.var.columns if SVG_COL in c]]
    adata.var[SVG_COL] = (flags == True).any(axis=1).to_numpy()          # noqa: E712  (NaN - gene filtered in a sample - is not True)

    _fold_sample_tables(adata, f"{SVG_COL}_", SVG_COL)
    _fold_sample_tables(adata, f"{MORAN_COL}_", MORAN_COL)
    return adata


# ---------------------------------------------------------------------------------------------------------------------
# estimate_compartments
# ---------------------------------------------------------------------------------------------------------------------
def estimate_compartments(adata, n_pcs=20, range_archetypes=(3, 50), max_iter=10):
    """
    Archetypal analysis for every archetype count in ``range(*range_archetypes)`` (drop-in for
    /root/reference/chrysalis/utils.py:115-136): ``.obsm['entropy']`` receives the per-spot entropy of the proportions
    (spots x counts), ``.uns['chr_aa']['RSSs']`` the residual sum of squares per count (what ``plot_rss`` reads).

    The fits share one upload of the embedding and run concurrently on the device (``core.aa_sweep``); each is the fit
    ``aa`` would make (3 restarts, ``tol=0.001``, ``random_state=42``).
    """
    if "chr_X_pca" not in adata.obsm.keys():
        raise ValueError(".obsm['chr_X_pca'] cannot be found, run chrysalis_pca first.")
    counts = list(range(range_archetypes[0], range_archetypes[1]))
    sweep = core.aa_sweep(np.asarray(adata.obsm["chr_X_pca"])[:, :n_pcs], counts, max_iter=max_iter)
    adata.obsm["entropy"] = np.stack([sweep[a]["entropy"] for a in counts], axis=1)
    adata.uns.setdefault("chr_aa", {})["RSSs"] = {a: sweep[a]["rss"] for a in counts}
