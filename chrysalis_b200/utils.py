"""The two callers of the hot path that live in ``/root/reference/chrysalis/utils.py``:
``integrate_adatas`` (multi-sample: per-sample Moran's I with per-sample W, union of the SVG sets -
utils.py:161-250) and ``estimate_compartments`` (archetype-count sweep - utils.py:115-136).
Colour helpers, harmony wrapper and plotting stay out of scope.
"""
from __future__ import annotations

from typing import List

import numpy as np
import pandas as pd
import scipy.sparse as sp

from . import core
from .anndata_lite import AnnDataLite


def _concat_inner(adatas_dict: dict) -> AnnDataLite:
    """``anndata.concat(adatas_dict, index_unique='-', uns_merge='unique', merge='first')`` for duck-typed
    inputs: inner join on var_names (order of the first sample), obs names suffixed with '-<sample>',
    ``.var`` columns taken from the first sample that has them, ``.obsm`` keys common to all stacked."""
    names = list(adatas_dict)
    first = adatas_dict[names[0]]
    common = pd.Index(first.var_names)
    for k in names[1:]:
        common = common[common.isin(pd.Index(adatas_dict[k].var_names))]
    blocks, obs_names, obs_frames = [], [], []
    for k in names:
        ad = adatas_dict[k]
        cols = pd.Index(ad.var_names).get_indexer(common)
        blocks.append(sp.csr_matrix(ad.X)[:, cols])
        obs_names += [f"{o}-{k}" for o in ad.obs_names]
        obs_frames.append(ad.obs)
    obsm_keys = set(first.obsm.keys())
    for k in names[1:]:
        obsm_keys &= set(adatas_dict[k].obsm.keys())
    obsm = {key: np.concatenate([np.asarray(adatas_dict[k].obsm[key]) for k in names], axis=0) for key in sorted(obsm_keys)}
    out = AnnDataLite(sp.vstack(blocks, format="csr"), obsm=obsm, var_names=list(common), obs_names=obs_names)
    obs = pd.concat(obs_frames, axis=0)
    obs.index = pd.Index(obs_names)
    out.obs = obs
    # merge='first': every var column from the first sample that carries it, restricted to the common genes
    for k in names:
        v = adatas_dict[k].var
        for c in v.columns:
            if c not in out.var.columns:
                out.var[c] = v[c].reindex(common).to_numpy()
    # uns_merge='unique': keys whose values are identical (or present once)
    for k in names:
        for key, val in adatas_dict[k].uns.items():
            if key not in out.uns:
                out.uns[key] = val
    return out


def integrate_adatas(adatas: List, sample_names: List[str] = None, calculate_svgs: bool = False, sample_col: str = 'sample',
                     **kwargs):
    """
    Integrate multiple samples (drop-in for ``chrysalis.integrate_adatas``, utils.py:161-250).

    Moran's I is computed PER SAMPLE with that sample's own W and its own centring (``detect_svgs`` per input);
    ``.var['spatially_variable']`` of the result is the union over the samples, the sample-wise columns move to
    ``.varm['spatially_variable']`` and ``.varm["Moran's I"]``.  If ``.var['gene_ids']`` is present it replaces the
    gene symbols as var names.

    :param adatas: list of AnnData (or duck-typed) objects.
    :param sample_names: list of sample names (default 0, 1, ...).
    :param calculate_svgs: run ``detect_svgs`` for every sample that has no ``spatially_variable`` column.
    :param sample_col: ``.obs`` column that receives the sample labels.
    :param kwargs: keyword arguments for ``detect_svgs``.
    :return: the integrated object.
    """
    if sample_names is None:
        sample_names = np.arange(len(adatas))
    assert len(adatas) == len(sample_names)

    adatas_dict = {}
    for ad, name in zip(adatas, sample_names):
        # replace .uns['spatial'] with the specified sample name
        if 'spatial' in ad.uns.keys():
            assert len(ad.uns['spatial'].keys()) == 1
            curr_key = list(ad.uns['spatial'].keys())[0]
            ad.uns['spatial'][name] = ad.uns['spatial'][curr_key]
            if name != curr_key:
                del ad.uns['spatial'][curr_key]

        # check if column is already used
        if sample_col not in ad.obs.columns:
            ad.obs[sample_col] = name
        else:
            raise Exception('sample_id_col is already present in adata.obs, specify another column.')

        if 'gene_symbols' not in ad.var.columns:
            ad.var['gene_symbols'] = ad.var_names

        if 'gene_ids' in ad.var.columns:
            ad.var.index = pd.Index(ad.var['gene_ids'])

        # check if SVGs are already present
        if 'spatially_variable' not in ad.var.columns:
            if calculate_svgs:
                core.detect_svgs(ad, **kwargs)
            else:
                raise Exception('spatially_variable column is not found in adata.var. Run `chrysalis.detect_svgs` '
                                'first or set the calculate_svgs argument to True.')

        ad.var[f'spatially_variable_{name}'] = ad.var['spatially_variable']
        ad.var[f"Moran's I_{name}"] = ad.var["Moran's I"]

        adatas_dict[name] = ad

    # concat samples
    adata = _concat_inner(adatas_dict)
    adata.obs[sample_col] = adata.obs[sample_col].astype('category')
    # get SVGs for all samples
    svg_columns = [c for c in adata.var.columns if 'spatially_variable' in c]
    svg_list = [list(adata.var[c][adata.var[c] == True].index) for c in svg_columns]  # noqa: E712

    # union of SVGs
    spatially_variable = set().union(*svg_list)
    adata.var['spatially_variable'] = [x in spatially_variable for x in adata.var_names]

    # save sample-wise spatially_variable and Morans's I columns
    sv_cols = [x for x in adata.var.columns if 'spatially_variable_' in x]
    adata.varm['spatially_variable'] = adata.var[sv_cols].copy()
    if 'gene_ids' in adata.var.columns:
        adata.varm['spatially_variable']['gene_ids'] = adata.var['gene_ids']
    if 'gene_symbols' in adata.var.columns:
        adata.varm['spatially_variable']['gene_symbols'] = adata.var['gene_symbols']
    adata.var = adata.var.drop(columns=sv_cols)

    mi_cols = [x for x in adata.var.columns if "Moran's I_" in x]
    adata.varm["Moran's I"] = adata.var[mi_cols].copy()
    if 'gene_ids' in adata.var.columns:
        adata.varm["Moran's I"]['gene_ids'] = adata.var['gene_ids']
    if 'gene_symbols' in adata.var.columns:
        adata.varm["Moran's I"]['gene_symbols'] = adata.var['gene_symbols']
    adata.var = adata.var.drop(columns=mi_cols)

    return adata


def estimate_compartments(adata, n_pcs=20, range_archetypes=(3, 50), max_iter=10):
    """
    Archetypal analysis for a range of archetype counts (utils.py:115-136): stores the per-spot entropy of the
    proportions in ``.obsm['entropy']`` (spots x counts) and the residual sums of squares in
    ``.uns['chr_aa']['RSSs']`` (what ``plot_rss`` reads for the elbow).
    """
    from scipy.stats import entropy

    if 'chr_X_pca' not in adata.obsm.keys():
        raise ValueError(".obsm['chr_X_pca'] cannot be found, run chrysalis_pca first.")

    counts = range(range_archetypes[0], range_archetypes[1])
    entropy_arr = np.zeros((len(counts), len(adata)))
    rss_dict = {}
    emb = np.asarray(adata.obsm['chr_X_pca'][:, :n_pcs])
    for i, a in enumerate(counts):
        fit = core.aa_stage(emb, a, max_iter=max_iter)
        rss_dict[a] = fit["rss"]
        entropy_arr[i, :] = entropy(fit["alphas"].T)

    adata.obsm['entropy'] = entropy_arr.T

    if 'chr_aa' not in adata.uns.keys():
        adata.uns['chr_aa'] = {'RSSs': rss_dict}
    else:
        adata.uns['chr_aa']['RSSs'] = rss_dict
