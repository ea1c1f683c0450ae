"""Drop-in ``detect_svgs`` / ``pca`` / ``aa`` on B200.

Same signatures, same AnnData fields and same error behaviour as
``/root/reference/chrysalis/core.py`` (``detect_svgs`` :12-92, ``pca`` :95-137, ``aa`` :140-208);
the numeric bodies (scanpy / libpysal / esda / sklearn / archetypes calls) are replaced by
hand-written sm_100a kernels reached through the C ABI (``include/chrysalis_b200.h``).
"""
from __future__ import annotations

import os
import threading
import time

import numpy as np
import pandas as pd
import torch

from . import device as dv
from . import distributed as dd


class _perf:
    """``CHR_PERF=1``: wall seconds of a public call (device synchronised on both sides) stashed under
    ``adata.uns['chr_perf'][stage]`` together with what the stage reports about itself."""

    def __init__(self, adata, stage):
        self.adata, self.stage, self.on = adata, stage, os.environ.get("CHR_PERF", "0") == "1"

    def __enter__(self):
        if self.on:
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            self.t0 = time.perf_counter()
        return self

    def __exit__(self, exc_type, exc, tb):
        if self.on and exc_type is None:
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            rec = {"seconds": time.perf_counter() - self.t0, "n_obs": int(self.adata.X.shape[0]), "n_vars": int(self.adata.X.shape[1])}
            if torch.cuda.is_available():
                rec["device"] = torch.cuda.get_device_name()
                rec["peak_device_bytes"] = int(torch.cuda.max_memory_allocated())
            self.adata.uns.setdefault("chr_perf", {})[self.stage] = rec
        return False


def _make_unique(names) -> pd.Index:
    """``AnnData.var_names_make_unique()`` (core.py:54): 2nd+ duplicates get '-1', '-2', ..."""
    names = pd.Index(names)
    if names.is_unique:
        return names
    out = np.array(names, dtype=object)
    seen = set(out)
    counter = {}
    dup = names.duplicated(keep="first")
    for i in np.flatnonzero(dup):
        base = out[i]
        c = counter.get(base, 0)
        while True:
            c += 1
            cand = f"{base}-{c}"
            if cand not in seen:
                break
        counter[base] = c
        seen.add(cand)
        out[i] = cand
    return pd.Index(out)


# Limits of the device kernels that the reference does not have (documented deviations; checked up front so that a
# call outside them raises a clear ValueError instead of failing inside the library)
MAX_PCS = 128            # chr_eigh_topk_f64: Ritz pairs kept
MAX_ARCHETYPES = 64      # aa.cu kAaMaxA
MAX_AA_DIMS = 64         # aa.cu kAaMaxD

_SIDE_STREAMS = {}
_POOL_LOCK = threading.Lock()
SVG_TRACE = []     # CHR_SVG_TRACE=1: (label, host seconds since the call started) of the last svg_stage call


def _mark(t0, label):
    if t0 is not None:
        SVG_TRACE.append((label, round((time.perf_counter() - t0) * 1e3, 3)))



_W_BUILDER = []


def _w_builder():
    """The host thread that assembles W next to the calling thread (one per process, kept)."""
    from concurrent.futures import ThreadPoolExecutor

    with _POOL_LOCK:
        if not _W_BUILDER:
            _W_BUILDER.append(ThreadPoolExecutor(max_workers=1, thread_name_prefix="chr-w"))
        return _W_BUILDER[0]


def _side_stream():
    """A second CUDA stream (one per device, kept: its allocator pool is reused across calls) for work that
    overlaps the upload of the counts.  None without a device: the device calls then raise ChrysalisNativeError."""
    if not torch.cuda.is_available():
        return None
    dev = torch.cuda.current_device()
    if dev not in _SIDE_STREAMS:
        _SIDE_STREAMS[dev] = torch.cuda.Stream(device=dev)
    return _SIDE_STREAMS[dev]


class _stream_ctx:
    def __init__(self, stream):
        self.ctx = torch.cuda.stream(stream) if stream is not None else None

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


def svg_stage(X, spatial, min_spots: float, neighbors: int, geary: bool, already_log1p: bool, timing: bool = False,
              shard: bool = True, permutations: int = 0, seed: int | None = None, local_shard: bool = False):
    """Device part of ``detect_svgs``: filter -> normalise/log1p -> kNN W -> Moran's I.

    Returns host arrays ``(keep_mask (G,), I (G',), C (G',) or None)``; also usable directly
    (bench, multi-GPU driver).  Mirrors core.py:53-76.

    Under an initialised ``torch.distributed`` group (one process per GPU) every rank takes a
    contiguous range of the genes: the filter is per gene (local), the per-spot totals of
    ``normalize_total`` are all-reduced, W is built on every rank, Moran's I needs no exchange,
    and the per-gene results are all-gathered so every rank returns the full arrays.
    ``shard=False`` keeps the call local to this rank (every rank works on its own matrix).
    ``local_shard=True``: ``X`` already is this rank's gene shard (a loader that reads column ranges per rank): nothing
    is sliced, the exchanges are the same; the returned arrays cover the genes of all ranks in rank order.
    ``permutations=P > 0`` (esda's ``Moran(..., permutations=P)``) appends a dict of the simulated
    quantities (p_sim, EI_sim, VI_sim, seI_sim, z_sim, p_z_sim; host arrays over the scored genes).
    """
    n, g_all = X.shape
    t0 = None
    if os.environ.get("CHR_SVG_TRACE"):
        SVG_TRACE.clear()
        t0 = time.perf_counter()
    rank, size = dd.world() if shard else (0, 1)
    lo, hi = (0, g_all) if local_shard else dd.shard_range(g_all, rank, size)
    # While the counts are in flight: W (core.py:61-65, needs the coordinates only) and the zero fill of the dense
    # matrix run on a side stream; the side stream waits for everything enqueued so far, not for the upload enqueued next.
    main = torch.cuda.current_stream() if torch.cuda.is_available() else None
    side = _side_stream()
    # core.py:61-62  coordinates (x, -y): uploaded BEFORE the counts, otherwise this small copy queues behind the
    # multi-GB transfer on the copy engine and the side-stream work below cannot start until the upload has finished
    # (the sign flip of y happens on the device: the host makes no copy of the coordinates)
    pts = np.asarray(spatial)
    if pts.dtype.kind not in "fiu":
        pts = pts.astype(np.float64)
    xy = dv.upload_array(pts, torch.float64)
    xy[:, 1].neg_()
    if side is not None:
        side.wait_stream(main)
    _mark(t0, "coords uploaded")
    # counts already in pinned (or device) memory upload asynchronously: enqueue that first, then the side-stream work.
    # Anything else (scipy CSR in pageable memory) keeps the host busy filling bounce buffers: the side-stream work is
    # launched before it.  A rank's gene shard of a host matrix is selected while uploading (no host-side slicing).
    split = size > 1 and not local_shard

    def upload_counts():
        if not split:
            return dv.upload_csr(X)
        if isinstance(X, (dv.DeviceCSR, dv.PinnedCSR, dv.CompactCSR)):
            return dv.csr_columns(dv.upload_csr(X), lo, hi)
        col_map = np.full(g_all, -1, dtype=np.int32)
        col_map[lo:hi] = np.arange(hi - lo, dtype=np.int32)
        return dv.upload_csr_selected(X, col_map)

    async_upload = not split and dv.host_arrays_pinned(X)
    csr = None
    if async_upload:
        csr = upload_counts()
        _mark(t0, "counts upload enqueued")
    def build_w():
        # core.py:64-65  kNN weights, row-standardised.  kNN on the spots as given: distance ties go to the lower ORIGINAL
        # spot id (knn.cu), so W does not depend on the internal row order and is the W local_morans_i builds
        with _stream_ctx(side):
            bbox = dv.spatial_bbox(xy)
            nbr0 = dv.knn_build(xy, neighbors, bbox)
            # square lattice with k = 8 (Visium HD / Stereo-seq bins): lattice layout + stencil kernel; otherwise rows in
            # Hilbert order + the quad plan of W
            lat = None
            if neighbors == 8 and os.environ.get("CHR_MORAN_LATTICE", "1") != "0":
                lat = dv.lattice_plan(xy, nbr0, bbox, simple_lists=n > neighbors)     # knn_build: k distinct others per spot
            if lat is not None:
                rank_of, nbr, plan, n_rows = lat.row_of_spot, None, None, lat.n_rows
            else:
                order = dv.spatial_order(xy, bbox)          # spatially coherent row order of the dense matrix
                rank_of = torch.empty_like(order)
                rank_of[order] = torch.arange(n, device=order.device)
                nbr = rank_of.to(torch.int32)[nbr0[order].long()].contiguous()      # relabelled into the row order
                plan = dv.moran_plan(nbr)                    # W as the device plan the Moran kernel streams against
                n_rows = n
            # on a lattice the kernel is fed from the CSR counts themselves (no dense matrix); permutation inference and every
            # other W stream a dense matrix, whose zero fill runs here, under the upload
            use_csr = lat is not None and not permutations and os.environ.get("CHR_MORAN_CSR", "1") != "0"
            dense_buf = None
            if torch.cuda.is_available() and not use_csr:
                dense_buf = torch.empty((n_rows, dv.dense_ld(hi - lo)), dtype=torch.float32, device="cuda")
                dense_buf.zero_()
        return lat, rank_of, nbr, plan, n_rows, use_csr, dense_buf

    # The W build is a chain of ~100 short launches with a few read-backs: host-bound (6 - 9 ms at 700k spots).  It runs on a
    # second host thread (and the side stream) while this thread moves the counts and applies the gene filter; on one GPU it
    # hides under the upload, on a gene shard of a multi-GPU run it is the longest item of the call.
    w_job = None
    if side is not None:
        device_index = torch.cuda.current_device()

        def build_w_on_worker():
            torch.cuda.set_device(device_index)
            return build_w()

        w_job = _w_builder().submit(build_w_on_worker)
    else:
        w_built = build_w()
    if csr is None:
        csr = upload_counts()
        _mark(t0, "counts uploaded (bounce ring)")

    def join_w():
        built = w_job.result() if w_job is not None else w_built
        lat, rank_of, nbr, plan, n_rows, use_csr, dense_buf = built
        if side is not None:
            main.wait_stream(side)
            keep_alive = [rank_of] + ([dense_buf] if dense_buf is not None else []) + ([nbr, plan.buf] if plan is not None else []) + \
                ([lat.bands, lat.chunks, lat.corr, lat.absent, lat.pilot_rows] if lat is not None else [])
            for t in keep_alive:
                t.record_stream(main)
        _mark(t0, "W build joined (side stream, second host thread)")
        return built

    # core.py:53  sc.pp.filter_genes(min_cells=int(len(adata) * min_spots))
    gene_nnz = dv.csr_gene_nnz(csr)
    keep = gene_nnz >= int(n * min_spots)
    gene_col = torch.where(keep, torch.cumsum(keep.to(torch.int32), 0, dtype=torch.int32) - 1,
                           torch.full_like(gene_nnz, -1))
    n_keep = int(keep.sum().item())
    _mark(t0, "gene filter known on the host (upload finished)")
    gather = dd.allgather_ragged if size > 1 else (lambda t, n_total=None: t)
    reduce_ = dd.allreduce_sum_ if size > 1 else (lambda t: t)
    keep_all = gather(keep, None if local_shard else g_all)
    if int(keep_all.sum().item()) == 0:
        join_w()
        return keep_all.cpu().numpy(), np.empty(0), (np.empty(0) if geary else None)
    # core.py:55-57  normalize_total + log1p on the filtered copy unless the caller already did
    scale = None
    if not already_log1p:
        scale = dv.median_scale(reduce_(dv.csr_spot_totals(csr, gene_col)))
    lat, rank_of, nbr, plan, n_rows, use_csr, dense_buf = join_w()
    unsorted = None
    if n_keep > 0 and use_csr and csr.indices.numel() < 2 ** 32 - 1:
        # core.py:59 + 71-76 in one kernel: the producers of the stencil kernel build the tiles from the stored entries
        I, C, _, unsorted = dv.moran_lattice_csr(csr, gene_col, n_keep, scale, lat, apply_log1p=not already_log1p, geary=geary, timing=timing)
        sim = None
        if int(unsorted.item()):                                        # columns not sorted inside a row: the dense path does not care
            unsorted, use_csr = None, False
    else:
        use_csr = False
    if n_keep > 0 and not use_csr:
        # core.py:59  gene_matrix = ad.to_df()  (dense float32, in lattice layout / Hilbert row order)
        if dense_buf is None and torch.cuda.is_available():
            dense_buf = torch.zeros((n_rows, dv.dense_ld(hi - lo)), dtype=torch.float32, device="cuda")
        dense = dv.csr_densify(csr, gene_col, n_keep, scale, rank_of, apply_log1p=not already_log1p, out=dense_buf)
        # core.py:71-76  the gene loop
        if lat is not None:
            I, C, _ = dv.moran_lattice(dense, lat, n_keep, geary=geary, timing=timing)
            sim = dv.moran_perm_lattice(dense, lat, n_keep, permutations, seed) if permutations else None
        else:
            I, C, _ = dv.moran_dense(dense, nbr, n_keep, geary=geary, timing=timing, plan=plan)
            sim = dv.moran_permutation_test(dense, nbr, n_keep, permutations, seed, plan) if permutations else None
    elif n_keep == 0:
        I = torch.empty(0, dtype=torch.float64, device=keep.device)
        C = torch.empty(0, dtype=torch.float64, device=keep.device) if geary else None
        sim = {k: torch.empty(0, dtype=torch.float64, device=keep.device)
               for k in ("p_sim", "EI_sim", "VI_sim", "seI_sim", "z_sim", "p_z_sim")} if permutations else None
    _mark(t0, "densify + Moran enqueued")
    I = gather(I)
    C = gather(C) if geary else None
    out = (keep_all.cpu().numpy(), I.cpu().numpy(), (C.cpu().numpy() if geary else None))
    _mark(t0, "results on the host")
    if permutations:
        out += ({k: gather(v).cpu().numpy() for k, v in sim.items() if k != "I"},)
    return out


def svg_stage_multi(samples, min_spots: float, neighbors: int, geary: bool):
    """``detect_svgs``'s device part for SEVERAL samples that share one gene list, in one Moran launch
    (``chr_moran_blockdiag_f32``): what the per-sample loop of ``integrate_adatas`` computes
    (/root/reference/chrysalis/utils.py:210-212 calling core.py:53-76 once per sample) - every sample keeps its own gene
    filter, its own ``normalize_total`` target, its own kNN graph, mean and variance; W over the stacked spots is block
    diagonal.  ``samples``: list of ``(X, spatial, already_log1p)``.  Returns a list of ``(keep, I, C)`` host arrays per
    sample, as ``svg_stage`` does.

    Under an initialised ``torch.distributed`` group the samples are dealt out to the ranks (sample b -> rank b mod size,
    one sample per GPU for an 8-sample cohort on 8 GPUs), each rank scores its samples in one launch, and the per-sample
    results (8 bytes per gene) are all-gathered."""
    B = len(samples)
    g = samples[0][0].shape[1]
    assert all(s[0].shape[1] == g for s in samples), "svg_stage_multi: the samples must share one gene list"
    rank, size = dd.world()
    mine = [b for b in range(B) if b % size == rank]
    results = {}
    if mine:
        block_rows, block_n, pad_rows = dv.block_layout([samples[b][0].shape[0] for b in mine])
        n_rows = int(block_rows[-1])
        csrs, keeps, scales, row_maps = [], [], [], []
        nbr_all = torch.arange(n_rows, dtype=torch.int32, device="cuda")[:, None].repeat(1, neighbors)   # padding rows list themselves
        for t, b in enumerate(mine):
            X, spatial, logged = samples[b]
            n_b = X.shape[0]
            csr = dv.upload_csr(X)
            keep = dv.csr_gene_nnz(csr) >= int(n_b * min_spots)                      # core.py:53 on this sample
            gene_col = torch.where(keep, torch.cumsum(keep.to(torch.int32), 0, dtype=torch.int32) - 1,
                                   torch.full((g,), -1, dtype=torch.int32, device=keep.device))
            scales.append(None if logged else dv.median_scale(dv.csr_spot_totals(csr, gene_col)))   # core.py:55-57
            pts = np.array(spatial, dtype=np.float64, copy=True)
            pts[:, 1] = pts[:, 1] * -1                                                # core.py:61-62
            xy = dv.to_device(pts, torch.float64)
            bbox = dv.spatial_bbox(xy)
            order = dv.spatial_order(xy, bbox)
            rank_of = torch.empty_like(order)
            rank_of[order] = torch.arange(n_b, device=order.device)
            nbr = rank_of.to(torch.int32)[dv.knn_build(xy, neighbors, bbox)[order].long()]   # core.py:64-65
            r0 = int(block_rows[t])
            nbr_all[r0:r0 + n_b] = nbr + r0
            csrs.append(csr); keeps.append(keep); row_maps.append(rank_of + r0)
        union = torch.stack(keeps).any(0)
        col_u = torch.where(union, torch.cumsum(union.to(torch.int32), 0, dtype=torch.int32) - 1,
                            torch.full((g,), -1, dtype=torch.int32, device=union.device))
        g_u = int(union.sum().item())
        if g_u > 0:
            dense = torch.zeros((n_rows, dv.dense_ld(g_u)), dtype=torch.float32, device="cuda")
            for csr, scale, rows, (_, _, logged) in zip(csrs, scales, row_maps, [samples[b] for b in mine]):
                dv.csr_densify(csr, col_u, g_u, scale, rows, apply_log1p=not logged, out=dense)    # core.py:59 per sample
            plan = dv.moran_plan(nbr_all.contiguous())
            I, C = dv.moran_blockdiag(dense, plan, block_rows, block_n, torch.from_numpy(pad_rows).cuda(), g_u, geary=geary)
        for t, b in enumerate(mine):
            pos = col_u[keeps[t]].long()                                              # this sample's kept genes inside the union
            I_b = I[t][pos].cpu().numpy() if g_u > 0 else np.empty(0)
            C_b = (C[t][pos].cpu().numpy() if g_u > 0 else np.empty(0)) if geary else None
            results[b] = (keeps[t].cpu().numpy(), I_b, C_b)
    if size > 1:
        for part in dd.allgather_objects(results):
            results.update(part)
    return [results[b] for b in range(B)]


class _UniqueNames:
    """Stand-in for the names of all ranks when they are known to be distinct (only their count and uniqueness are used)."""
    is_unique = True

    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n


def _write_svg_fields(adata, keep, I, C, top_svg, min_morans, geary, all_names=None, offset=0):
    """The tail of ``detect_svgs`` (core.py:78-92) for one AnnData given the scored genes.

    ``keep`` / ``I`` / ``C`` may cover more genes than this AnnData holds (gene-sharded AnnData: the genes of all ranks
    in rank order, ``all_names`` their names, this AnnData's genes starting at ``offset``).
    With unique gene names the reference's pandas statements reduce to array arithmetic (a descending sort, a count above
    the threshold, index-aligned writes) and are done that way; duplicate names take the statements themselves, so that
    ``var_names_make_unique`` renames and index alignment behave exactly as in the reference."""
    local = pd.Index(adata.var_names)
    names_all = local if all_names is None else all_names
    g_local = len(local)
    if names_all.is_unique:
        keep = np.asarray(keep, dtype=bool)
        I = np.asarray(I, dtype=np.float64)
        full = np.full(len(keep), np.nan)
        full[keep] = I
        adata.var["Moran's I"] = full[offset:offset + g_local]                     # core.py:80: NaN for filtered genes
        if geary:
            full_c = np.full(len(keep), np.nan)
            full_c[keep] = np.asarray(C, dtype=np.float64)
            adata.var["Geary's C"] = full_c[offset:offset + g_local]                # core.py:85
        above = I > min_morans                                                     # strict, NaN is not above
        chosen = np.zeros(len(keep), dtype=bool)
        if min(top_svg, len(I)) < int(above.sum()):                                # core.py:88: len(df[:top_svg]) < len(df[df > min])
            order = np.argsort(-I, kind="stable")                                  # descending, NaN last (core.py:79)
            chosen[np.flatnonzero(keep)[order[:top_svg]]] = True
        else:
            chosen[np.flatnonzero(keep)[above]] = True
        adata.var['spatially_variable'] = chosen[offset:offset + g_local]          # core.py:89,92
        return
    names = _make_unique(names_all[keep])
    moran_df = pd.DataFrame(data=I, index=names, columns=["Moran's I"])
    moran_df = moran_df.sort_values(ascending=False, by="Moran's I")
    adata.var["Moran's I"] = moran_df["Moran's I"]
    if geary:
        geary_df = pd.DataFrame(data=C, index=names, columns=["Geary's C"])
        geary_df = geary_df.sort_values(ascending=False, by="Geary's C")
        adata.var["Geary's C"] = geary_df["Geary's C"]
    if len(moran_df[:top_svg]) < len(moran_df[moran_df["Moran's I"] > min_morans]):
        chosen = set(moran_df[:top_svg].index)
    else:
        chosen = set(moran_df[moran_df["Moran's I"] > min_morans].index)
    # membership of every var name in the chosen set (core.py:89,92 builds the same booleans with a list comprehension)
    adata.var['spatially_variable'] = np.asarray(local.isin(list(chosen)), dtype=bool)


def detect_svgs_multi(adatas, min_spots: float = 0.05, top_svg: int = 1000, min_morans: float = 0.20, neighbors: int = 6,
                      geary: bool = False):
    """``detect_svgs`` for a list of samples: same arguments, same per-sample results and ``.var`` fields as calling
    ``detect_svgs`` on every element (what ``integrate_adatas(..., calculate_svgs=True)`` does, utils.py:210-212), but
    samples that share a gene list are scored together in one block-diagonal launch."""
    assert 0 < min_spots < 1
    groups = {}
    for i, ad in enumerate(adatas):
        groups.setdefault(tuple(ad.var_names), []).append(i)
    for idx in groups.values():
        res = svg_stage_multi([(adatas[i].X, adatas[i].obsm['spatial'], "log1p" in adatas[i].uns_keys()) for i in idx],
                              min_spots, neighbors, geary)
        for i, (keep, I, C) in zip(idx, res):
            _write_svg_fields(adatas[i], keep, I, C, top_svg, min_morans, geary)


def detect_svgs(adata, min_spots: float = 0.05, top_svg: int = 1000, min_morans: float = 0.20, neighbors: int = 6,
                geary: bool = False):
    """
    Calculate spatial autocorrelation (Moran's I) to define spatially variable genes.

    Drop-in for ``chrysalis.detect_svgs`` (/root/reference/chrysalis/core.py:12-92).

    :param adata: AnnData (or duck-typed equivalent) of shape ``n_obs`` x ``n_vars``; spatial coordinates in
        ``.obsm['spatial']``.
    :param min_spots: only genes expressed in at least this fraction of capture spots are scored.
    :param top_svg: cutoff for top ranked spatially variable genes.
    :param min_morans: Moran's I cutoff; only active when it retains fewer genes than ``top_svg``.
    :param neighbors: number of nearest neighbours used for Moran's I.
    :param geary: also compute Geary's C into ``.var["Geary's C"]``.
    :return: None; updates ``.var["Moran's I"]`` and ``.var["spatially_variable"]``.
    """
    assert 0 < min_spots < 1
    with _perf(adata, "detect_svgs"):
        _detect_svgs(adata, min_spots, top_svg, min_morans, neighbors, geary)


def _detect_svgs(adata, min_spots, top_svg, min_morans, neighbors, geary):
    # Multi-GPU extension (no counterpart in the reference): under torch.distributed an AnnData flagged with
    # .uns['chr_gene_shard'] holds this rank's GENES only (all spots) - a loader that reads column ranges per rank.
    # Scores and the top_svg / min_morans rule are then global over the genes of all ranks; every rank writes its own .var.
    local_shard = bool(dict(adata.uns).get('chr_gene_shard', False)) and dd.world()[1] > 1
    keep, I, C = svg_stage(adata.X, adata.obsm['spatial'], min_spots, neighbors, geary,
                           already_log1p="log1p" in adata.uns_keys(), local_shard=local_shard)

    # names of the filtered copy after var_names_make_unique (core.py:53-54)
    all_names, offset = None, 0
    if local_shard:
        # the names themselves only matter when some are duplicated (var_names_make_unique renames): exchange one 64-bit
        # hash per name (a fixed-size tensor all-gather) and fall back to the names when two hashes coincide
        rank = dd.world()[0]
        local = pd.Index(adata.var_names)
        h = pd.util.hash_array(local.to_numpy(), categorize=False).view(np.int64)     # siphash of every name, same key on every rank
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        hashes, lens = dd.allgather_ragged(torch.from_numpy(h.copy()).to(dev), return_lens=True)
        offset = int(sum(lens[:rank]))
        hashes = hashes.cpu().numpy()
        if np.unique(hashes).size != hashes.size:
            parts = dd.allgather_objects(list(local))
            all_names = pd.Index([v for part in parts for v in part])
        else:
            all_names = _UniqueNames(int(hashes.size))
    _write_svg_fields(adata, keep, I, C, top_svg, min_morans, geary, all_names=all_names, offset=offset)


def morans_i_permutation(adata, min_spots: float = 0.05, neighbors: int = 6, permutations: int = 999, seed: int | None = None):
    """
    Moran's I with permutation inference for every gene: what ``esda.moran.Moran(y, w, permutations=P)``
    reports per gene (the reference's ``detect_svgs`` hard-codes ``permutations=0`` at core.py:72; the
    authors call esda directly for this, ``article/A2_human_lymph_node/morans_i.ipynb:186``).

    Same gene filter, normalisation and W as ``detect_svgs``.  Updates ``.var`` with ``"Moran's I"``,
    ``"Moran's I p_sim"``, ``"Moran's I z_sim"``, ``"Moran's I EI_sim"``, ``"Moran's I VI_sim"`` (NaN for filtered genes).
    """
    assert 0 < min_spots < 1
    keep, I, _, sim = svg_stage(adata.X, adata.obsm['spatial'], min_spots, neighbors, False,
                                already_log1p="log1p" in adata.uns_keys(), permutations=int(permutations), seed=seed)
    names = _make_unique(pd.Index(adata.var_names)[keep])
    adata.var["Moran's I"] = pd.Series(I, index=names)
    for key in ("p_sim", "z_sim", "EI_sim", "VI_sim"):
        adata.var[f"Moran's I {key}"] = pd.Series(sim[key], index=names)


def local_morans_i(adata, keys, neighbors: int = 6, permutations: int = 999, seed: int | None = None, obsm_key: str | None = None):
    """
    Local Moran's I (LISA) per spot for a few variables: ``esda.moran.Moran_Local(y, w, permutations=P)`` with the
    W of ``detect_svgs`` (core.py:61-65), as the authors run it next to the global statistic
    (``article/A2_human_lymph_node/morans_i.ipynb:150``; ``chrysalis/core.py`` has no such call).

    ``keys``: gene names (columns of ``adata.X`` as stored - normalise first, like for esda), or with ``obsm_key``
    column indices of ``adata.obsm[obsm_key]`` (e.g. the compartment scores ``'chr_aa'``).
    Writes ``.obsm['local_morans_i']``, ``['local_morans_q']`` (1 HH, 2 LH, 3 LL, 4 HL) and, with permutations,
    ``['local_morans_p_sim']`` (all [n, len(keys)]) and ``.uns['local_morans']``.
    """
    keys = list(keys)
    n = len(adata)
    if obsm_key is None:
        from scipy.sparse import issparse
        cols = pd.Index(adata.var_names).get_indexer(keys)
        assert (cols >= 0).all(), "local_morans_i: unknown gene name"
        sub = adata.X[:, cols]
        vals = np.asarray(sub.todense() if issparse(sub) else sub, dtype=np.float32)
    else:
        vals = np.asarray(adata.obsm[obsm_key], dtype=np.float32)[:, [int(c) for c in keys]]
    pts = np.array(adata.obsm['spatial'], dtype=np.float64, copy=True)
    pts[:, 1] = pts[:, 1] * -1
    xy = dv.to_device(pts, torch.float64)
    nbr = dv.knn_build(xy, neighbors, dv.spatial_bbox(xy))
    # one independent 64-bit key per variable (SeedSequence.spawn): adjacent variables never share redraw streams
    seeds = [int(c.generate_state(1, np.uint64)[0]) for c in np.random.SeedSequence(seed).spawn(len(keys))]
    Is, q, ps = (np.empty((n, len(keys)), dtype=dt) for dt in (np.float32, np.int32, np.float32))
    yd = dv.to_device(np.ascontiguousarray(vals.T), torch.float32)
    for j in range(len(keys)):
        res = dv.local_moran(yd[j], nbr, permutations=int(permutations), seed=seeds[j])
        Is[:, j], q[:, j] = res["Is"].cpu().numpy(), res["q"].cpu().numpy()
        if permutations:
            ps[:, j] = res["p_sim"].cpu().numpy()
    adata.obsm['local_morans_i'] = Is
    adata.obsm['local_morans_q'] = q
    if permutations:
        adata.obsm['local_morans_p_sim'] = ps
    adata.uns['local_morans'] = {'keys': keys, 'neighbors': neighbors, 'permutations': int(permutations), 'obsm_key': obsm_key}


def pca_stage(X, sel: np.ndarray, n_pcs: int, timing: bool = False):
    """Device part of ``pca``: densify the SVG columns -> column means -> tcgen05 Gram ->
    eigensolver -> scores.  Returns host arrays (scores (n, r) float32, components (r, S) float32,
    explained_variance_ratio (r,) float32).  Mirrors sklearn's PCA(arpack) as called at core.py:127-128:
    centre, top-r right singular vectors, ``svd_flip`` (v-based), scores = U * S.

    Under an initialised ``torch.distributed`` group every rank takes a contiguous range of the
    SPOTS: column sums and the partial Grams are all-reduced, the small eigenproblem is solved on
    every rank, the scores of the local spots are all-gathered.
    """
    n = X.shape[0]
    s = int(sel.sum())
    # sklearn's arpack solver needs 0 < n_components < min(n_samples, n_features)  (_pca.py)
    if not 0 < n_pcs < min(n, s):
        raise ValueError(f"n_components={n_pcs} must be strictly less than min(n_samples, n_features)={min(n, s)} "
                         "with svd_solver='arpack'")
    if n_pcs > MAX_PCS:       # deviation from the reference (no such limit there): the eigensolver keeps <= 128 Ritz pairs
        raise ValueError(f"n_pcs={n_pcs} exceeds this implementation's limit of {MAX_PCS} principal components")
    rank, size = dd.world()
    lo, hi = dd.shard_range(n, rank, size)
    col_map = np.where(sel, np.cumsum(sel) - 1, -1).astype(np.int32)
    if isinstance(X, (dv.DeviceCSR, dv.PinnedCSR)) or (size == 1 and dv.host_arrays_pinned(X)):
        # resident or page-locked arrays: one asynchronous copy of everything, columns picked on the device
        csr = dv.csr_rows(dv.upload_csr(X), lo, hi)
        gene_col = dv.to_device(col_map)
    else:
        # ordinary host arrays: only the entries of the SVG columns (and of this rank's spots) cross PCIe
        csr = dv.upload_csr_selected(X, col_map, lo, hi)
        gene_col = torch.arange(s, dtype=torch.int32, device=csr.data.device)
    P = dv.csr_densify(csr, gene_col, s, None, None, apply_log1p=False)     # core.py:126  .X.todense()
    mean = dd.allreduce_sum_(dv.column_sums(P, s)) / n
    C = dd.allreduce_sum_(dv.gram_centered(P, s, mean, timing=timing)) / (n - 1)   # covariance (ddof=1)
    total_var = torch.diagonal(C).sum()
    evals, V, worst = dv.eigh_topk(C, n_pcs)
    if worst > 1e-6:          # Krylov depth capped before the Ritz residuals met the bound: say so, do not hand back silently
        import warnings
        warnings.warn(f"chrysalis_b200.pca: eigensolver stopped at a relative residual of {worst:.2e} "
                      f"({s} features, {n_pcs} components); the trailing components may be inaccurate", RuntimeWarning)
    # svd_flip(u, v, u_based_decision=False): largest-|.| entry of every component is positive
    idx = V.abs().argmax(dim=1)
    signs = torch.sign(V[torch.arange(V.shape[0], device=V.device), idx])
    signs[signs == 0] = 1.0
    V = V * signs[:, None]
    scores = dd.allgather_ragged(dv.pca_scores(P, s, mean, V), n)
    ratio = (evals / total_var).to(torch.float32)
    return dv.to_host(scores), V.to(torch.float32).cpu().numpy(), ratio.cpu().numpy()


def pca(adata, n_pcs: int = 50):
    """
    Perform PCA (Principal Component Analysis) to calculate PCA coordinates, loadings, and variance decomposition.

    Drop-in for ``chrysalis.pca`` (/root/reference/chrysalis/core.py:95-137).  Spatially variable genes need to
    be defined in ``.var['spatially_variable']`` using ``detect_svgs``.

    :param adata: AnnData (or duck-typed equivalent) of shape ``n_obs`` x ``n_vars``.
    :param n_pcs: Number of principal components to be calculated.
    :return: None; adds PCs to ``.obsm['chr_X_pca']`` and ``.uns['chr_pca']`` with 'variance_ratio',
        'loadings' and 'features'.
    """
    sel = np.asarray(adata.var['spatially_variable'] == True)  # noqa: E712  (same test as core.py:126)
    with _perf(adata, "pca"):
        scores, components, ratio = pca_stage(adata.X, sel, n_pcs)
    adata.obsm['chr_X_pca'] = scores

    features = list(np.asarray(adata.var_names)[sel])
    if 'chr_pca' not in adata.uns.keys():
        adata.uns['chr_pca'] = {'variance_ratio': ratio,
                                'loadings': components,
                                'features': features}
    else:
        adata.uns['chr_pca']['variance_ratio'] = ratio
        adata.uns['chr_pca']['loadings'] = components
        adata.uns['chr_pca']['features'] = features


def _furthest_sum(X: np.ndarray, n_archetypes: int, rs: np.random.RandomState) -> np.ndarray:
    """FurthestSum seeding (Morup & Hansen 2012) as archetypes 0.4.x draws it: a random start, greedy
    picks of the point with the largest summed distance to the chosen set, then 10 re-pick rounds.
    Host statement of what ``chr_aa_furthest_sum`` does on the device (``aa_stage`` uses the device
    kernel; this stays as the readable definition the tests pin it against)."""
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    sq = np.einsum("ij,ij->i", X, X)

    def dist(i):
        return np.sqrt(np.maximum(sq - 2.0 * (X @ X[i]) + sq[i], 0.0))

    picks = [int(rs.randint(n))]
    free = np.ones(n, dtype=bool)
    free[picks[0]] = False
    total = np.zeros(n)
    cur = picks[0]
    for step in range(1, n_archetypes + 11):
        if step > n_archetypes - 1:
            oldest = picks.pop(0)
            total -= dist(oldest)
            free[oldest] = True
        total += dist(cur)
        cand = np.flatnonzero(free)
        cur = int(cand[np.argmax(total[cand])])
        picks.append(cur)
        free[cur] = False
    return np.array(picks[:n_archetypes], dtype=np.int64)


def _aa_upload(X: np.ndarray) -> torch.Tensor:
    """Embedding -> float64 device tensor.  numpy upcasts the float32 scores to float64 inside archetypes; here the (exact)
    widening happens on the device, so half the bytes cross PCIe and the host never builds the float64 copy."""
    X = np.asarray(X)
    if X.dtype == np.float32:
        return dv.upload_array(np.ascontiguousarray(X), torch.float64)
    return dv.upload_array(np.ascontiguousarray(X, dtype=np.float64), torch.float64)


_RESTART_WORKERS = {}
_RESTART_LOCK = threading.Lock()


def _restart_workers(device, n: int):
    """Host threads and CUDA streams that drive concurrent restarts: created once per device and kept (a fresh stream per
    call would make the caching allocator open a new small-block pool each time)."""
    from concurrent.futures import ThreadPoolExecutor

    key = (device.index if device.index is not None else torch.cuda.current_device(), n)
    with _RESTART_LOCK:
        if key not in _RESTART_WORKERS:
            _RESTART_WORKERS[key] = (ThreadPoolExecutor(max_workers=n, thread_name_prefix="chr-aa"),
                                     [torch.cuda.Stream(device=device) for _ in range(n)])
        return _RESTART_WORKERS[key]


def _aa_fit_device(Xd: torch.Tensor, n_archetypes: int, max_iter: int, n_init: int, tol: float, random_state: int, init_rows=None,
                   restart_workers: int = None):
    """``archetypes.AA(n_archetypes, n_init, max_iter, tol, random_state).fit`` on a device-resident embedding.

    The restarts are independent fits (their initial rows are drawn up front, so the RNG stream does not depend on them);
    ``restart_workers`` host threads drive one CUDA stream each, so the launch gaps and the convergence read-backs of one
    restart are covered by the kernels of the others.  Each restart is the same deterministic chain of launches either way."""
    n, d = Xd.shape
    if n_archetypes > MAX_ARCHETYPES or d > MAX_AA_DIMS:
        raise ValueError(f"aa: this implementation supports n_archetypes <= {MAX_ARCHETYPES} and <= {MAX_AA_DIMS} embedding "
                         f"dimensions (got n_archetypes={n_archetypes}, n_pcs={d})")
    if not n_archetypes < n:
        raise ValueError(f"aa: n_archetypes={n_archetypes} must be smaller than the number of observations ({n})")
    rs = np.random.RandomState(random_state)
    # initial archetypes of every restart first: the RNG stream (one randint + one choice per restart) does not depend
    # on the fits
    used = []
    for r in range(n_init):
        if init_rows is not None:
            rows = np.asarray(init_rows[r], dtype=np.int64)
        else:
            rows = dv.aa_furthest_sum(Xd, n_archetypes, int(rs.randint(n)))
            rs.choice(n_archetypes, n, replace=True)            # alphas0 draw: unused by the updates, keeps the RNG stream
        used.append(rows)

    def fit_restart(rows, st=None):
        st = dv.AAState(Xd, n_archetypes) if st is None else st
        st.init_from_rows(rows)
        rss_prev, rss, it = np.inf, np.inf, 0
        for it in range(max_iter):
            rss = st.iterate()
            if abs(rss_prev - rss) < tol:
                break
            rss_prev = rss
        return {"alphas": st.alphas, "archetypes": st.Z, "rss": rss, "n_iter": it + 1 if max_iter else 0}

    workers = int(os.environ.get("CHR_AA_RESTART_WORKERS", n_init)) if restart_workers is None else int(restart_workers)
    if workers > 1 and len(used) > 1:
        device = Xd.device
        # the restarts' buffers come from the caller's stream (the caching allocator cannot hand blocks cached for one stream
        # to another: allocating 150 MB per restart on fresh streams would go to cudaMalloc, or to a cache flush on a full device)
        states = [dv.AAState(Xd, n_archetypes) for _ in used]
        pool, streams = _restart_workers(device, min(workers, len(used)))
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(device))          # Xd (and the seeding kernels) precede every restart

        def on_own_stream(job):
            i, rows, st = job
            torch.cuda.set_device(device)
            stream = streams[i % len(streams)]
            stream.wait_event(ready)
            with torch.cuda.stream(stream):
                f = fit_restart(rows, st)
            stream.synchronize()                                 # the caller's stream may reuse the buffers afterwards
            return f

        fits = list(pool.map(on_own_stream, [(i, rows, st) for i, (rows, st) in enumerate(zip(used, states))]))
    else:
        fits = [fit_restart(rows) for rows in used]
    best = None
    for f in fits:                                              # first restart with the smallest RSS, as a sequential loop keeps it
        if best is None or f["rss"] < best["rss"]:
            best = f
    best["init_rows"] = used
    return best


def aa_stage(X: np.ndarray, n_archetypes: int, max_iter: int = 200, n_init: int = 3, tol: float = 0.001,
             random_state: int = 42, init_rows=None):
    """Device part of ``aa``: ``archetypes.AA(n_archetypes, n_init, max_iter, tol, random_state).fit(X)``
    (core.py:186-191).  ``init_rows``: optional list of n_init index arrays (matched initialisation).
    Returns a dict with alphas (N, A), archetypes (A, d), rss, n_iter and the initial rows used."""
    best = _aa_fit_device(_aa_upload(X), n_archetypes, max_iter, n_init, tol, random_state, init_rows)
    return {"alphas": dv.to_host(best["alphas"]), "archetypes": best["archetypes"].cpu().numpy(), "rss": best["rss"],
            "n_iter": best["n_iter"], "init_rows": best["init_rows"]}


def aa_sweep(X: np.ndarray, archetype_counts, max_iter: int = 10, n_init: int = 3, tol: float = 0.001, random_state: int = 42,
             workers: int = 4):
    """The fits of ``estimate_compartments`` (/root/reference/chrysalis/utils.py:123-130: one ``arch.AA(n_archetypes=a,
    n_init=3, max_iter, tol=0.001, random_state=42).fit`` per archetype count) as CONCURRENT device fits: the embedding is
    uploaded once, ``workers`` host threads drive one CUDA stream each (a fit of a few thousand spots is a chain of short
    launches that leaves the GPU mostly idle on its own).  Returns {a: {"rss", "entropy" (per spot, natural log), "n_iter"}}."""
    from concurrent.futures import ThreadPoolExecutor

    Xd = _aa_upload(X)
    counts = [int(a) for a in archetype_counts]
    device = Xd.device
    torch.cuda.synchronize(device)

    def one(a):
        torch.cuda.set_device(device)
        stream = torch.cuda.Stream(device=device)
        with torch.cuda.stream(stream):
            f = _aa_fit_device(Xd, a, max_iter, n_init, tol, random_state, restart_workers=1)
            p = f["alphas"] / f["alphas"].sum(1, keepdim=True)          # scipy.stats.entropy normalises, then -sum p log p
            ent = -(torch.where(p > 0, p * torch.log(p), torch.zeros_like(p))).sum(1)
            out = {"rss": f["rss"], "entropy": ent.cpu().numpy(), "n_iter": f["n_iter"]}
        stream.synchronize()
        return a, out

    if workers <= 1 or len(counts) <= 1:
        return dict(one(a) for a in counts)
    with ThreadPoolExecutor(max_workers=min(workers, len(counts))) as pool:
        return dict(pool.map(one, counts))


def aa(adata, n_archetypes: int = 8, pca_key: str = None, n_pcs: int = None, max_iter: int = 200):
    """
    Run archetypal analysis on the low-dimensional embedding.

    Drop-in for ``chrysalis.aa`` (/root/reference/chrysalis/core.py:140-208).  Needs ``chrysalis.pca`` first.

    :param adata: AnnData (or duck-typed equivalent) of shape ``n_obs`` x ``n_vars``.
    :param n_archetypes: Number of archetypes (tissue compartments) to be identified.
    :param pca_key: alternative PCA input from ``.obsm``; by default ``.obsm['chr_X_pca']``.
    :param n_pcs: Number of PCs to be used (default ``n_archetypes - 1``).
    :param max_iter: Maximum number of iterations.
    :return: None; adds ``.obsm['chr_aa']`` and ``.uns['chr_aa']`` with 'archetypes', 'alphas', 'loadings', 'RSS'.
    """
    if not isinstance(n_archetypes, int):
        raise TypeError
    elif n_archetypes < 2:
        raise ValueError(f"n_archetypes cannot be less than 2.")

    if n_pcs is None:
        pcs = n_archetypes - 1
    else:
        pcs = n_pcs

    if pca_key is None:
        emb = adata.obsm['chr_X_pca'][:, :pcs]
    else:
        emb = adata.obsm[pca_key][:, :pcs]
    with _perf(adata, "aa"):
        fit = aa_stage(np.asarray(emb), n_archetypes, max_iter=max_iter)

    adata.obsm['chr_aa'] = fit["alphas"]

    # archetype loadings (core.py:197)
    aa_loadings = np.dot(fit["archetypes"], adata.uns['chr_pca']['loadings'][:pcs, :])

    if 'chr_aa' not in adata.uns.keys():
        adata.uns['chr_aa'] = {'archetypes': fit["archetypes"],
                               'alphas': fit["alphas"],
                               'loadings': aa_loadings,
                               'RSS': fit["rss"]}
    else:
        adata.uns['chr_aa']['archetypes'] = fit["archetypes"]
        adata.uns['chr_aa']['alphas'] = fit["alphas"]
        adata.uns['chr_aa']['loadings'] = aa_loadings
        adata.uns['chr_aa']['RSS'] = fit["rss"]
