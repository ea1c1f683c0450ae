"""Device-side building blocks of the hot path: thin Python over the C ABI.

PyTorch is plumbing here - it owns device buffers, pinned staging and the current stream;
every computation on the path is a call into ``libchrysalis_b200.so`` (hand-written sm_100a
kernels).  There is no CPU fallback: without CUDA or without the library these raise.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib


def _require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.ChrysalisNativeError("chrysalis_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


_BOUNCE_MIN_BYTES = 8 << 20


def to_device(a: np.ndarray, dtype=None) -> torch.Tensor:
    """Host ndarray -> device tensor: a plain copy from pageable memory (about 11 GB/s on this box).  Pinning a fresh
    buffer first only pays when the pinned pool is already warm (the first cudaHostAlloc of a size costs 100-200 ms,
    ``scripts/xfer_test.py``); callers that upload repeatedly hold their data in pinned memory themselves (``PinnedCSR``)."""
    dev = _require_cuda()
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.to(dev)


def to_host(t: torch.Tensor) -> np.ndarray:
    """Device tensor -> fresh ndarray.  Results of 8 MB and more (67 MB of proportions, 84 MB of scores at cfg4) come back through
    the library's pinned ring with all host threads copying out of it (``chr_host_download``): ``t.cpu().numpy()`` takes 31 ms for
    the 67 MB, a copy into a numpy buffer 14 ms - most of it one thread first-touching the pages of the new array."""
    t = t.contiguous()
    nbytes = t.numel() * t.element_size()
    if nbytes < (1 << 20):
        return t.cpu().numpy()
    out = np.empty(tuple(t.shape), dtype=torch.empty(0, dtype=t.dtype).numpy().dtype)
    if nbytes >= _BOUNCE_MIN_BYTES and t.is_cuda:
        _lib.call("chr_host_download", out.ctypes.data, t.data_ptr(), nbytes, _stream())
    else:
        torch.from_numpy(out).copy_(t)
    return out


@dataclass
class DeviceCSR:
    """CSR counts (spots x genes) resident on the device."""
    indptr: torch.Tensor   # int64 [n+1]
    indices: torch.Tensor  # int32 [nnz]
    data: torch.Tensor     # float32 [nnz]
    shape: tuple

    @property
    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.indptr, self.indices, self.data))


@dataclass
class PinnedCSR:
    """CSR counts staged in pinned host memory (what a loader hands to the device path)."""
    indptr: torch.Tensor   # int64 [n+1], pinned
    indices: torch.Tensor  # int32 [nnz], pinned
    data: torch.Tensor     # float32 [nnz], pinned
    shape: tuple

    @property
    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.indptr, self.indices, self.data))




def upload_array(a: np.ndarray, dtype=None) -> torch.Tensor:
    """Host ndarray -> device tensor at the best rate its memory allows: page-locked arrays (e.g. numpy views of pinned
    torch tensors) go as one asynchronous copy; large pageable arrays through the library's multi-threaded pinned
    bounce ring (``chr_host_upload``); small ones as a plain copy.  A dtype change happens on the DEVICE (the host never
    makes a converted copy)."""
    dev = _require_cuda()
    t = torch.from_numpy(np.ascontiguousarray(a))
    nbytes = t.numel() * t.element_size()
    if t.is_pinned():
        d = t.to(dev, non_blocking=True)
    elif nbytes >= _BOUNCE_MIN_BYTES:
        d = torch.empty(t.shape, dtype=t.dtype, device=dev)
        _lib.call("chr_host_upload", d.data_ptr(), t.data_ptr(), nbytes, _stream())
    else:
        d = t.to(dev)
    if dtype is not None and d.dtype != dtype:
        d = d.to(dtype)
    return d


def _as_csr(X):
    import scipy.sparse as sp

    if not sp.issparse(X) or X.format != "csr":
        X = sp.csr_matrix(X)
    return X


@dataclass
class CompactCSR:
    """Raw counts as a loader can stage them for the device: 2 - 3 bytes per stored entry instead of scipy's 8 (PCIe is the floor
    of ``detect_svgs`` end to end).  Values: ``vals`` uint8 with 255 = escape to the overflow list (``over_pos`` positions into
    ``vals``, ``over_val`` float32).  Columns, either
      * ``cols`` uint16 absolute indices (needs n_genes <= 65536), or
      * ``deltas`` uint8: column minus the previous column of the row (first entry of a row: from 0), columns sorted inside
        every row, 255 = escape to the gap list (``gap_pos`` sorted positions, ``gap_val`` int32 true deltas) - at a few per cent
        density a delta >= 255 occurs about once in a million entries.
    Host tensors, page-locked when ``pin``.  ``detect_svgs`` accepts it as ``adata.X`` (expanded on the device by
    ``chr_csr_decode_compact[_delta]``); ``to_scipy()`` gives the ordinary matrix back."""
    indptr: torch.Tensor
    cols: torch.Tensor
    vals: torch.Tensor
    over_pos: torch.Tensor
    over_val: torch.Tensor
    shape: tuple
    deltas: torch.Tensor | None = None
    gap_pos: torch.Tensor | None = None
    gap_val: torch.Tensor | None = None

    def _arrays(self):
        return [t for t in (self.indptr, self.cols, self.vals, self.over_pos, self.over_val, self.deltas, self.gap_pos, self.gap_val)
                if t is not None]

    @property
    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self._arrays())

    @staticmethod
    def encode_deltas(indptr: np.ndarray, indices: np.ndarray):
        """(deltas uint8, gap_pos int64, gap_val int32) of CSR column indices that are sorted inside every row."""
        idx = np.asarray(indices, dtype=np.int64)
        d = idx.copy()
        d[1:] -= idx[:-1]
        starts = np.asarray(indptr[:-1])[np.diff(indptr) > 0]
        d[starts] = idx[starts]
        if d.size and d.min() < 0:
            raise ValueError("CompactCSR: column indices must be sorted inside every row for the delta coding")
        esc = d >= 255
        return np.where(esc, 255, d).astype(np.uint8), np.flatnonzero(esc).astype(np.int64), d[esc].astype(np.int32)

    @staticmethod
    def from_scipy(X, pin: bool = True, delta: bool | None = None) -> "CompactCSR":
        """``delta``: None = delta-code the columns when that is the smaller form (few escapes), True / False = force."""
        X = _as_csr(X)
        n, g = X.shape
        data = np.asarray(X.data)
        if data.size and (data.min() < 0 or not np.array_equal(data, np.floor(data))):
            raise ValueError("CompactCSR holds raw counts: non-negative integers")
        big = data >= 255
        over_pos = np.flatnonzero(big).astype(np.int64)
        vals = np.where(big, 255, data).astype(np.uint8)

        def host(a):
            t = torch.from_numpy(np.ascontiguousarray(a))
            return t.pin_memory() if pin and torch.cuda.is_available() else t

        enc = None
        if delta is not False:
            if not X.has_sorted_indices:
                X = X.sorted_indices()
                data = np.asarray(X.data)
                big = data >= 255
                over_pos = np.flatnonzero(big).astype(np.int64)
                vals = np.where(big, 255, data).astype(np.uint8)
            enc = CompactCSR.encode_deltas(X.indptr, X.indices)
            if delta is None and g <= 65536 and enc[1].size * 12 > X.nnz:      # an escape costs 12 bytes: the uint16 form is smaller
                enc = None
        if enc is None:
            if g > 65536:
                raise ValueError("CompactCSR needs n_genes <= 65536 for uint16 column indices (or sorted rows for the delta coding)")
            return CompactCSR(host(X.indptr.astype(np.int64)), host(X.indices.astype(np.uint16)), host(vals), host(over_pos),
                              host(data[big].astype(np.float32)), (n, g))
        return CompactCSR(host(X.indptr.astype(np.int64)), host(np.empty(0, dtype=np.uint16)), host(vals), host(over_pos),
                          host(data[big].astype(np.float32)), (n, g), host(enc[0]), host(enc[1]), host(enc[2]))

    def columns(self) -> np.ndarray:
        """Absolute int32 column indices (host)."""
        if self.deltas is None:
            return self.cols.numpy().astype(np.int32)
        d = self.deltas.numpy().astype(np.int64)
        d[self.gap_pos.numpy()] = self.gap_val.numpy()
        run = np.cumsum(d)
        indptr = self.indptr.numpy()
        counts = np.diff(indptr)
        before = np.concatenate([[0], run])[indptr[:-1]]                 # running sum in front of every row
        return (run - np.repeat(before, counts)).astype(np.int32)

    def to_scipy(self):
        import scipy.sparse as sp

        data = self.vals.numpy().astype(np.float32)
        data[self.over_pos.numpy()] = self.over_val.numpy()
        return sp.csr_matrix((data, self.columns(), self.indptr.numpy()), shape=self.shape)


def _upload_compact(X: CompactCSR) -> DeviceCSR:
    dev = _require_cuda()
    up = lambda t: t.to(dev, non_blocking=True)
    indptr, vals, opos, oval = up(X.indptr), up(X.vals), up(X.over_pos), up(X.over_val)
    nnz = vals.numel()
    idx = torch.empty(nnz, dtype=torch.int32, device=dev)
    val = torch.empty(nnz, dtype=torch.float32, device=dev)
    if X.deltas is not None:
        deltas, gpos, gval = up(X.deltas), up(X.gap_pos), up(X.gap_val)
        _lib.call("chr_csr_decode_compact_delta", _ptr(indptr), _ptr(deltas), _ptr(vals), X.shape[0], nnz, _ptr(gpos) if gpos.numel() else 0,
                  _ptr(gval) if gval.numel() else 0, gpos.numel(), _ptr(opos) if opos.numel() else 0, _ptr(oval) if oval.numel() else 0,
                  opos.numel(), _ptr(idx), _ptr(val), _stream())
    else:
        cols = up(X.cols)
        _lib.call("chr_csr_decode_compact", _ptr(cols), _ptr(vals), nnz, _ptr(opos) if opos.numel() else 0, _ptr(oval) if oval.numel() else 0,
                  opos.numel(), _ptr(idx), _ptr(val), _stream())
    return DeviceCSR(indptr, idx, val, tuple(X.shape))


def upload_csr(X) -> DeviceCSR:
    """scipy CSR (or anything ``scipy.sparse.csr_matrix`` accepts, incl. dense), PinnedCSR or CompactCSR -> DeviceCSR."""
    if isinstance(X, DeviceCSR):
        return X
    if isinstance(X, CompactCSR):
        return _upload_compact(X)
    if isinstance(X, PinnedCSR):
        dev = _require_cuda()
        return DeviceCSR(X.indptr.to(dev, non_blocking=True), X.indices.to(dev, non_blocking=True),
                         X.data.to(dev, non_blocking=True), tuple(X.shape))
    X = _as_csr(X)
    return DeviceCSR(upload_array(X.indptr, torch.int64), upload_array(X.indices, torch.int32),
                     upload_array(X.data, torch.float32), tuple(X.shape))


def csr_rows(csr: DeviceCSR, lo: int, hi: int) -> DeviceCSR:
    """Rows [lo, hi) of a device CSR (views; one small read-back for the two offsets)."""
    n = csr.shape[0]
    if lo == 0 and hi == n:
        return csr
    b, e = (int(v) for v in csr.indptr[[lo, hi]].cpu())
    return DeviceCSR((csr.indptr[lo:hi + 1] - b).contiguous(), csr.indices[b:e], csr.data[b:e], (hi - lo, csr.shape[1]))


def csr_columns(csr: DeviceCSR, lo: int, hi: int) -> DeviceCSR:
    """Columns [lo, hi) of a device CSR, renumbered from 0 (a rank's gene shard of a matrix that is already on the device;
    plain tensor bookkeeping - host matrices are sliced while uploading instead, ``upload_csr_selected``)."""
    n, g = csr.shape
    if lo == 0 and hi == g:
        return csr
    m = (csr.indices >= lo) & (csr.indices < hi)
    csum = torch.zeros(m.numel() + 1, dtype=torch.int64, device=m.device)
    torch.cumsum(m, 0, out=csum[1:])
    return DeviceCSR(csum[csr.indptr].contiguous(), (csr.indices[m] - lo).contiguous(), csr.data[m].contiguous(), (n, hi - lo))


def host_arrays_pinned(X) -> bool:
    """True when the CSR arrays of ``X`` can be copied asynchronously as they are (pinned / device resident)."""
    if isinstance(X, (PinnedCSR, DeviceCSR)):
        return True
    if isinstance(X, CompactCSR):
        return X.vals.is_pinned() and (X.deltas if X.deltas is not None else X.cols).is_pinned()
    import scipy.sparse as sp

    if not (sp.issparse(X) and X.format == "csr") or not torch.cuda.is_available():
        return False
    return all(torch.from_numpy(a).is_pinned() for a in (X.indices, X.data))


_DATA_KINDS = {np.dtype(np.float32): 0, np.dtype(np.int32): 1, np.dtype(np.float64): 2}


def select_count(X, col_map: np.ndarray, row_lo: int = 0, row_hi: int | None = None):
    """Host count pass of the column selection over rows [row_lo, row_hi) of the scipy CSR ``X``:
    (indptr view int64, sel_indptr int64 [rows + 1], selected nnz).  Runs on the library's host thread pool."""
    import ctypes

    n, g = X.shape
    row_hi = n if row_hi is None else row_hi
    rows = row_hi - row_lo
    indptr = np.ascontiguousarray(X.indptr[row_lo:row_hi + 1], dtype=np.int64)
    col_map = np.ascontiguousarray(col_map, dtype=np.int32)
    assert col_map.shape == (g,) and X.indices.dtype == np.int32 and X.indices.flags.c_contiguous
    sel_indptr = np.empty(rows + 1, dtype=np.int64)
    nnz = ctypes.c_int64(0)
    _lib.call("chr_host_csr_select_count", indptr.ctypes.data, X.indices.ctypes.data, rows, g, col_map.ctypes.data,
              sel_indptr.ctypes.data, ctypes.addressof(nnz))
    return indptr, sel_indptr, int(nnz.value)


def upload_csr_selected(X, col_map: np.ndarray, row_lo: int = 0, row_hi: int | None = None) -> DeviceCSR:
    """Rows [row_lo, row_hi) of the host scipy CSR ``X`` restricted to the columns with ``col_map[c] >= 0`` (renumbered to
    ``col_map[c]``) as a DeviceCSR - only the selected entries cross PCIe (``chr_host_csr_select_*``).  What ``pca`` needs
    of ``adata.X`` (core.py:126) and what a rank needs of it under gene-sharded ``detect_svgs``."""
    dev = _require_cuda()
    X = _as_csr(X)
    if X.indices.dtype != np.int32 or X.data.dtype not in _DATA_KINDS:
        X = X.astype(np.float32) if X.data.dtype not in _DATA_KINDS else X
        if X.indices.dtype != np.int32:
            raise _lib.ChrysalisNativeError("CSR matrices with 64-bit column indices are not supported")
    col_map = np.ascontiguousarray(col_map, dtype=np.int32)
    indptr, sel_indptr, nnz = select_count(X, col_map, row_lo, row_hi)
    rows = sel_indptr.shape[0] - 1
    out_idx = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
    out_val = torch.empty(max(nnz, 1), dtype=torch.float32, device=dev)
    data = np.ascontiguousarray(X.data)
    _lib.call("chr_host_csr_select_upload", indptr.ctypes.data, X.indices.ctypes.data, data.ctypes.data, _DATA_KINDS[data.dtype], rows,
              col_map.ctypes.data, sel_indptr.ctypes.data, out_idx.data_ptr(), out_val.data_ptr(), _stream())
    n_sel = int(col_map.max()) + 1 if col_map.size and col_map.max() >= 0 else 0
    return DeviceCSR(upload_array(sel_indptr, torch.int64), out_idx[:nnz], out_val[:nnz], (rows, n_sel))


# ---- K2 ----------------------------------------------------------------------------------

def csr_gene_nnz(csr: DeviceCSR) -> torch.Tensor:
    n, g = csr.shape
    out = torch.empty(g, dtype=torch.int32, device=csr.data.device)
    _lib.call("chr_csr_gene_nnz", _ptr(csr.indptr), _ptr(csr.indices), _ptr(csr.data), n, g, _ptr(out), _stream())
    return out


def csr_spot_totals(csr: DeviceCSR, gene_col: torch.Tensor) -> torch.Tensor:
    n, _ = csr.shape
    out = torch.empty(n, dtype=torch.float64, device=csr.data.device)
    _lib.call("chr_csr_spot_totals", _ptr(csr.indptr), _ptr(csr.indices), _ptr(csr.data), n, _ptr(gene_col), _ptr(out),
              _stream())
    return out


def dense_ld(n_cols: int) -> int:
    """Row stride of the dense matrix: 128-byte aligned rows."""
    return max(32, (n_cols + 31) // 32 * 32)


def csr_densify(csr: DeviceCSR, gene_col: torch.Tensor, n_cols: int, scale: torch.Tensor | None,
                row_of_spot: torch.Tensor | None, apply_log1p: bool, out: torch.Tensor | None = None) -> torch.Tensor:
    """``out``: optional pre-ZEROED [>= n, ld >= n_cols] float32 buffer (e.g. filled on a side stream while the
    counts were still in flight from the host; more than n rows in lattice layout, where ``row_of_spot`` leaves
    filler rows); the kernel then only scatters."""
    n, _ = csr.shape
    if out is None:
        ld = dense_ld(n_cols)
        dense = torch.empty((n, ld), dtype=torch.float32, device=csr.data.device)
    else:
        dense, ld = out, out.stride(0)
        assert out.dtype == torch.float32 and out.shape[0] >= n and ld >= n_cols and out.stride(1) == 1
    _lib.call("chr_csr_densify_log1p", _ptr(csr.indptr), _ptr(csr.indices), _ptr(csr.data), n, _ptr(gene_col),
              _ptr(scale), _ptr(row_of_spot), _ptr(dense), ld, int(apply_log1p) | (2 if out is not None else 0), _stream())
    return dense


def median_scale(totals: torch.Tensor) -> torch.Tensor:
    """scanpy normalize_total scale factors: 1 / (T_i / median(T[T>0])), T_i == 0 -> T_i = 1."""
    pos = totals[totals > 0]
    if pos.numel() == 0:
        return torch.ones_like(totals)
    srt, _ = torch.sort(pos)
    m = srt.numel()
    after = srt[m // 2] if m % 2 else 0.5 * (srt[m // 2 - 1] + srt[m // 2])
    counts = totals + (totals == 0).to(totals.dtype)
    return 1.0 / (counts / after)


# ---- K1 ----------------------------------------------------------------------------------

def spatial_bbox(xy: torch.Tensor) -> torch.Tensor:
    out = torch.empty(4, dtype=torch.float64, device=xy.device)
    _lib.call("chr_spatial_bbox", _ptr(xy), xy.shape[0], _ptr(out), _stream())
    return out


def hilbert_order(xy: torch.Tensor, bbox: torch.Tensor | None = None) -> torch.Tensor:
    """Permutation that sorts spots along a Hilbert curve (stable; int64 [n])."""
    bbox = spatial_bbox(xy) if bbox is None else bbox
    keys = torch.empty(xy.shape[0], dtype=torch.int64, device=xy.device)
    _lib.call("chr_hilbert_keys", _ptr(xy), xy.shape[0], _ptr(bbox), _ptr(keys), _stream())
    return torch.sort(keys, stable=True).indices


def lattice_pitch_from_nn(xy: torch.Tensor, bbox: torch.Tensor, d1: torch.Tensor) -> float | None:
    """The decision behind ``square_lattice_pitch`` given every spot's nearest-neighbour distance ``d1`` (plain tensor
    arithmetic, device-agnostic).  The pitch candidate is the median of ``d1``; the nearest of four jittered neighbours
    is biased low by a fraction of the jitter, which adds up over hundreds of nodes, so the pitch is refitted by least
    squares on the node numbers, first near the corner, then over the whole lattice.  The spots form a lattice when,
    measured from the bounding-box corner, 98 % of them lie within a tenth of a pitch of a node on both axes."""
    p = d1.median()
    rel = xy - bbox[:2]
    for limit in (64.0, 1e18):
        u = rel / p
        k = u.round() * (u < limit)
        p = (rel * k).sum() / (k * k).sum().clamp_min(1.0)
    u = rel / p
    near = ((u - u.round()).abs() < 0.1).all(dim=1).double().mean()
    extent = (bbox[2:] - bbox[:2]).max() / p
    p_h, near_h, extent_h = (float(v) for v in torch.stack([p, near, extent]).cpu())    # one host round trip
    if not (p_h > 0.0 and near_h >= 0.98 and extent_h < 32000.0):
        return None
    return p_h


def square_lattice_pitch(xy: torch.Tensor, bbox: torch.Tensor) -> float | None:
    """Pitch of the square lattice the spots sit on (Visium HD / Stereo-seq bins), or None (hex lattice, random beads).
    Nearest-neighbour distances come from one exact k = 1 pass of the kNN kernel."""
    if xy.shape[0] < 64:
        return None
    nb = knn_build(xy, 1, bbox)
    d1 = (xy[nb[:, 0].long()] - xy).norm(dim=1)
    return lattice_pitch_from_nn(xy, bbox, d1)


def spatial_order(xy: torch.Tensor, bbox: torch.Tensor | None = None) -> torch.Tensor:
    """Row order of the dense matrix: spots sorted along a Hilbert curve.  On a square lattice the dyadic grid of the
    curve is aligned with the lattice (node i -> cell block [m i, m i + m), m a power of two), so that 4 consecutive
    spots are an exact 2 x 2 block: 6 of the 16 unordered neighbour pairs of a quad then live inside it and the W plan
    needs fewer list rows (11.83 -> 11.50 ms for the cfg4 Moran pass)."""
    bbox = spatial_bbox(xy) if bbox is None else bbox
    p = square_lattice_pitch(xy, bbox)
    if p is None:
        return hilbert_order(xy, bbox)
    b = [float(v) for v in bbox.cpu()]
    w = max(b[2] - b[0], b[3] - b[1]) / p + 1.0
    m = 1
    while (w + 1.0) * m * 2 <= 65536.0:
        m *= 2
    span = 65535.0 * p / m                               # chr_hilbert_keys quantises with 65535 / span: m cells per pitch
    snapped = torch.tensor([b[0] - p / 2, b[1] - p / 2, b[0] - p / 2 + span, b[1] - p / 2 + span], dtype=torch.float64, device=xy.device)
    return hilbert_order(xy, snapped)


def knn_build(xy: torch.Tensor, k: int, bbox: torch.Tensor | None = None) -> torch.Tensor:
    """k nearest other spots of every spot: int32 [n, k] (row-standardised W is implicit)."""
    assert xy.dtype == torch.float64 and xy.is_contiguous() and xy.shape[1] == 2
    n = xy.shape[0]
    bbox = spatial_bbox(xy) if bbox is None else bbox
    ws_bytes = _lib.query("chr_knn_workspace_bytes", n, k)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=xy.device)
    nbr = torch.empty((n, k), dtype=torch.int32, device=xy.device)
    _lib.call("chr_knn_build", _ptr(xy), n, k, _ptr(bbox), _ptr(nbr), _ptr(ws), ws_bytes, _stream())
    return nbr


# ---- K3 / K4 -----------------------------------------------------------------------------

@dataclass
class MoranPlan:
    """Device representation of W for the Moran tile kernel (``chr_moran_plan_build``)."""
    buf: torch.Tensor      # uint8 plan buffer
    chunk_units: int       # largest tile slice (16-byte units) - sizes the kernel's staging buffers
    total_units: int
    n: int
    k: int


def moran_plan(nbr: torch.Tensor) -> MoranPlan | None:
    """Quad plan of the kNN lists ``nbr`` [n, k] int32 (built once per W).  ``None`` for k > 64
    (the per-spot stream kernel handles those straight from ``nbr``)."""
    import ctypes

    assert nbr.dtype == torch.int32 and nbr.is_contiguous() and nbr.dim() == 2
    n, k = nbr.shape
    if k > 64:
        return None
    nbytes = _lib.query("chr_moran_plan_bytes", n, k)
    buf = torch.empty(nbytes + 256, dtype=torch.uint8, device=nbr.device)
    pad = (-buf.data_ptr()) % 256
    buf = buf[pad:pad + nbytes]
    ws_bytes = _lib.query("chr_moran_plan_workspace_bytes", n, k)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=nbr.device)
    info = (ctypes.c_int32 * 4)()
    _lib.call("chr_moran_plan_build", _ptr(nbr), n, k, _ptr(buf), nbytes, ctypes.addressof(info), _ptr(ws), ws_bytes, _stream())
    # keep only the used part of the (worst-case sized) buffer
    used = 256 + ((n + 3) // 4 + 1 + 64) * 4 + 256 + int(info[1]) * 16 + 256
    keep = buf[:min(nbytes, used)].clone() if used < nbytes // 2 else buf
    if keep.data_ptr() % 256:
        keep = buf
    return MoranPlan(keep, int(info[0]), int(info[1]), n, k)


def moran_dense(X: torch.Tensor, nbr: torch.Tensor, n_genes: int | None = None, geary: bool = False,
                timing: bool = False, want_stats: bool = False, variant: int | None = None, preshifted: bool = False,
                plan: MoranPlan | None = None, row_map: torch.Tensor | None = None):
    """Moran's I (and Geary's C) of every column of the dense [n, ld] float32 matrix ``X``.

    Returns ``(I, C, stats)`` float64 device tensors (``C``/``stats`` None unless requested).
    ``plan``: the quad plan of ``nbr`` (built here when not given; reuse it across calls with the same W).
    ``preshifted``: the columns of ``X`` already have (near-)zero mean, skip the in-kernel shift.
    ``row_map``: int32 [n]: spot i takes row ``row_map[i]`` of ``X`` (one permutation draw, see ``moran_permutation_test``).
    """
    assert X.dtype == torch.float32 and X.dim() == 2 and X.stride(1) == 1
    assert row_map is None or (row_map.dtype == torch.int32 and row_map.is_contiguous() and row_map.numel() == X.shape[0])
    assert nbr.dtype == torch.int32 and nbr.is_contiguous()
    n, k = nbr.shape
    assert X.shape[0] == n
    ld = X.stride(0)
    g = X.shape[1] if n_genes is None else n_genes
    if variant is None:
        variant = int(os.environ.get("CHR_MORAN_VARIANT", "0"))   # kernel-variant override for tuning runs
    if plan is None and variant != 4:
        plan = moran_plan(nbr)
    flags = ((_lib.FLAG_GEARY if geary else 0) | (_lib.FLAG_TIMING if timing else 0) |
             (_lib.FLAG_PRESHIFTED if preshifted else 0) | (variant << _lib.VARIANT_SHIFT))
    ws_bytes = _lib.query("chr_moran_workspace_bytes", n, g, k, flags)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=X.device)
    out_I = torch.empty(g, dtype=torch.float64, device=X.device)
    out_C = torch.empty(g, dtype=torch.float64, device=X.device) if geary else None
    stats = torch.empty((4, g), dtype=torch.float64, device=X.device) if want_stats else None
    _lib.call("chr_moran_dense_f32", _ptr(X), n, g, ld, _ptr(nbr), k, _ptr(plan.buf) if plan is not None else 0,
              plan.chunk_units if plan is not None else 0, _ptr(row_map), flags, _ptr(out_I), _ptr(out_C), _ptr(stats),
              _ptr(ws), ws_bytes, _stream())
    return out_I, out_C, stats


# ---- K3/K4 with a block-diagonal W (several samples, one launch) ---------------------------------------------------

BLOCK_ALIGN_ROWS = 448      # CHR_BLOCK_ALIGN_ROWS (include/chrysalis_b200.h): blocks start on whole CTA row ranges


def block_layout(sizes):
    """Row blocks of a multi-sample matrix: block b holds ``sizes[b]`` spots followed by padding up to the next multiple of
    ``BLOCK_ALIGN_ROWS``.  Returns (block_rows int64 [B + 1], block_n int64 [B], pad_rows int32 - all host arrays)."""
    sizes = np.asarray(list(sizes), dtype=np.int64)
    padded = (sizes + BLOCK_ALIGN_ROWS - 1) // BLOCK_ALIGN_ROWS * BLOCK_ALIGN_ROWS
    block_rows = np.concatenate([[0], np.cumsum(padded)]).astype(np.int64)
    pad = [np.arange(block_rows[b] + sizes[b], block_rows[b + 1]) for b in range(len(sizes))]
    return block_rows, sizes, (np.concatenate(pad) if pad else np.empty(0)).astype(np.int32)


def moran_blockdiag(X: torch.Tensor, plan: MoranPlan, block_rows: np.ndarray, block_n: np.ndarray, pad_rows: torch.Tensor,
                    n_genes: int | None = None, geary: bool = False, timing: bool = False):
    """Moran's I (and Geary's C) per (block, gene) of the dense [n_rows, ld] float32 matrix ``X`` whose rows are laid out
    by ``block_layout`` and whose W (``plan``) is block diagonal.  Returns (I [B, g], C [B, g] or None) float64 device
    tensors.  The padding rows of ``X`` are overwritten."""
    assert X.dtype == torch.float32 and X.dim() == 2 and X.stride(1) == 1 and plan is not None
    n_rows, ld = X.shape[0], X.stride(0)
    g = X.shape[1] if n_genes is None else n_genes
    B = len(block_n)
    block_rows = np.ascontiguousarray(block_rows, dtype=np.int64)
    block_n = np.ascontiguousarray(block_n, dtype=np.int64)
    assert block_rows.shape == (B + 1,) and int(block_rows[-1]) == n_rows == plan.n
    flags = (_lib.FLAG_GEARY if geary else 0) | (_lib.FLAG_TIMING if timing else 0)
    ws_bytes = _lib.query("chr_moran_blockdiag_workspace_bytes", n_rows, g, B)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=X.device)
    out_I = torch.empty((B, g), dtype=torch.float64, device=X.device)
    out_C = torch.empty((B, g), dtype=torch.float64, device=X.device) if geary else None
    _lib.call("chr_moran_blockdiag_f32", _ptr(X), n_rows, g, ld, plan.k, _ptr(plan.buf), plan.chunk_units, block_rows.ctypes.data,
              block_n.ctypes.data, B, _ptr(pad_rows) if pad_rows.numel() else 0, pad_rows.numel(), flags, _ptr(out_I), _ptr(out_C),
              _ptr(ws), ws_bytes, _stream())
    return out_I, out_C


# ---- K3/K4 on a square lattice: layout + corrections for the stencil kernel ----------------------------------------

LATTICE_BAND_ROWS = 60      # CHR_LATTICE_BAND_ROWS (include/chrysalis_b200.h)
LATTICE_CHUNK_COLS = 256    # longest run of lattice columns one CTA walks


@dataclass
class LatticePlan:
    """Square-lattice representation of W for ``chr_moran_lattice_f32`` (see csrc/moran_lattice.cu).

    ``row_of_spot`` [n] int64 maps a spot to its row of the dense matrix in lattice layout (``n_rows`` rows: the stored
    range of every band, absent nodes included as filler rows)."""
    row_of_spot: torch.Tensor
    n_rows: int
    n: int
    k: int
    bands: torch.Tensor        # int32 [n_bands, 4]
    chunks: torch.Tensor       # int32 [n_chunks, 4]
    corr: torch.Tensor         # int32 [n_corr, 4]
    absent: torch.Tensor       # int32 [n_absent]
    pilot_rows: torch.Tensor   # int32 [<= 512]
    lattice_shape: tuple       # (W, H) nodes
    last_band_rows: int = 0    # lattice rows of the last band (the others have LATTICE_BAND_ROWS)
    _maps: tuple | None = None

    def maps(self):
        """(spot_of_row int32 [n_rows] with -1 for filler rows, row_of_spot int32 [n]) - what the permutation kernel needs."""
        if self._maps is None:
            ros = self.row_of_spot.to(torch.int32).contiguous()
            sor = torch.full((self.n_rows,), -1, dtype=torch.int32, device=ros.device)
            sor[self.row_of_spot] = torch.arange(self.n, dtype=torch.int32, device=ros.device)
            self._maps = (sor, ros)
        return self._maps


def lattice_nodes(xy: torch.Tensor, bbox: torch.Tensor, pitch: float):
    """Integer lattice node (ix, iy) of every spot, counted from the bounding-box corner."""
    rel = (xy - bbox[:2]) / pitch
    node = rel.round().to(torch.int64)
    return node[:, 0].contiguous(), node[:, 1].contiguous()


def lattice_layout(ix: torch.Tensor, iy: torch.Tensor, band_rows: int = LATTICE_BAND_ROWS, chunk_cols: int = LATTICE_CHUNK_COLS):
    """Bands / chunks / row map of the lattice layout for spots on nodes (ix, iy) (unique).  Device-agnostic tensor
    arithmetic; the per-band tables are tiny and assembled on the host (one round trip)."""
    n = ix.numel()
    W, H = (int(v) + 1 for v in torch.stack([ix.max(), iy.max()]).cpu())
    n_bands = (H + band_rows - 1) // band_rows
    band = iy // band_rows
    # occupancy of the node grid (padded to whole bands): two spots on one node, and per band the first / last occupied column.
    # (Dense reductions over the grid: a scatter-min / max of 700k spots into a dozen band slots serialises on the atomics.)
    occ = torch.bincount(iy * W + ix, minlength=n_bands * band_rows * W)
    crowded = (occ > 1).any().to(torch.int64)
    col_used = (occ.view(n_bands, band_rows, W) > 0).any(1)                                        # [n_bands, W]
    first = col_used.to(torch.int8).argmax(1)
    last = W - 1 - col_used.flip(1).to(torch.int8).argmax(1)
    used = col_used.any(1)
    xmin = torch.where(used, first, torch.full_like(first, torch.iinfo(torch.int64).max))
    xmax = torch.where(used, last, torch.full_like(last, -1))
    packed = torch.cat([xmin, xmax, crowded[None]]).cpu().numpy()                                  # one round trip
    if packed[-1]:
        return None, 0, None, None, (W, H)
    xmin_h, xmax_h = packed[:n_bands], packed[n_bands:2 * n_bands]
    bands_h = np.zeros((n_bands, 4), dtype=np.int64)
    chunks = []
    row = 0
    for b in range(n_bands):
        bh = min(band_rows, H - b * band_rows)
        ncols = int(xmax_h[b] - xmin_h[b] + 1) if xmax_h[b] >= 0 else 0
        bands_h[b] = (row, xmin_h[b] if ncols else 0, ncols, bh)
        row += ncols * bh
        pieces = (ncols + chunk_cols - 1) // chunk_cols
        size = (ncols + pieces - 1) // pieces if pieces else 0
        for c0 in range(0, ncols, size if size else 1):
            chunks.append((b, c0, min(size, ncols - c0), 0))
    bands_d = torch.from_numpy(bands_h).to(ix.device)
    row_of_spot = bands_d[band, 0] + (ix - bands_d[band, 1]) * bands_d[band, 3] + (iy - band * band_rows)
    return row_of_spot, int(row), bands_d.to(torch.int32).contiguous(), \
        torch.tensor(chunks, dtype=torch.int32, device=ix.device).reshape(-1, 4).contiguous(), (W, H)


def _lattice_corrections_general(ix, iy, nbr, shape):
    """D = A - B by coalescing ALL ordered pairs (any list: duplicates, self loops): exact integer coefficients."""
    n, k = nbr.shape
    W, H = shape
    dev = ix.device
    grid = torch.full(((H + 2) * (W + 2),), -1, dtype=torch.int64, device=dev)
    grid[(iy + 1) * (W + 2) + ix + 1] = torch.arange(n, device=dev)
    me = torch.arange(n, device=dev)
    keys, vals = [me[:, None].expand(n, k).reshape(-1) * n + nbr.reshape(-1).long()], [torch.ones(n * k, dtype=torch.int64, device=dev)]
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dx == 0 and dy == 0:
                continue
            j = grid[(iy + 1 + dy) * (W + 2) + ix + 1 + dx]
            ok = j >= 0
            keys.append(me[ok] * n + j[ok])
            vals.append(torch.full((keys[-1].numel(),), -1, dtype=torch.int64, device=dev))
    key, inv = torch.unique(torch.cat(keys), return_inverse=True)
    coef = torch.zeros(key.numel(), dtype=torch.int64, device=dev).scatter_add_(0, inv, torch.cat(vals))
    nz = coef != 0
    key, coef = key[nz], coef[nz]
    return key // n, key % n, coef


def lattice_corrections(ix: torch.Tensor, iy: torch.Tensor, nbr: torch.Tensor, shape: tuple, simple_lists: bool = False):
    """D = A - B as a COO list (i, j, coefficient) plus the (indeg - k) weights of W's column sums.

    A: kNN adjacency counts of ``nbr`` [n, k] (spot ids); B: 1 for every ordered pair of present nodes that are
    8-neighbours on the lattice.  kNN lists (no repeats, no self) only differ from the stencil at tissue edges and holes:
    list entries that are not stencil neighbours (+1) and stencil neighbours the list misses (-1) are found with dense
    [n, k] / [n, 8] masks and two host round trips; lists with repeats or self loops take the general coalescing path."""
    n, k = nbr.shape
    W, H = shape
    dev = ix.device
    me = torch.arange(n, device=dev)
    nb = nbr.long()
    grid = torch.full(((H + 2) * (W + 2),), -1, dtype=torch.int64, device=dev)
    grid[(iy + 1) * (W + 2) + ix + 1] = me
    offs = torch.tensor([dy * (W + 2) + dx for dy in (-1, 0, 1) for dx in (-1, 0, 1) if dx or dy], device=dev)
    J = grid[((iy + 1) * (W + 2) + ix + 1)[:, None] + offs[None, :]]                          # [n, 8] stencil neighbours, -1 = absent
    extra = ~(((ix[nb] - ix[:, None]).abs() <= 1) & ((iy[nb] - iy[:, None]).abs() <= 1)) | (nb == me[:, None])    # [n, k]
    missing = (J >= 0) & ~(nb[:, None, :] == J[:, :, None]).any(2)                            # [n, 8]
    if simple_lists:                                            # lists of the kNN builder: k distinct neighbours, never the spot itself
        irregular = torch.zeros((), dtype=torch.bool, device=dev)
    else:
        srt = nb.sort(dim=1).values
        irregular = (srt[:, 1:] == srt[:, :-1]).any() | (nb == me[:, None]).any()
    indeg = torch.bincount(nb.reshape(-1), minlength=n)
    wmask = indeg != k
    counts = torch.stack([extra.sum(), missing.sum(), wmask.sum(), irregular.to(torch.int64)]).cpu()   # round trip 1
    n_extra, n_missing, n_w, irr = (int(v) for v in counts)
    hits = torch.nonzero(torch.cat([extra.reshape(-1), missing.reshape(-1), wmask]))[:, 0]               # round trip 2 (sizes known)
    e, m, w = hits[:n_extra], hits[n_extra:n_extra + n_missing] - n * k, hits[n_extra + n_missing:] - n * (k + 8)
    wj = w
    if irr:
        ci, cj, cc = _lattice_corrections_general(ix, iy, nbr, shape)
    else:
        ei, mi = e // k, m // 8
        ci = torch.cat([ei, mi])
        cj = torch.cat([nb.reshape(-1)[e], J.reshape(-1)[m]])
        cc = torch.cat([torch.ones_like(ei), -torch.ones_like(mi)])
    return ci, cj, cc, wj, (indeg[wj] - k)


def lattice_plan(xy: torch.Tensor, nbr: torch.Tensor, bbox: torch.Tensor | None = None, pitch: float | None = None,
                 max_fill: float = 1.6, max_corr_frac: float = 0.2, simple_lists: bool = False) -> LatticePlan | None:
    """Lattice layout + corrections for the kNN lists ``nbr`` [n, 8] of spots on a square lattice, or None when the
    stencil kernel does not apply (k != 8, no lattice, two spots on one node, a tissue so ragged that the stored
    range or the correction list would dominate): the caller then takes the general quad-plan kernel."""
    n, k = nbr.shape
    if k != 8 or n < 64:
        return None
    bbox = spatial_bbox(xy) if bbox is None else bbox
    if pitch is None:                                                   # the lists are sorted by distance: column 0 is the nearest neighbour
        pitch = lattice_pitch_from_nn(xy, bbox, (xy[nbr[:, 0].long()] - xy).norm(dim=1))
    if pitch is None:
        return None
    ix, iy = lattice_nodes(xy, bbox, pitch)
    row_of_spot, n_rows, bands, chunks, shape = lattice_layout(ix, iy)
    if row_of_spot is None:                                             # two spots on one node
        return None
    if n_rows > max_fill * n or n_rows >= 2 ** 31 or chunks.shape[0] >= 65536:
        return None
    ci, cj, cc, wj, ww = lattice_corrections(ix, iy, nbr, shape, simple_lists)
    if ci.numel() + wj.numel() > max_corr_frac * n * k:
        return None
    f32bits = lambda t: t.to(torch.float32).view(torch.int32)
    quad = torch.stack([row_of_spot[ci].to(torch.int32), row_of_spot[cj].to(torch.int32), f32bits(cc), torch.zeros_like(ci, dtype=torch.int32)], 1)
    lin = torch.stack([row_of_spot[wj].to(torch.int32), torch.zeros_like(wj, dtype=torch.int32), f32bits(ww), torch.ones_like(wj, dtype=torch.int32)], 1)
    corr = torch.cat([quad, lin], 0)
    corr = corr[torch.argsort(corr[:, 0].long() * 2 + corr[:, 3].long(), stable=True)].contiguous()    # by row: neighbouring entries share rows
    present = torch.zeros(n_rows, dtype=torch.bool, device=xy.device)
    present[row_of_spot] = True
    absent = torch.nonzero(~present).squeeze(1).to(torch.int32).contiguous()
    ns = min(n, 512)
    pilot_rows = row_of_spot[(torch.arange(ns, device=xy.device) * n) // ns].to(torch.int32).contiguous()
    last_rows = shape[1] - (bands.shape[0] - 1) * LATTICE_BAND_ROWS
    return LatticePlan(row_of_spot, n_rows, n, k, bands, chunks, corr, absent, pilot_rows, shape, int(last_rows))


def moran_lattice(X: torch.Tensor, plan: LatticePlan, n_genes: int | None = None, geary: bool = False, timing: bool = False,
                  want_stats: bool = False, preshifted: bool = False, variant: int | None = None):
    """Moran's I (and Geary's C) of every column of ``X`` [plan.n_rows, ld] float32 in lattice layout.
    Returns ``(I, C, stats)`` like ``moran_dense``.  The filler rows of ``X`` (absent nodes) are overwritten."""
    assert X.dtype == torch.float32 and X.dim() == 2 and X.stride(1) == 1 and X.shape[0] == plan.n_rows
    ld = X.stride(0)
    g = X.shape[1] if n_genes is None else n_genes
    if variant is None:                                              # 0: TMA producer (default), 2: cp.async producer warps
        variant = int(os.environ.get("CHR_LATTICE_VARIANT", "0"))
    flags = ((_lib.FLAG_GEARY if geary else 0) | (_lib.FLAG_TIMING if timing else 0) | (_lib.FLAG_PRESHIFTED if preshifted else 0) |
             (variant << _lib.VARIANT_SHIFT))
    n_chunks, n_corr = plan.chunks.shape[0], plan.corr.shape[0]
    ws_bytes = _lib.query("chr_moran_lattice_workspace_bytes", g, n_chunks, n_corr)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=X.device)
    out_I = torch.empty(g, dtype=torch.float64, device=X.device)
    out_C = torch.empty(g, dtype=torch.float64, device=X.device) if geary else None
    stats = torch.empty((4, g), dtype=torch.float64, device=X.device) if want_stats else None
    _lib.call("chr_moran_lattice_f32", _ptr(X), plan.n_rows, g, ld, plan.n, plan.k, _ptr(plan.bands), plan.bands.shape[0], plan.last_band_rows,
              _ptr(plan.chunks), n_chunks, _ptr(plan.corr) if n_corr else 0, n_corr, _ptr(plan.absent) if plan.absent.numel() else 0,
              plan.absent.numel(), _ptr(plan.pilot_rows), plan.pilot_rows.numel(), flags, _ptr(out_I), _ptr(out_C), _ptr(stats),
              _ptr(ws), ws_bytes, _stream())
    return out_I, out_C, stats


def moran_lattice_csr(csr: DeviceCSR, gene_col: torch.Tensor, n_keep: int, scale: torch.Tensor | None, plan: LatticePlan,
                      apply_log1p: bool, geary: bool = False, timing: bool = False, want_stats: bool = False):
    """Moran's I (and Geary's C) of the kept genes straight from the device CSR counts (``chr_moran_lattice_csr_f32``): the
    column map of the gene filter, the per-spot scale, log1p and the densification are fused into the stencil kernel's
    producers; no dense matrix is built.  ``csr`` in spot order (the order ``plan`` was built for), columns sorted inside
    every row.  Returns ``(I, C, stats, unsorted)``: ``unsorted`` is a device int32 that is 1 when a row was not sorted -
    the caller must then discard the results and take the dense path."""
    n, g_all = csr.shape
    assert n == plan.n and gene_col.dtype == torch.int32 and gene_col.numel() == g_all
    nnz = csr.indices.numel()
    n_chunks, n_corr = plan.chunks.shape[0], plan.corr.shape[0]
    sor, _ = plan.maps()
    flags = (_lib.FLAG_GEARY if geary else 0) | (_lib.FLAG_TIMING if timing else 0)
    ws_bytes = _lib.query("chr_moran_lattice_csr_workspace_bytes", nnz, plan.n_rows, n_keep, n_chunks, n_corr)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=csr.data.device)
    out_I = torch.empty(n_keep, dtype=torch.float64, device=csr.data.device)
    out_C = torch.empty(n_keep, dtype=torch.float64, device=csr.data.device) if geary else None
    stats = torch.empty((4, n_keep), dtype=torch.float64, device=csr.data.device) if want_stats else None
    unsorted = torch.zeros(1, dtype=torch.int32, device=csr.data.device)
    _lib.call("chr_moran_lattice_csr_f32", _ptr(csr.indptr), _ptr(csr.indices), _ptr(csr.data), nnz, n, g_all, _ptr(gene_col), n_keep,
              _ptr(scale), int(apply_log1p), plan.n_rows, plan.k, _ptr(plan.bands), plan.bands.shape[0], _ptr(plan.chunks), n_chunks,
              _ptr(plan.corr) if n_corr else 0, n_corr, _ptr(sor), _ptr(plan.pilot_rows), plan.pilot_rows.numel(), flags, _ptr(out_I),
              _ptr(out_C), _ptr(stats), _ptr(unsorted), _ptr(ws), ws_bytes, _stream())
    return out_I, out_C, stats, unsorted


def perm_seed(seed: int | None) -> int:
    """64-bit key of the device's counter-based permutations (``csrc/perm.cuh``); ``None`` draws fresh entropy.  torch's and
    numpy's global RNG states are never touched."""
    return int(np.random.SeedSequence(seed).generate_state(1, np.uint64)[0])


def perm_row_map(n: int, seed: int, draw: int, device) -> torch.Tensor:
    """pi_draw as an int32 row map (spot i takes the row of spot pi(i))."""
    out = torch.empty(n, dtype=torch.int32, device=device)
    _lib.call("chr_perm_row_map", n, seed & (2 ** 64 - 1), int(draw), _ptr(out), _stream())
    return out


def _perm_summary(I, larger, s1, s2, permutations: int):
    """esda's simulation summaries from #{I_sim >= I}, sum I_sim, sum I_sim^2 (float64 device tensors over the genes)."""
    P = float(permutations)
    larger = torch.minimum(larger, int(permutations) - larger)                 # esda: two-sided fold
    ei = s1 / P
    vi = (s2 / P - ei * ei).clamp_min(0.0)
    se = vi.sqrt()
    z = (I - ei) / se
    p_z = 0.5 * torch.erfc(z.abs() / 2.0 ** 0.5)                               # 1 - Phi(|z|)
    return {"I": I, "p_sim": (larger.to(torch.float64) + 1.0) / (P + 1.0), "EI_sim": ei, "VI_sim": vi, "seI_sim": se,
            "z_sim": z, "p_z_sim": p_z}


def moran_permutation_test(X: torch.Tensor, nbr: torch.Tensor, n_genes: int | None = None, permutations: int = 999,
                           seed: int | None = None, plan: MoranPlan | None = None):
    """Permutation inference of ``esda.moran.Moran(y, w, permutations=P)`` for every column of ``X`` (general W).

    Each draw permutes the spots (rows) against a fixed W and is ONE pass of the Moran kernel over the
    whole matrix (``row_map`` = the device's counter-based permutation of that draw), so a draw is shared by all
    genes; esda draws per gene from NumPy's global RNG, so parity of the simulated quantities is statistical, parity
    of ``I`` exact.  Returns a dict of float64 device tensors [g]: I, p_sim, EI_sim, VI_sim, seI_sim, z_sim, p_z_sim.
    """
    n = X.shape[0]
    g = X.shape[1] if n_genes is None else n_genes
    if plan is None:
        plan = moran_plan(nbr)
    if plan is None:
        raise _lib.ChrysalisNativeError("permutation inference needs the plan-based kernel (k <= 64)")
    I, _, _ = moran_dense(X, nbr, g, plan=plan)
    key = perm_seed(seed)
    larger = torch.zeros(g, dtype=torch.int64, device=X.device)
    s1 = torch.zeros(g, dtype=torch.float64, device=X.device)
    s2 = torch.zeros(g, dtype=torch.float64, device=X.device)
    for draw in range(int(permutations)):
        Ip, _, _ = moran_dense(X, nbr, g, plan=plan, row_map=perm_row_map(n, key, draw, X.device))
        larger += (Ip >= I)
        s1 += Ip
        s2 += Ip * Ip
    return _perm_summary(I, larger, s1, s2, permutations)


def moran_perm_lattice(X: torch.Tensor, plan: LatticePlan, n_genes: int | None = None, permutations: int = 999,
                       seed: int | None = None, timing: bool = False, variant: int | None = None):
    """The same inference on a square lattice (``chr_moran_perm_lattice_f32``): all P draws inside the library, batches of
    draws per launch, the gene slab reused from L2 across the draws of a batch.  ``X`` in lattice layout.
    ``variant`` (env CHR_PERM_VARIANT): 0 = one bulk copy per gathered row (default), 2 = 16-byte cp.async producers."""
    assert X.dtype == torch.float32 and X.dim() == 2 and X.stride(1) == 1 and X.shape[0] == plan.n_rows
    ld = X.stride(0)
    g = X.shape[1] if n_genes is None else n_genes
    n_chunks, n_corr = plan.chunks.shape[0], plan.corr.shape[0]
    sor, ros = plan.maps()
    if variant is None:
        variant = int(os.environ.get("CHR_PERM_VARIANT", "0"))
    flags = (_lib.FLAG_TIMING if timing else 0) | (variant << _lib.VARIANT_SHIFT)
    ws_bytes = _lib.query("chr_moran_perm_lattice_workspace_bytes", plan.n_rows, g, n_chunks, n_corr, int(permutations))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=X.device)
    I = torch.empty(g, dtype=torch.float64, device=X.device)
    larger = torch.empty(g, dtype=torch.int64, device=X.device)
    s1 = torch.empty(g, dtype=torch.float64, device=X.device)
    s2 = torch.empty(g, dtype=torch.float64, device=X.device)
    _lib.call("chr_moran_perm_lattice_f32", _ptr(X), plan.n_rows, g, ld, plan.n, plan.k, _ptr(plan.bands), plan.bands.shape[0],
              _ptr(plan.chunks), n_chunks, _ptr(plan.corr) if n_corr else 0, n_corr, _ptr(plan.absent) if plan.absent.numel() else 0,
              plan.absent.numel(), _ptr(plan.pilot_rows), plan.pilot_rows.numel(), _ptr(sor), _ptr(ros), int(permutations),
              perm_seed(seed) & (2 ** 64 - 1), flags, _ptr(I), _ptr(larger), _ptr(s1), _ptr(s2), _ptr(ws), ws_bytes,
              _stream())
    return _perm_summary(I, larger, s1, s2, permutations)


def local_moran(y: torch.Tensor, nbr: torch.Tensor, permutations: int = 999, seed: int = 0):
    """``esda.moran.Moran_Local(y, w, permutations=P)`` for one variable ``y`` [n] (float32, device, spot order of ``nbr``).
    Returns a dict of device tensors [n]: Is, lag, q, and with permutations p_sim, EI_sim, VI_sim."""
    assert y.dim() == 1 and nbr.dtype == torch.int32 and nbr.is_contiguous() and nbr.shape[0] == y.numel()
    n, k = nbr.shape
    yd = y.to(torch.float64)
    z = ((yd - yd.mean()) / yd.std(unbiased=False)).to(torch.float32).contiguous()
    z2ss = float((z.to(torch.float64) ** 2).sum().item())
    f = lambda dt: torch.empty(n, dtype=dt, device=y.device)
    lag, Is, q = f(torch.float32), f(torch.float32), f(torch.int32)
    p_sim, ei, vi = (f(torch.float32), f(torch.float32), f(torch.float32)) if permutations else (None, None, None)
    _lib.call("chr_local_moran_f32", _ptr(z), n, _ptr(nbr), k, z2ss, int(permutations), int(seed) & (2 ** 64 - 1), _ptr(lag), _ptr(Is),
              _ptr(q), _ptr(p_sim), _ptr(ei), _ptr(vi), _stream())
    out = {"Is": Is, "lag": lag, "q": q, "z": z}
    if permutations:
        out.update({"p_sim": p_sim, "EI_sim": ei, "VI_sim": vi})
    return out


def last_kernel_ms(family: int = _lib.FAMILY_MORAN) -> float:
    return float(_lib.query("chr_last_kernel_ms", family))


# ---- K6 - K8: PCA ------------------------------------------------------------------------

def _workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(int(nbytes), dtype=torch.uint8, device=device)


def column_sums(P: torch.Tensor, s: int) -> torch.Tensor:
    """float64 column sums of the first ``s`` columns of the dense [n, ld] float32 matrix."""
    n, ld = P.shape[0], P.stride(0)
    ws_bytes = _lib.query("chr_colsum_workspace_bytes", n, s)
    ws = _workspace(ws_bytes, P.device)
    out = torch.empty(s, dtype=torch.float64, device=P.device)
    _lib.call("chr_colsum_f64", _ptr(P), n, s, ld, _ptr(out), _ptr(ws), ws_bytes, _stream())
    return out


def gram_centered(P: torch.Tensor, s: int, mean: torch.Tensor, timing: bool = False, variant: int | None = None) -> torch.Tensor:
    """(P - mean)^T (P - mean) as float64 [s, s] (tcgen05, split-bf16 operands).  ``variant`` (env CHR_GRAM_VARIANT): 0 = 256 x 256
    super-tile kernel (default), 2 = 128 x 128 tile kernel; both produce the same bits."""
    n, ld = P.shape[0], P.stride(0)
    if variant is None:
        variant = int(os.environ.get("CHR_GRAM_VARIANT", "0"))
    assert mean.dtype == torch.float64 and mean.numel() == s
    ws_bytes = _lib.query("chr_gram_workspace_bytes", n, s)
    ws = _workspace(ws_bytes, P.device)
    C = torch.empty((s, s), dtype=torch.float64, device=P.device)
    _lib.call("chr_gram_centered_bf16x3", _ptr(P), n, s, ld, _ptr(mean), _ptr(C),
              (_lib.FLAG_TIMING if timing else 0) | (variant << _lib.VARIANT_SHIFT), _ptr(ws), ws_bytes, _stream())
    return C


def eigh_topk(C: torch.Tensor, r: int, rtol: float = 1e-9):
    """r largest eigenpairs of the symmetric float64 matrix C: (evals [r] desc, evecs [r, s]).

    The Lanczos depth is doubled until every Ritz residual is below ``rtol * lambda_max``
    (or the Krylov space is the whole space)."""
    s = C.shape[0]
    assert C.dtype == torch.float64 and C.is_contiguous() and C.shape[1] == s
    m = _lib.query("chr_eigh_default_steps", s, r)
    while True:
        ws_bytes = _lib.query("chr_eigh_workspace_bytes", s, r, m)
        ws = _workspace(ws_bytes, C.device)
        evals = torch.empty(r, dtype=torch.float64, device=C.device)
        evecs = torch.empty((r, s), dtype=torch.float64, device=C.device)
        res = torch.empty(r, dtype=torch.float64, device=C.device)
        _lib.call("chr_eigh_topk_f64", _ptr(C), s, r, m, _ptr(evals), _ptr(evecs), _ptr(res), _ptr(ws), ws_bytes, _stream())
        worst = float((res / evals[0].abs().clamp_min(1e-300)).max().item())
        if worst <= rtol or m >= min(s, 1024):
            order = torch.argsort(evals, descending=True)
            return evals[order], evecs[order], worst
        m = min(s, 1024, 2 * m)


def pca_scores(P: torch.Tensor, s: int, mean: torch.Tensor, V: torch.Tensor) -> torch.Tensor:
    """(P - mean) V^T as float32 [n, r];  V: [r, s] float64."""
    n, ld = P.shape[0], P.stride(0)
    r = V.shape[0]
    ws_bytes = _lib.query("chr_pca_scores_workspace_bytes", s, r)
    ws = _workspace(ws_bytes, P.device)
    Y = torch.empty((n, r), dtype=torch.float32, device=P.device)
    _lib.call("chr_pca_scores_f32", _ptr(P), n, s, ld, _ptr(mean), _ptr(V.contiguous()), r, _ptr(Y), r, _ptr(ws), ws_bytes,
              _stream())
    return Y


# ---- K9 - K11: archetypal analysis ---------------------------------------------------------

AA_PENALTY_M = 200.0   # weight of the sum-to-one row in the penalised NNLS problems (archetypes 0.4.x)


def aa_furthest_sum(X: torch.Tensor, n_archetypes: int, first: int) -> np.ndarray:
    """FurthestSum seeding on the device: indices (int64 [n_archetypes]) of the initial archetypes."""
    import ctypes

    assert X.dtype == torch.float64 and X.is_contiguous() and X.dim() == 2
    n, d = X.shape
    ws_bytes = _lib.query("chr_aa_furthest_sum_workspace_bytes", n)
    ws = _workspace(ws_bytes, X.device)
    rows = (ctypes.c_int64 * n_archetypes)()
    _lib.call("chr_aa_furthest_sum", _ptr(X), n, d, n_archetypes, int(first), ctypes.addressof(rows), _ptr(ws), ws_bytes,
              _stream())
    return np.array(list(rows), dtype=np.int64)


class AAState:
    """Device buffers of one AA restart: data X [n, d] float64, archetypes Z, coefficients alphas."""

    def __init__(self, X: torch.Tensor, n_archetypes: int):
        assert X.dtype == torch.float64 and X.is_contiguous() and X.dim() == 2
        self.X, self.na = X, int(n_archetypes)
        n, d = X.shape
        self.Z = torch.empty((self.na, d), dtype=torch.float64, device=X.device)
        self.alphas = torch.empty((n, self.na), dtype=torch.float64, device=X.device)
        self.rss = torch.empty(1, dtype=torch.float64, device=X.device)
        self.ws_bytes = _lib.query("chr_aa_workspace_bytes", n, d, self.na)
        self.ws = _workspace(self.ws_bytes, X.device)
        self.beta_passes = 0
        self.n_done = 0          # iterations since the last (re-)initialisation: > 0 warm-starts the active sets

    def init_from_rows(self, rows):
        """Z = X[rows]: one-hot betas0 (archetypes 0.4.x initialises archetypes on data points)."""
        n, d = self.X.shape
        r = torch.as_tensor(np.asarray(rows, dtype=np.int64)).to(self.X.device)
        _lib.call("chr_aa_init_archetypes", _ptr(self.X), n, d, _ptr(r), self.na, _ptr(self.Z), _stream())
        self.n_done = 0

    def iterate(self) -> float:
        """One alternating-NNLS iteration in place; returns the residual sum of squares."""
        import ctypes

        n, d = self.X.shape
        passes = ctypes.c_int(0)
        _lib.call("chr_aa_iterate", _ptr(self.X), n, d, self.na, AA_PENALTY_M, int(self.n_done > 0), _ptr(self.Z),
                  _ptr(self.alphas), _ptr(self.rss), ctypes.addressof(passes), _ptr(self.ws), self.ws_bytes, _stream())
        self.beta_passes = passes.value
        self.n_done += 1
        return float(self.rss.item())
