"""CPU oracle for the chrysalis hot path (detect_svgs -> pca -> aa).

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or as
the timed CPU baseline.  ``chrysalis_b200`` never imports this package.

What it restates (reference = /root/reference, rockdeme/chrysalis 0.2.0):

* ``chrysalis/core.py:12-92``  ``detect_svgs``  -> :mod:`oracle.svg`
* ``chrysalis/core.py:95-137`` ``pca``          -> :mod:`oracle.pca`
* ``chrysalis/core.py:140-208`` ``aa``          -> :mod:`oracle.aa`

The arithmetic of that path lives in third-party packages that are NOT in
the reference tree and NOT installed in this image (no network):
``libpysal`` / ``esda`` (via the unpinned meta-package ``pysal``,
``pyproject.toml:28``), ``scanpy`` (unpinned, ``pyproject.toml:29``) and
``archetypes==0.4.2`` (pinned, ``pyproject.toml:24``).  Their published
algorithms are restated here with numpy / scipy (``cKDTree`` *is* what
libpysal's KNN calls, so neighbour sets and tie order are reproduced
exactly); ``sklearn.decomposition.PCA`` is installed and is called directly.

PARITY PINNING STATUS
---------------------
* Moran's I: pinned against the only in-repo statement of the formula,
  ``chrysalis/functions.py:15-35`` (``get_moransI``, dense W, 8 dp), restated
  in :func:`oracle.moran.morans_I_dense_formula`, plus hand-computed lattices.
  The reference's own tests (``chrysalis/test/test.py``) contain no
  assertions, so there are no golden vectors to check against.
* SVG rule / preprocessing: restated from ``core.py:50-92`` and scanpy's
  documented semantics - **parity unpinned** (no reference fixtures exist).
* PCA: the reference call itself (sklearn, installed) - exact.
* AA: ``archetypes==0.4.2`` source is absent - **parity unpinned**; items the
  restatement could not verify offline are marked UNVERIFIED in oracle/aa.py.
"""
