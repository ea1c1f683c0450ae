"""Golden vectors produced by RUNNING reference code (not the oracle).

``/root/reference/chrysalis/functions.py:15-35`` (``get_moransI``) is the only statement of the
Moran statistic that lives in the reference repository itself and it needs numpy only.  The module
around it cannot be imported (scanpy / archetypes / pysal at functions.py:4-9 are not installed),
so the function's source is cut out of the file with ``ast`` and executed as it stands - no line
of it is re-typed here or stored in this repository.  Its outputs on the committed fixtures
(dense row-standardised kNN W of the fixture, one normalised log1p gene column at a time) are
written to ``tests/golden/reference_get_moransI.npz``; ``tests/test_oracle.py`` pins
``oracle.moran.morans_I`` (and through it every GPU parity test) to these reference-run numbers.

    python tests/golden/make_reference_vectors.py      # needs /root/reference (build container only)
"""
import ast
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
REF_FILE = "/root/reference/chrysalis/functions.py"


def load_reference_get_moransI(path: str = REF_FILE):
    """The reference's ``get_moransI`` as a callable, cut out of its module by ``ast`` (nothing else of the
    module is executed).  Returns (function, (first_line, last_line))."""
    src = open(path).read()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "get_moransI")
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {"np": np}
    exec(compile(mod, path, "exec"), ns)   # noqa: S102 - executing the reference is the point
    return ns["get_moransI"], (fn.lineno, fn.end_lineno)


def fixture_inputs(name: str):
    """(dense float64 W, dense float32 normalised log1p matrix of the scored genes, scored gene ids)."""
    import scipy.sparse as sp

    from oracle import svg as osvg
    from oracle import weights as ow

    z = np.load(os.path.join(HERE, name + ".npz"))
    X = sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=tuple(z["shape"]))
    mask = z["gene_mask"]
    dense = np.asarray(osvg.normalize_total_log1p(X[:, mask]).todense(), dtype=np.float32)
    # W exactly as the reference builds it: cKDTree kNN on (x, -y), row-standardised (core.py:61-65)
    nbr = ow.knn_neighbors(ow.reference_points(z["xy"]), int(z["k"]))
    w = np.asarray(ow.row_standardised_csr(nbr).todense())
    return w, dense, np.flatnonzero(mask)


def main():
    fn, lines = load_reference_get_moransI()
    out = {"source": np.array(f"{REF_FILE}:{lines[0]}-{lines[1]}")}
    for name in ("svg_hex_k6", "svg_random_k8", "svg_square_k8"):
        w, dense, genes = fixture_inputs(name)
        cols = np.arange(0, dense.shape[1], 3)                 # every third scored gene
        vals32 = np.array([fn(w, dense[:, j]) for j in cols])                       # float32 column, as ad.to_df() gives it
        vals64 = np.array([fn(w, dense[:, j].astype(np.float64)) for j in cols])
        out[f"{name}_genes"] = genes[cols]
        out[f"{name}_I_f32in"] = vals32.astype(np.float64)
        out[f"{name}_I_f64in"] = vals64
        print(name, "genes", len(cols), "I range", float(vals64.min()), float(vals64.max()))
    np.savez_compressed(os.path.join(HERE, "reference_get_moransI.npz"), **out)


if __name__ == "__main__":
    main()
