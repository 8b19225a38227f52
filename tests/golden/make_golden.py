"""Generates tests/golden/*.json from the REAL reference (oracle/_ref/libhpccg_ref_*.so, i.e. the
unmodified /root/reference sources compiled by oracle/build.sh).  Run in the build container only:

    python tests/golden/make_golden.py [--big]

Fixtures hold, per configuration: every scalar field, SHA-256 of every array the reference builds
(matrix, halo lists, vectors), the residual history of HPCCG() at 17 significant digits for every
iteration, and per-kernel outputs for seeded inputs.  --big adds the 256^3 history (config #2 of
BASELINE.json; ~2 minutes serial, 12 GB).
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "oracle"))
import refwrap  # noqa: E402

ARRAYS = ["nnz_in_row", "list_of_inds", "list_of_vals", "ind_offsets", "val_offsets", "diag_offsets", "x", "b", "xexact",
          "external_index", "external_local_index", "elements_to_send", "neighbors", "recv_length", "send_length"]

CONFIGS = [  # (nx, ny, nz, ranks, stencil)
    (10, 10, 10, 1, 27), (20, 30, 10, 1, 27), (20, 30, 10, 1, 7), (16, 16, 16, 1, 27), (33, 17, 5, 1, 27), (1, 1, 1, 1, 27),
    (7, 1, 1, 1, 27), (4, 3, 2, 3, 27), (4, 3, 1, 4, 27), (5, 4, 3, 2, 7), (16, 16, 8, 8, 27), (16, 16, 8, 2, 27),
    (12, 10, 2, 3, 27), (32, 32, 16, 4, 27), (64, 64, 64, 1, 27), (64, 64, 32, 2, 27), (64, 64, 64, 1, 7),
]


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def seeded(n, seed):
    return np.random.default_rng(seed).uniform(-1.0, 1.0, n)


def config_record(nx, ny, nz, size, stencil, max_iter=150):
    with refwrap.RefWorld(nx, ny, nz, size=size, stencil=stencil) as R:
        rec = {"dims": [nx, ny, nz], "ranks": size, "stencil": stencil, "variant": R.variant, "rank": []}
        for r in range(size):
            d = {"scalars": {s: R.scalar(r, s) for s in refwrap.SCALARS}, "sha256": {}, "small": {}}
            for a in ARRAYS:
                arr = R.array(r, a)
                d["sha256"][a] = sha(arr)
                if a in ("external_index", "external_local_index", "elements_to_send", "neighbors", "recv_length",
                         "send_length") and arr.size <= 64:
                    d["small"][a] = arr.tolist()
            rec["rank"].append(d)
        # per-kernel: SpMV (with halo exchange) and ddot on seeded vectors
        xs = [np.concatenate([seeded(R.scalar(r, "local_nrow"), 12345 + r),
                              np.zeros(R.scalar(r, "local_ncol") - R.scalar(r, "local_nrow"))]) for r in range(size)]
        ys = R.spmv(xs, exchange=True)
        rec["spmv_sha256"] = [sha(y) for y in ys]
        rec["spmv_head"] = [[v.hex() for v in y[:4].tolist()] for y in ys]
        v2 = [seeded(R.scalar(r, "local_nrow"), 54321 + r) for r in range(size)]
        xloc = [x[:R.scalar(r, "local_nrow")].copy() for r, x in enumerate(xs)]
        rec["ddot_xy"] = float(R.ddot(xloc, v2)[0]).hex()
        rec["ddot_xx"] = float(R.ddot(xloc, xloc)[0]).hex()
        s = R.solve(max_iter, 0.0, hist=True)
        rec["max_iter"] = max_iter
        rec["niters"] = s["niters"]
        rec["normr"] = float(s["normr"]).hex()
        rec["hist"] = [float(v).hex() for v in s["hist"]]
        rec["x_max_err"] = float(max(np.abs(x - 1.0).max() for x in s["x"]))
    return rec


def main():
    out = {"generator": "tests/golden/make_golden.py", "source": "unmodified /root/reference via oracle/build.sh",
           "rng": "numpy.random.default_rng(seed).uniform(-1,1,n); seeds 12345+rank / 54321+rank", "configs": []}
    for cfg in CONFIGS:
        print("golden", cfg, flush=True)
        out["configs"].append(config_record(*cfg))
    # waxpby: the three branches on seeded vectors
    n = 1001
    x, y = seeded(n, 12345), seeded(n, 54321)
    wax = {}
    for name, (a, b) in {"alpha1": (1.0, -1.4142135623730951), "beta1": (0.7071067811865476, 1.0),
                         "general": (0.7071067811865476, -1.4142135623730951)}.items():
        wax[name] = {"alpha": a, "beta": b, "sha256": sha(refwrap.waxpby(a, x, b, y))}
    out["waxpby"] = {"n": n, "cases": wax}
    (ROOT / "tests" / "golden" / "golden.json").write_text(json.dumps(out, indent=1))
    if "--big" in sys.argv:
        print("golden 256^3 (serial reference, ~2 min)", flush=True)
        with refwrap.RefWorld(256, 256, 256) as R:
            s = R.solve(150, 0.0, hist=True, want_x=True)
            big = {"dims": [256, 256, 256], "ranks": 1, "stencil": 27, "variant": R.variant, "max_iter": 150,
                   "niters": s["niters"], "normr": float(s["normr"]).hex(), "hist": [float(v).hex() for v in s["hist"]],
                   "x_max_err": float(np.abs(s["x"][0] - 1.0).max()), "nnz_sum": R.scalar(0, "nnz_sum"),
                   "times": s["times"].tolist()}
        (ROOT / "tests" / "golden" / "golden_256.json").write_text(json.dumps(big, indent=1))


if __name__ == "__main__":
    main()
