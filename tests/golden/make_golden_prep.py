"""Generate tests/golden/prep_cases.npz from the reference's own preprocessing functions
(utility.py: normalize_feature, write_matrix_image_v1, near_interpolation1, near_interpolation1a,
write_matrix_array_v1; select_valuesPosition1_2 -> select_cases.npz), executed in memory through ref_loader.  Build container only.

usage: python tests/golden/make_golden_prep.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402


def contact_triples(rng, W, d, start, fill):
    """Upper-triangle bin pairs of a W-bin window with Hi-C-like values (decay off the diagonal,
    a fraction `fill` of pairs present, a few negatives for the clamp)."""
    ii, jj = np.triu_indices(W)
    keep = rng.random(len(ii)) < fill
    ii, jj = ii[keep], jj[keep]
    base = 50.0 / (1.0 + (jj - ii))
    val = base[:, None] * rng.gamma(2.0, 0.5, size=(len(ii), d))
    val[rng.random(val.shape) < 0.05] = 0.0
    val[rng.random(val.shape) < 0.01] *= -1.0
    pos = np.stack([ii + start, jj + start], axis=1).astype(np.int64)
    return val, pos


def make_select_cases():
    """utility.py:1331-1364 select_valuesPosition1_2 on a 30-bin triangle, the three border types."""
    util = ref_loader.load_utility(["select_valuesPosition1_2"])
    rng = np.random.default_rng(3)
    W = 30
    ii, jj = np.triu_indices(W)
    position = np.stack([ii + 5, jj + 5, np.arange(len(ii)) + 1000], axis=1).astype(np.int64)
    x = rng.random((len(ii), 3))
    out = {"position": position, "x": x}
    res = 10000
    cases = [(0, 80000, 250000, 80000, 250000), (0, 60000, 150000, 200000, 330000),
             (1, 80000, 250000, 80000, 250000), (2, 80000, 250000, 100000, 300000)]
    for k, (bt, p1, p2, p1a, p2a) in enumerate(cases):
        xs, b1 = util["select_valuesPosition1_2"](position, x, "", p1, p2, p1a, p2a, res, bt)
        out["sel%d_args" % k] = np.array([bt, p1, p2, p1a, p2a, res])
        out["sel%d_idx" % k] = b1
        out["sel%d_x" % k] = xs
    path = os.path.join(HERE, "select_cases.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


def main():
    if not ref_loader.available():
        raise SystemExit("reference not present")
    make_select_cases()
    util = ref_loader.load_utility(["normalize_feature", "write_matrix_image_v1", "near_interpolation1",
                                    "near_interpolation1a", "write_matrix_array_v1"])
    util["THRESH1"] = 1e-05
    out = {}
    cases = [("a", 1, 12, 3, 100, 0.7), ("b", 2, 23, 5, 0, 0.5), ("c", 3, 40, 2, 7, 0.85)]
    names = []
    for name, seed, W, d, start, fill in cases:
        rng = np.random.default_rng(seed)
        val, pos = contact_triples(rng, W, d, start, fill)
        p = "prep_%s_" % name
        out[p + "value"] = val
        out[p + "pos"] = pos
        for tag, xmin, xmax in (("auto", -1, -1), ("fixed", 0.0, 40.0)):
            x1, vec1, xmin_o, xmax_o = util["normalize_feature"](val.copy(), xmin, xmax)
            out[p + "norm_" + tag] = x1
            out[p + "vec1_" + tag] = vec1
            out[p + "lim_" + tag] = np.array([xmin_o, xmax_o])
        x = np.log(1 + out[p + "norm_auto"])
        out[p + "x"] = x
        mtx1, start_region = util["write_matrix_image_v1"](x, pos, "")
        out[p + "mtx"] = mtx1.copy()
        out[p + "start"] = np.array([start_region])
        filled = np.stack([util["near_interpolation1"](mtx1[:, :, c].copy(), 3) for c in range(d)], axis=2)
        out[p + "filled"] = filled
        # the general-block variant on a rectangular cut of the same image
        rect = mtx1[: W - 3, 2:, 0].copy()
        out[p + "rect"] = rect.copy()
        out[p + "rect_filled"] = util["near_interpolation1a"](rect, 3)
        data1, pos_idx, serial = util["write_matrix_array_v1"](filled, start_region, "", 1)
        out[p + "data1"] = data1
        out[p + "pos_idx"] = pos_idx
        out[p + "serial"] = serial
        names.append(name)
    out["names"] = np.array(names)
    path = os.path.join(HERE, "prep_cases.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
