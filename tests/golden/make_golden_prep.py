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


def write_loader_inputs(root, G):
    """Recreate the mini data set of loader_cases.npz as the files the loaders read."""
    os.makedirs(root, exist_ok=True)
    with open(os.path.join(root, "chrom.sizes"), "w") as f:
        f.write("chr1\t248956422\nchr3\t%d\n" % int(G["chrom_size"]))
    dirs = []
    for s in range(int(G["n_species"])):
        d = os.path.join(root, "hic_sp%d" % s)
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "chr3.%dK.txt" % (int(G["resolution"]) // 1000)), "w") as f:
            for a, b, v in zip(G["sp%d_x1" % s], G["sp%d_x2" % s], G["sp%d_v" % s]):
                f.write("%d\t%d\t%s\n" % (a, b, "NaN" if np.isnan(v) else repr(float(v))))
        dirs.append(d)
    np.savetxt(os.path.join(root, "chr3.synteny.txt"), G["synteny"], fmt="%d", delimiter="\t")
    return dirs


def make_loader_cases():
    """utility.py:2507-2570, 2631-2662, 824-863 (multi_contact_matrix3A, output_multi_contactMtx,
    mapping_Idx) and :2111-2189 (subregion1) on a mini chr3 (1 Mb bins, a synteny block across the
    hg38 centromere).  Python-2 classic division in the alignment is patched to ``//``."""
    import math
    import tempfile
    import pandas as pd
    patches = [("math.ceil(chrom_size/resolution)", "math.ceil(chrom_size//resolution)"),
               ("x1, x2 = x1/resolution, x2/resolution", "x1, x2 = x1//resolution, x2//resolution"),
               # pandas >= 2 hands out read-only views; the reference writes -1 over NaN in place
               ("np.asarray(data2[2])", "np.array(data2[2])")]
    util = ref_loader.load_utility(["mapping_Idx", "output_multi_contactMtx", "multi_contact_matrix3A", "subregion1",
                                    "multi_contact_matrix3A_single", "quantile_contact", "quantile_contact_vec"],
                                   patches)
    util.update({"pd": pd, "math": math, "os": os})
    rng = np.random.default_rng(11)
    res = 1000000
    G = {"resolution": np.array(res), "chrom_size": np.array(198295559), "n_species": np.array(3),
         "synteny": np.array([[60000000, 110000000, 50000000], [120000000, 140000000, 20000000]])}
    ii, jj = np.triu_indices(80)
    for s in range(3):
        keep = rng.random(len(ii)) < (0.9, 0.6, 0.75)[s]
        a, b = (ii[keep] + 60) * res, (jj[keep] + 60) * res
        v = 100.0 / (1.0 + (jj[keep] - ii[keep])) * rng.gamma(2.0, 0.5, size=keep.sum())
        v[rng.random(len(v)) < 0.02] = np.nan
        G["sp%d_x1" % s], G["sp%d_x2" % s], G["sp%d_v" % s] = a, b, v
    with tempfile.TemporaryDirectory() as root:
        dirs = write_loader_inputs(root, G)
        species = ["sp0", "sp1", "sp2"]
        data = util["multi_contact_matrix3A"]("3", res, os.path.join(root, "chrom.sizes"), dirs, species, "", 0)
        G["aligned_position"] = np.asarray(data.loc[:, [0, 1, 2]])
        G["aligned_x"] = np.asarray(data.loc[:, species], dtype=np.float64)
        G["quantiles"] = util["quantile_contact_vec"](["3"], res, os.path.join(root, "chrom.sizes"), dirs, species)
        points = [np.array([90279522, 93797661])]
        region_list, list1 = util["subregion1"](os.path.join(root, "chr3.synteny.txt"), 3, res, points, 0)
        G["region_list"] = np.asarray([list(map(int, r)) for r in region_list])
        G["list1"] = np.asarray([list(map(int, r)) for r in list1])
        region_list, list1 = util["subregion1"](os.path.join(root, "chr3.synteny.txt"), 3, res, [], 0)
        G["list1_nosplit"] = np.asarray([list(map(int, r)) for r in list1])
    path = os.path.join(HERE, "loader_cases.npz")
    np.savez_compressed(path, **G)
    print("wrote", path, os.path.getsize(path), "bytes", G["aligned_x"].shape, G["list1"].shape)


def main():
    if not ref_loader.available():
        raise SystemExit("reference not present")
    make_select_cases()
    make_loader_cases()
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
