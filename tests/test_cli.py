"""The command-line contract: option names and code defaults equal the reference's own
parse_args() (fixture: tests/golden/cli_defaults.json, produced by running it)."""
import json
import os

import pytest

from phylo_hmrf_b200 import cli

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "cli_defaults.json")))


def test_defaults_match_reference_parse_args():
    opts = vars(cli.parse_args([]))
    assert opts == GOLD and len(opts) == 33
    # the code defaults differ from the README
    assert opts["beta1"] == "0.5" and opts["num_neighbor"] == "8" and opts["estimate_type"] == "0"


def test_flags_parse_like_the_reference_example():
    o = cli.parse_args(["-n", "20", "-r", "1", "--reload", "0", "--chromvec", "21,22", "--miter", "100"])
    assert (o.num_states, o.run_id, o.reload, o.chromvec, o.miter) == ("20", "1", "0", "21,22", "100")


def test_raw_loading_path_builds_and_caches_the_inputs(tmp_path, monkeypatch):
    """`--reload 0` (phylo_hmrf.py:1622-1706): species / path lists, x_max = median of the per-species maxima,
    load_data_chromosome2 with the option values, caches written for the next `--reload 1` run."""
    import numpy as np
    from phylo_hmrf_b200 import hmrf, loader
    (tmp_path / "edge.1.txt").write_text("0\t1\n1\t2\n1\t3\n")
    (tmp_path / "species_name.1.txt").write_text("spA\nspB\n")
    (tmp_path / "path_list.txt").write_text("%s/a\n%s/b\n" % (tmp_path, tmp_path))
    seen = {}

    def fake_quant(chrom_vec, resolution, ref_filename, filename_list, species):
        seen["quant"] = (list(chrom_vec), resolution, ref_filename, list(filename_list), list(species))
        m = np.zeros((2, 10))
        m[:, 6] = [30.0, 50.0]
        return m

    def fake_load(chrom_vec, x_max, x_min, resolution, num_neighbor, filter_mode, sigma, diagonal_typeId,
                  ref_filename, filename_list, species, data_path, annotation="", device=0):
        seen["load"] = (list(chrom_vec), x_max, x_min, resolution, num_neighbor, filter_mode, sigma, diagonal_typeId)
        return np.ones((6, 2)), [[6, 0, 6, 3, 3, 0, 0, 0, 1, 21]], [np.zeros((4, 3))]

    class Stub:
        def __init__(self, **kw):
            seen["model"] = kw

        def fit_accumulate_test(self, samples, len_vec, threshold, filename, miter):
            return (np.zeros(1),) * 4 + (np.zeros(1), np.zeros(1), np.zeros(6))

        def close(self):
            pass

    monkeypatch.setattr(loader, "quantile_contact_vec", fake_quant)
    monkeypatch.setattr(loader, "load_data_chromosome2", fake_load)
    monkeypatch.setattr(hmrf, "phyloHMRF", Stub)
    monkeypatch.chdir(tmp_path)
    out = tmp_path / "out"
    o = cli.parse_args(["-p", str(tmp_path), "--output", str(out), "--reload", "0", "--chromvec", "21,22",
                        "--resolution", "50000", "--num_neighbor", "8", "--filter_mode", "0", "-w", "0.25", "--dtype", "1"])
    cli.run(o)
    assert seen["quant"][0] == [21, 22] and seen["quant"][2].endswith("hg38.chrom.sizes") and seen["quant"][4] == ["spA", "spB"]
    assert seen["load"] == ([21, 22], 40.0, 0, 50000, 8, 0, 0.25, 1)
    assert seen["model"]["n_samples"] == 6 and seen["model"]["n_features"] == 2
    for f in ("data.50Kb.observed.0.npy", "edgelist.50Kb.observed.0.npy", "lenvec.50Kb.observed.0.txt"):
        assert (out / f).exists()
    assert (tmp_path / "chrom_quantile_test.txt").exists()
