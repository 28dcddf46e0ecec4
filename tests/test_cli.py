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


def test_raw_loading_is_refused_with_a_clear_message(tmp_path):
    (tmp_path / "edge.1.txt").write_text("0\t1\n1\t2\n1\t3\n")
    o = cli.parse_args(["-p", str(tmp_path), "--output", str(tmp_path), "--reload", "0"])
    with pytest.raises(SystemExit, match="--reload 1"):
        cli.run(o)
