"""Loader parity on CPU: phylo_b200.loader against the golden produced by the reference's own lines (runner.py:83-115)."""
import os

import numpy as np
import pytest

from phylo_b200 import loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_form_dataset_matches_reference_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "loader.npz"))
    dd = loader.form_dataset_from_strings([str(s) for s in z["toy_strings"]], loader.ALPHABET_DIR_BLANK)
    np.testing.assert_array_equal(dd["genome"], z["toy_genome"])
    assert dd["taxa"] == [str(t) for t in z["toy_taxa"]]
    dd = loader.form_dataset_from_strings([str(s) for s in z["primate_strings"]], loader.ALPHABET_DIR_BLANK)
    np.testing.assert_array_equal(dd["genome"], z["primate_genome"].astype(np.float64))


def test_load_dataset_primate_pickle():
    dd = loader.load_dataset("primate_data", os.path.join(ROOT, "data"))
    assert dd["genome"].shape == (12, 898, 4) and dd["genome"].dtype == np.float64
    assert (dd["genome"].sum(axis=2) == 4).sum() == 30            # 30 gap characters -> all-ones
    dd = loader.load_dataset("primate_data_wang", os.path.join(ROOT, "data"))
    assert dd["genome"].shape == (9, 738, 4)


def test_unknown_character_raises_like_reference():
    with pytest.raises(KeyError):                                   # alphabet_dir[ch] in runner.py:111
        loader.form_dataset_from_strings(["ACGN", "ACGT"], loader.ALPHABET_DIR_BLANK)
    dd = loader.form_dataset_from_strings(["ACGN", "ACGT"], loader.ALPHABET_DIR_BLANK, unknown_as_gap=True)
    assert dd["genome"][0, 3].tolist() == [1, 1, 1, 1]


def test_toy_and_synthetic():
    dd = loader.load_dataset("load_strings")
    assert dd["genome"].shape == (4, 10, 4) and dd["genome"].sum() == 40
    dd = loader.load_dataset("synthetic_27x1949")
    assert dd["genome"].shape == (27, 1949, 4)
    assert (dd["genome"].sum(axis=2) == 1).all()
    np.testing.assert_array_equal(dd["genome"], loader.synthetic_alignment(27, 1949, 0)["genome"])
    with pytest.raises(ValueError):
        loader.load_dataset("no_such_dataset")


def test_unknown_as_gap_is_one_uniform_option(tmp_path):
    """DS7 holds 'N' (the reference raises KeyError at runner.py:111): one flag, applied to every string dataset."""
    import pickle
    os.makedirs(tmp_path / "hohna_datasets")
    for n in (7, 10):
        with open(tmp_path / "hohna_datasets" / ("DS%d.pickle" % n), "wb") as f:
            pickle.dump({"a": "ACGTN", "b": "AC.T-"}, f)
    for n in (7, 10):
        dd = loader.load_dataset("hohna_data_%d" % n, str(tmp_path))
        assert dd["genome"][0, 4].tolist() == [1, 1, 1, 1] and dd["genome"][1, 2].tolist() == [1, 1, 1, 1]
        with pytest.raises(KeyError):
            loader.load_dataset("hohna_data_%d" % n, str(tmp_path), unknown_as_gap=False)
    from phylo_b200.runner import parse_args
    assert parse_args([]).unknown_as_gap is True and parse_args(["--unknown_as_gap=false"]).unknown_as_gap is False


def test_corona_data_falls_back_to_the_repaired_betacorona_alignment(tmp_path):
    """coronavirus.p is absent from the reference repository; betacorona1.p (taxa: a 1-tuple with 16 names for 17 rows)
    is the stand-in, repaired.  The shipped state-mask copy decodes to the same [17,3260,4] genome."""
    dd = loader.load_dataset("corona_data", os.path.join(ROOT, "data"))
    assert dd["genome"].shape == (17, 3260, 4) and len(dd["taxa"]) == 17 and len(set(dd["taxa"])) == 17
    s = dd["genome"].sum(axis=2)
    assert set(np.unique(s)) == {1.0, 4.0} and 0.16 < (s == 4).mean() < 0.17
    broken = {"taxa": (["S%d" % i for i in range(16)],), "gemome": dd["genome"][:, :5]}
    fixed = loader.repair_datadict(broken)
    assert fixed["taxa"] == ["S%d" % i for i in range(17)] and fixed["genome"].shape == (17, 5, 4)
    import pickle
    os.makedirs(tmp_path / "betacoronavirus")
    with open(tmp_path / "betacoronavirus" / "betacorona1.p", "wb") as f:
        pickle.dump({"taxa": (["S%d" % i for i in range(16)],), "genome": dd["genome"]}, f)
    np.testing.assert_array_equal(loader.load_dataset("corona_data", str(tmp_path))["genome"], dd["genome"])
