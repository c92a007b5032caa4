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
