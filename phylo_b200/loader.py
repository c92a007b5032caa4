"""Alignment loaders: the dataset switch of the reference's runner.py:83-192 as functions.

Each loader returns the reference's ``datadict`` -- ``{'taxa': [str]*N, 'genome': ndarray[N,S,4] float64}`` --
so ``VCSMC(datadict, K, args)`` receives exactly what the reference's class receives (vcsmc.py:110-118).
Packing into 4-bit device codes happens once, inside VCSMC (ops.pack_alignment).
"""
from __future__ import annotations

import io
import os
import pickle
import random
import zipfile
from typing import Dict, List, Optional

import numpy as np

# runner.py:83-96
ALPHABET_DIR = {"A": [1, 0, 0, 0], "C": [0, 1, 0, 0], "G": [0, 0, 1, 0], "T": [0, 0, 0, 1]}
ALPHABET_DIR_BLANK = dict(ALPHABET_DIR, **{"-": [1, 1, 1, 1], "?": [1, 1, 1, 1]})


def form_dataset_from_strings(genome_strings: List[str], alphabet_dir: Dict[str, List[int]], alphabet_num: int = 4,
                              unknown_as_gap: bool = False) -> dict:
    """runner.py:107-115.  ``unknown_as_gap`` maps characters outside the dict (e.g. 'N', '.', 'n' in DS7/DS10/
    DS11, where the reference raises KeyError) to the all-ones mask; off by default to keep reference behaviour."""
    n, s = len(genome_strings), len(genome_strings[0])
    table = np.zeros((256, alphabet_num))
    known = np.zeros(256, dtype=bool)
    for ch, row in alphabet_dir.items():
        table[ord(ch)] = row
        known[ord(ch)] = True
    raw = np.frombuffer("".join(genome_strings).encode("latin-1"), dtype=np.uint8).reshape(n, s)
    if not known[raw].all():
        if not unknown_as_gap:
            bad = chr(int(raw[~known[raw]][0]))
            raise KeyError(bad)  # what alphabet_dir[ch] raises in the reference
        table[~known] = 1.0
    genome = table[raw]
    return {"taxa": ["S" + str(i) for i in range(n)], "genome": genome}


def simulateDNA(nsamples: int, seqlength: int, alphabet: Optional[np.ndarray] = None) -> np.ndarray:
    """runner.py:100-104 (python ``random``, unseeded in the reference)."""
    alphabet = np.eye(4) if alphabet is None else alphabet
    genomes = np.zeros([nsamples, seqlength, alphabet.shape[0]])
    for n in range(nsamples):
        genomes[n] = np.array([random.choice(alphabet) for _ in range(seqlength)])
    return genomes


def synthetic_alignment(n_taxa: int, n_sites: int, seed: int = 0) -> dict:
    """i.i.d. uniform nucleotides, no gaps: the benchmark's synthetic alignments of the named shapes."""
    rng = np.random.Generator(np.random.PCG64(seed))
    genome = np.eye(4)[rng.integers(0, 4, (n_taxa, n_sites))]
    return {"taxa": ["S" + str(i) for i in range(n_taxa)], "genome": genome}


def _read_pickle(path: str):
    with open(path, "rb") as f:
        return pickle.load(f)


def _hohna(data_dir: str, n: int):
    """runner.py:117-156 expects data/hohna_datasets/DSn.pickle; also read it straight from the shipped zip."""
    path = os.path.join(data_dir, "hohna_datasets", "DS%d.pickle" % n)
    if os.path.exists(path):
        return _read_pickle(path)
    zpath = os.path.join(data_dir, "hohna_dataset_pickle.zip")
    if os.path.exists(zpath):
        with zipfile.ZipFile(zpath) as z:
            for name in z.namelist():
                if name.endswith("DS%d.pickle" % n):
                    return pickle.load(io.BytesIO(z.read(name)))
    raise FileNotFoundError(path)


DATASETS = ("primate_data", "corona_data", "hohna_data", "load_strings", "simulate_data", "primate_data_wang") + \
    tuple("hohna_data_%d" % i for i in range(1, 12))


def load_dataset(name: str, data_dir: str = "data") -> dict:
    """The ``exec(args.dataset + ' = True')`` switch of runner.py:81,117-184 by name.

    Extra names: ``synthetic_NxS`` (e.g. synthetic_64x10000) for the benchmark shapes.
    """
    if name in ("hohna_data", "hohna_data_1"):
        return form_dataset_from_strings(list(_hohna(data_dir, 1).values()), ALPHABET_DIR_BLANK)
    if name.startswith("hohna_data_"):
        n = int(name.rsplit("_", 1)[1])
        return form_dataset_from_strings(list(_hohna(data_dir, n).values()), ALPHABET_DIR_BLANK, unknown_as_gap=n > 8)
    if name == "corona_data":
        return _read_pickle(os.path.join(data_dir, "coronavirus.p"))  # runner.py:159-160: already a datadict
    if name == "primate_data":
        return form_dataset_from_strings(list(_read_pickle(os.path.join(data_dir, "primate.p")).values()), ALPHABET_DIR_BLANK)
    if name == "primate_data_wang":
        return form_dataset_from_strings(list(_read_pickle(os.path.join(data_dir, "primates_small.p")).values()), ALPHABET_DIR)
    if name == "simulate_data":
        g = simulateDNA(3, 5)
        return {"taxa": ["S" + str(i) for i in range(g.shape[0])], "genome": g}
    if name == "load_strings":
        return form_dataset_from_strings(["ACTTTGAGAG", "ACTTTGACAG", "ACTTTGACTG", "ACTTTGACTC"], ALPHABET_DIR)
    if name.startswith("synthetic_"):
        n, s = name[len("synthetic_"):].split("x")
        return synthetic_alignment(int(n), int(s))
    raise ValueError("unknown dataset %r (known: %s, synthetic_NxS)" % (name, ", ".join(DATASETS)))
