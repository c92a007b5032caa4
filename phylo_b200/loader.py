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


def repair_datadict(d: dict) -> dict:
    """The betacoronavirus pickles as shipped are not loadable by the reference's class: ``taxa`` is a 1-tuple holding
    a 16-name list for 17 genome rows (betacorona1.p) and the genome key is misspelt ``gemome`` (betacorona2.p).
    Returns a well-formed ``{'taxa': [N names], 'genome': [N,S,A]}``."""
    genome = np.asarray(d["genome"] if "genome" in d else d["gemome"], dtype=np.float64)
    taxa = d.get("taxa", [])
    if isinstance(taxa, tuple) and len(taxa) == 1 and isinstance(taxa[0], (list, tuple)):
        taxa = taxa[0]
    taxa = [str(t) for t in taxa]
    for i in range(len(taxa), genome.shape[0]):
        taxa.append("S" + str(i))
    return {"taxa": taxa[:genome.shape[0]], "genome": genome}


def _corona(data_dir: str) -> dict:
    """runner.py:159-160 reads data/coronavirus.p as a ready datadict.  That file is not in the reference repository
    (.MISSING_LARGE_BLOBS); the closest shipped alignment is data/betacoronavirus/betacorona1.p (17 x 3260, 16.5 % gap
    sites), which is used -- repaired -- when coronavirus.p is absent (a copy of its state masks ships in data/)."""
    path = os.path.join(data_dir, "coronavirus.p")
    if os.path.exists(path):
        return repair_datadict(_read_pickle(path))
    path = os.path.join(data_dir, "betacoronavirus", "betacorona1.p")
    if os.path.exists(path):
        return repair_datadict(_read_pickle(path))
    path = os.path.join(data_dir, "betacorona1_codes.npz")
    if os.path.exists(path):
        codes = np.load(path)["codes"]
        genome = ((codes[..., None] >> np.arange(4, dtype=np.uint8)) & 1).astype(np.float64)
        return {"taxa": ["S" + str(i) for i in range(genome.shape[0])], "genome": genome}
    raise FileNotFoundError(os.path.join(data_dir, "coronavirus.p"))


def load_dataset(name: str, data_dir: str = "data", unknown_as_gap: bool = True) -> dict:
    """The ``exec(args.dataset + ' = True')`` switch of runner.py:81,117-184 by name.

    ``unknown_as_gap`` applies to every string dataset alike: characters outside the alphabet dict ('N' in DS7, '.', 'n'
    in DS10/DS11) become the all-ones mask; with False they raise KeyError like the reference's ``alphabet_dir[ch]``
    (runner.py:111).  Extra names: ``hohna_data_9`` .. ``_11``, ``fish_data``, ``synthetic_NxS`` (e.g.
    synthetic_64x10000) for the benchmark shapes.
    """
    def strings(d, alphabet):
        return form_dataset_from_strings(list(d.values()), alphabet, unknown_as_gap=unknown_as_gap)

    if name in ("hohna_data", "hohna_data_1"):
        return strings(_hohna(data_dir, 1), ALPHABET_DIR_BLANK)
    if name.startswith("hohna_data_"):
        return strings(_hohna(data_dir, int(name.rsplit("_", 1)[1])), ALPHABET_DIR_BLANK)
    if name == "corona_data":
        return _corona(data_dir)
    if name == "fish_data":      # data/fish.p (12 x 1047) ships with the reference but has no branch in runner.py
        return strings(_read_pickle(os.path.join(data_dir, "fish.p")), ALPHABET_DIR_BLANK)
    if name == "primate_data":
        return strings(_read_pickle(os.path.join(data_dir, "primate.p")), ALPHABET_DIR_BLANK)
    if name == "primate_data_wang":
        return strings(_read_pickle(os.path.join(data_dir, "primates_small.p")), ALPHABET_DIR)
    if name == "simulate_data":
        g = simulateDNA(3, 5)
        return {"taxa": ["S" + str(i) for i in range(g.shape[0])], "genome": g}
    if name == "load_strings":
        return form_dataset_from_strings(["ACTTTGAGAG", "ACTTTGACAG", "ACTTTGACTG", "ACTTTGACTC"], ALPHABET_DIR)
    if name.startswith("synthetic_"):
        n, s = name[len("synthetic_"):].split("x")
        return synthetic_alignment(int(n), int(s))
    raise ValueError("unknown dataset %r (known: %s, synthetic_NxS)" % (name, ", ".join(DATASETS)))
