"""Drop-in for the reference's phyloformer/data.py (inference surface only)."""
from phyloformer_b200.data import ALPHABET, LOOKUP, load_alignment, load_alignment_idx  # noqa: F401
