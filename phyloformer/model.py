"""Drop-in for the reference's phyloformer/model.py (inference surface only)."""
from phyloformer_b200.model import Phyloformer  # noqa: F401
