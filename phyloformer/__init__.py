"""Import-name shim: `phyloformer.model` / `phyloformer.data` resolve to the B200-native
implementation in `phyloformer_b200`, so code written against the reference
(`from phyloformer.model import Phyloformer`, infer_alns.py:10-11) runs unchanged."""
