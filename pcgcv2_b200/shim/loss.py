"""Same-named module as the reference's ``loss.py`` (imported by its trainer.py:10): the occupancy loss as ONE fused
kernel over the ground truth's hash table instead of D2H + ``np.isin`` + torch BCE (``pcgcv2_b200.train``)."""
from pcgcv2_b200.train import get_bce, get_bits, get_cls_metrics, get_metrics  # noqa: F401
