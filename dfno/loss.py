"""``dfno.loss``: distributed losses under their reference module path."""
from dfno_b200.models.loss import DistributedMSELoss, DistributedRelativeLpLoss   # noqa: F401
