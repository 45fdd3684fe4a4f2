import torch

from ..functional import ZeroVolumeCorrectorFunction
from ._wire import collect_from_everyone
from .sum_reduce import SumReduce


class DistributedMSELoss(torch.nn.Module):
    """Global mean squared error of a domain-decomposed field: local sum of squares -> SumReduce onto the
    root -> divide by the global element count; a well-defined scalar 0 off the root."""

    def __init__(self, P_x, reduction="mean"):
        super().__init__()
        self.P_x, self.reduction = P_x, reduction
        self.P_0 = P_x.create_partition_inclusive([0]).create_cartesian_topology_partition([1] * P_x.dim)
        self.to_root = SumReduce(P_x, self.P_0)
        self.count = None

    def forward(self, y_hat, y):
        if self.count is None:
            self.count = sum(collect_from_everyone(int(y.numel()) if self.P_x.active else 0))
        total = self.to_root(torch.sum((y_hat - y) ** 2).reshape(1))
        if self.P_0.active and self.reduction == "mean":
            total = total / self.count
        return ZeroVolumeCorrectorFunction.apply(total.reshape(()) if total.numel() else total)
