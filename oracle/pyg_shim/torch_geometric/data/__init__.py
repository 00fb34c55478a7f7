import torch.utils.data

from oracle.pyg import Batch, Data  # noqa: F401


class Dataset(torch.utils.data.Dataset):
    """torch_geometric.data.Dataset stand-in (the reference only subclasses it, data.py:80)."""

    def __init__(self, *args, **kwargs):
        super().__init__()
