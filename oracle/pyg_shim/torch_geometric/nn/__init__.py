from oracle.pyg import GATConv, GATv2Conv, GCNConv, GraphConv, Sequential  # noqa: F401
from . import norm  # noqa: F401
