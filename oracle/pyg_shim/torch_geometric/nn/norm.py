from oracle.pyg import GraphNorm  # noqa: F401
