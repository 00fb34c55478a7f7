"""Stand-in for the absent ``torch-geometric==2.6.1`` wheel (reference requirements.txt:11).

Used ONLY by oracle/make_golden.py so that the UNMODIFIED reference sources under
/root/reference (models.py, trainer.py, data.py) can be imported in the build container and
golden vectors recorded.  Every class here is the oracle's restatement from oracle/pyg.py.
"""
__version__ = "2.6.1+oracle-shim"
from . import nn, data  # noqa: F401
