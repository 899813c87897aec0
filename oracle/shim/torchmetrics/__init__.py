"""ORACLE shim: makes ``import torchmetrics.functional`` resolve to the CPU
restatement in ``oracle/torchmetrics_port.py`` so the UNMODIFIED reference modules
(``models/utils.py:4-8``, ``report.py:3-7``) import in this container.  Test
infrastructure only."""
__version__ = "0.11.4+oracle-port"
from . import functional  # noqa: F401
