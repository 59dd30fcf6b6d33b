"""smeagol_b200 -- B200-native (sm_100a) implementation of SMEAGOL's PWM scanning hot path.

Drop-in for the reference's ``smeagol.encode`` / ``smeagol.models.PWMModel`` / ``smeagol.scan.scan_sequences`` (plus the
scan calls of ``smeagol.enrich`` and a batched ``smeagol.variant``): ``import smeagol_b200 as smeagol``.
"""
from . import encode, io, models, scan, utils  # noqa: F401
from . import enrich, matrices, variant  # noqa: F401
from .seqrecord import Seq, SeqRecord  # noqa: F401

__version__ = "0.1.0"
