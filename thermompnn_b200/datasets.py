"""API types shared with the reference's callers."""
from dataclasses import dataclass
from typing import Optional

ALPHABET = "ACDEFGHIKLMNPQRSTVWYX"  # transfer_model.py:11


@dataclass
class Mutation:
    """Same fields as the reference's ``datasets.Mutation`` (datasets.py:25-31);
    ``position`` is 0-based into the concatenated chain sequence."""
    position: int
    wildtype: str
    mutation: str
    ddG: Optional[float] = None
    pdb: Optional[str] = ""
