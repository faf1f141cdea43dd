"""Site-saturation mutation lists, as ``analysis/SSM.py:16-29`` builds them."""
from typing import List, Optional

from .datasets import ALPHABET, Mutation


def get_ssm_mutations(pdb: dict) -> List[Optional[str]]:
    """``wt + str(pos) + mut`` for the 20 letters (identity included) at every residue;
    a single ``None`` for each gap ('-') residue."""
    out: List[Optional[str]] = []
    for pos, wt in enumerate(pdb["seq"]):
        if wt != "-":
            out.extend(wt + str(pos) + m for m in ALPHABET[:-1])
        else:
            out.append(None)
    return out


def to_mutation_objects(mutation_list, pdb_name: str = "") -> List[Optional[Mutation]]:
    """Parse 'A12G'-style strings into ``Mutation`` objects (custom_inference.py:79-90)."""
    out: List[Optional[Mutation]] = []
    for m in mutation_list:
        if m is None:
            out.append(None)
            continue
        m = m.strip()
        wt, pos, mut = m[0], int(m[1:-1]), m[-1]
        assert wt in ALPHABET, f"Wild type residue {wt} invalid, please try again with one of the following options: {ALPHABET}"
        assert mut in ALPHABET, f"Mutant residue {mut} invalid, please try again with one of the following options: {ALPHABET}"
        out.append(Mutation(position=pos, wildtype=wt, mutation=mut, ddG=None, pdb=pdb_name))
    return out
