"""Persistent bookkeeping for one experiment tag: which FASTA sets map to which hex names, what
each sketch base name means, and every cardinality computed so far.

Drop-in for the reference module of the same name (reference lib/species_specifics.py:8-97): same
class, attributes, pickle file names and pickle contents (plain dicts), so a sketchdb directory
written by either implementation can be continued by the other.  Nothing here touches the GPU.
"""
import os
import pickle
import shutil
from typing import Dict

FASTA_SUFFIXES = (".fa.gz", ".fasta.gz", ".fna.gz", ".fasta", ".fa")


class SpeciesSpecifics:
    """State that outlives a single command.

    fastahex   {''.join(sorted basenames): hex digest / hex sum}      dandd_fastahex.pickle
    sketchinfo {sketch base name: {sketchbase, files, ngen, kval, registers}}  dandd_sketchinfo.pickle
    cardkey    {full sketch path: cardinality}               <tag>_<tool>_cardinalities.pickle
    """

    def __init__(self, tag: str, genomedir: str, sketchdir: str, kstart: int, tool: str, flist_loc=None):
        self.tag = tag
        self.sketchdir = sketchdir
        self.fastahex = self._read_fastahex()
        self.cardkey = self._read_cardkey(tool=tool)
        self.inputdir = genomedir
        self.card0 = []
        self.kstart = kstart
        self.orderings = None
        self.flist_loc = flist_loc
        self.sketchinfo = self._read_sketchinfo()

    # -- locations ------------------------------------------------------------------------------
    def _fastahex_loc(self) -> str:
        return os.path.join(self.sketchdir, "dandd_fastahex.pickle")

    def _sketchinfo_loc(self) -> str:
        return os.path.join(self.sketchdir, "dandd_sketchinfo.pickle")

    def _cardkey_loc(self, tool: str) -> str:
        return os.path.join(self.sketchdir, f"{self.tag}_{tool}_cardinalities.pickle")

    # -- reading ----------------------------------------------------------------------------------
    def read_pickle(self, filepath) -> Dict:
        """The stored dict, the .bkp copy if the main file is damaged, {} if there is nothing
        (reference :23-38; there a damaged .bkp raises, here it also yields {})."""
        for candidate in (filepath, filepath + ".bkp"):
            if not os.path.exists(candidate):
                if candidate == filepath:
                    return dict()
                continue
            try:
                with open(candidate, "rb") as fh:
                    return pickle.load(fh)
            except (pickle.UnpicklingError, EOFError):
                continue
        return dict()

    def _read_fastahex(self):
        return self.read_pickle(self._fastahex_loc())

    def _read_sketchinfo(self) -> Dict[str, Dict]:
        return self.read_pickle(self._sketchinfo_loc())

    def _read_cardkey(self, tool) -> Dict[str, float]:
        return self.read_pickle(self._cardkey_loc(tool))

    def update(self, tool) -> None:
        """Re-read all three dictionaries from disk (used after unpickling a tree)."""
        self.fastahex = self._read_fastahex()
        self.cardkey = self._read_cardkey(tool=tool)
        self.sketchinfo = self._read_sketchinfo()

    # -- writing: always to .bkp first, then copied over the live file (reference :57-69,83-89) ----
    @staticmethod
    def _dump_via_backup(obj, loc: str) -> None:
        os.makedirs(os.path.dirname(loc) or ".", exist_ok=True)
        with open(loc + ".bkp", "wb") as fh:
            pickle.dump(obj=obj, file=fh)
        shutil.copyfile(loc + ".bkp", loc)

    def _save_fastahex(self) -> None:
        self._dump_via_backup(self.fastahex, self._fastahex_loc())

    def _save_sketchinfo(self) -> None:
        self._dump_via_backup(self.sketchinfo, self._sketchinfo_loc())

    def save_references(self, fast=False) -> None:
        if fast:
            return
        self._save_fastahex()
        self._save_sketchinfo()

    def save_cardkey(self, tool: str, fast=False) -> None:
        if fast:
            return
        self._dump_via_backup(self.cardkey, self._cardkey_loc(tool))

    # -- inputs ---------------------------------------------------------------------------------------
    def retrieve_fasta_files(self, full=True) -> list:
        """Every entry of the input directory.  The reference compiles an extension filter but never
        applies it (`if reg_compile` is always true, reference :93-94), so every file counts as a
        FASTA; FASTA_SUFFIXES documents what the filter was meant to accept."""
        names = list(os.listdir(self.inputdir))
        return [os.path.join(self.inputdir, n) for n in names] if full else names
