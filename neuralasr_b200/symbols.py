"""Output symbol table and the string post-processing of decoded id sequences (host side).

Behavioural counterpart of the reference's ``symbols.py`` (``Symbols``, lines 8-68): ids are handed out in
insertion order, the preprocessing inserts ``<padding>`` first and ``<blank>`` last
(``preprocess_mfcc.py:81-92``) so that blank = num_classes - 1, which is the blank the CTC kernels assume;
``convert_to_str`` (``symbols.py:53-60``) strips the label context from every symbol, drops blanks and
turns ``_`` into a space.  The file format is ``<symbol> <id>`` per line, sorted by symbol.
"""
from __future__ import annotations

import os

BLANK = "<blank>"
PADDING = "<padding>"


class Symbols:
    def __init__(self, label_context=0, filename=None):
        self.label_context = int(label_context)
        self.blank, self.padding = BLANK, PADDING
        self.filename = filename
        self.sym_to_id, self.id_2_sym = {}, {}
        self.counter = 0
        if filename and os.path.exists(filename):
            self.read(filename)

    # -- construction -----------------------------------------------------------------------
    def read(self, filename):
        with open(filename) as f:
            pairs = [line.split() for line in f if line.strip()]
        self.sym_to_id = {sym: int(idx) for sym, idx in pairs}
        self.id_2_sym = {idx: sym for sym, idx in self.sym_to_id.items()}
        self.counter = max(self.id_2_sym, default=-1) + 1

    def insert_sym(self, sym):
        idx = self.sym_to_id.get(sym)
        if idx is None:
            idx = self.counter
            self.sym_to_id[sym], self.id_2_sym[idx] = idx, sym
            self.counter += 1
        return idx

    def insert_blank(self):
        return self.insert_sym(self.blank)

    def insert_padding(self):
        return self.insert_sym(self.padding)

    # -- lookups ----------------------------------------------------------------------------
    def get_padding_id(self):
        return self.sym_to_id[self.padding]

    def get_blank_id(self):
        return self.sym_to_id[self.blank]

    def get_id(self, sym):
        return self.sym_to_id[sym]

    def get_sym(self, idx):
        return self.id_2_sym[int(idx)]

    def get_all_ids(self, _unused=None):
        return list(self.sym_to_id.values())

    @property
    def num_classes(self):
        """What the reference passes to every network as ``num_classes`` (``tfnetwork.py:18``)."""
        return self.counter

    # -- decoded ids -> text ----------------------------------------------------------------
    def convert_to_str(self, ids):
        k = self.label_context
        parts = (self.get_sym(i) for i in ids)
        text = "".join(p[k:-k] for p in parts) if k > 0 else "".join(parts)
        return text.replace(self.blank, "").replace("_", " ")

    def write(self, filename=None):
        filename = filename or self.filename
        with open(filename, "w") as f:
            for sym in sorted(self.sym_to_id):
                f.write("%s %d\n" % (sym, self.sym_to_id[sym]))
