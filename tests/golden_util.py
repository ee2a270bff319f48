"""Helpers to read tests/golden/*.npz (written by oracle/make_goldens.py from the unmodified reference)."""
import json
import os

import numpy as np

from oracle import fixtures, replay, stryke_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz"))


class Golden:
    def __init__(self, name, subdir=""):
        self.z = np.load(os.path.join(GOLDEN_DIR, subdir, name + ".npz"))
        self.meta = json.loads(str(self.z["meta"]))
        self.prj = getattr(fixtures, self.meta["fixture"])(**self.meta["kwargs"])
        self.model = O.Model.from_project(self.prj)
        self.pops = {(r["Species"], r["Scenario"]): r for r in self.prj["population"]}
        self.cells = self.meta["cells"]

    def species(self, cm):
        prow = self.pops[(cm["species"], cm["scenario"])]
        return O.species_record(self.model, prow, self.meta["families"].get(cm["species"]))

    def draws(self, ci):
        return replay.unpack_cell(self.z, f"c{ci}_")

    def ref(self, ci, key):
        return self.z[f"c{ci}_{key}"]

    def ref_state_daily(self, ci):
        mv, st, vals = self.ref(ci, "sd_move"), self.ref(ci, "sd_state"), self.ref(ci, "sd_vals")
        return {(int(m), self.model.nodes[int(s)]): (float(v[0]), float(v[1])) for m, s, v in zip(mv, st, vals)}
