"""Extract the input tables of one of the reference's example workbooks into a small JSON project description.

Run in the build container only (needs /root/reference):
    python -m oracle.extract_example "Spreadsheet Interface/Examples/Run_of_River_Complex_Network_Cabot_Station.xlsx" \
        tests/golden/cabot_station_project.json

The workbook is parsed with the repository's own zip+XML reader (openpyxl is not installed, so the reference cannot read
it here), the sheets are laid out as the reference's ``worksheet_import`` expects (stryke.py:356-469) and written as
plain rows.  ``oracle/fixtures.cabot_station`` turns the rows into a project dict; the reference itself is then driven
through the attribute set its unit tests use (oracle/harness.py), because its spreadsheet path cannot run at HEAD
(SURVEY.md Appendix E2).  TEST INFRASTRUCTURE.
"""
from __future__ import annotations

import json
import math
import os
import sys
import tempfile


def _rows(df):
    out = []
    for rec in df.to_dict("records"):
        out.append({k: (None if isinstance(v, float) and math.isnan(v) else (v.item() if hasattr(v, "item") else v))
                    for k, v in rec.items() if not str(k).startswith("Unnamed")})
    return out


def extract(workbook, out_path):
    from stryke_b200 import simulation
    tmp = tempfile.mkdtemp(prefix="stryke_extract_")
    sim = simulation(os.path.dirname(workbook), os.path.join(tmp, "extract"), wks=os.path.basename(workbook))
    prj = dict(source=os.path.relpath(workbook, "/root/reference"), output_units=sim.output_units,
               nodes=_rows(sim.nodes), edges=_rows(sim.edges), unit_params=_rows(sim.unit_params.reset_index(drop=True)),
               facility_params=_rows(sim.facility_params.reset_index()), flow_scenarios=_rows(sim.flow_scenarios_df),
               operating_scenarios=_rows(sim.operating_scenarios_df), population=_rows(sim.pop))
    with open(out_path, "w") as fh:
        json.dump(prj, fh, indent=1)
    return prj


if __name__ == "__main__":
    wb = sys.argv[1]
    if not os.path.isabs(wb):
        wb = os.path.join("/root/reference", wb)
    p = extract(wb, sys.argv[2])
    print({k: (len(v) if isinstance(v, list) else v) for k, v in p.items()})
