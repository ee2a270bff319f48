"""Minimal .xlsx (OOXML) reader / sheet appender — openpyxl and xlrd are not available in this image.

``Workbook.frame(sheet, usecols, skiprows)`` reproduces what the reference gets from
``pd.read_excel(path, sheet_name=..., header=0, usecols="B:D", skiprows=n)`` (Stryke/stryke.py:356-469; sheet
geometry in SURVEY.md Appendix F).  ``append_sheets`` adds result sheets to an existing workbook the way
``pd.ExcelWriter(path, engine='openpyxl', mode='a')`` does at stryke.py:3895-3900.
"""
from __future__ import annotations

import datetime as _dt
import re
import shutil
import tempfile
import zipfile
from xml.etree import ElementTree as ET
from xml.sax.saxutils import escape

import numpy as np
import pandas as pd

NS = {"m": "http://schemas.openxmlformats.org/spreadsheetml/2006/main",
      "r": "http://schemas.openxmlformats.org/officeDocument/2006/relationships",
      "p": "http://schemas.openxmlformats.org/package/2006/relationships"}
_BUILTIN_DATE_FMTS = set(range(14, 23)) | set(range(45, 48)) | {27, 30, 36, 50, 57}
_CELL = re.compile(r"([A-Z]+)(\d+)")


def col_index(letters):
    n = 0
    for ch in letters:
        n = n * 26 + (ord(ch) - 64)
    return n - 1


def col_letters(idx):
    s = ""
    idx += 1
    while idx:
        idx, r = divmod(idx - 1, 26)
        s = chr(65 + r) + s
    return s


def _excel_date(serial):
    base = _dt.datetime(1899, 12, 30)      # 1900 date system, incl. the phantom 1900-02-29
    return pd.Timestamp(base + _dt.timedelta(days=float(serial)))


class Workbook:
    def __init__(self, path):
        self.path = path
        with zipfile.ZipFile(path) as z:
            names = set(z.namelist())
            self.shared = self._shared_strings(z) if "xl/sharedStrings.xml" in names else []
            self.date_styles = self._date_styles(z) if "xl/styles.xml" in names else set()
            wb = ET.fromstring(z.read("xl/workbook.xml"))
            rels = ET.fromstring(z.read("xl/_rels/workbook.xml.rels"))
            target = {r.get("Id"): r.get("Target") for r in rels.findall("p:Relationship", NS)}
            self.sheets = {}
            for s in wb.find("m:sheets", NS).findall("m:sheet", NS):
                t = target[s.get("{%s}id" % NS["r"])]
                self.sheets[s.get("name")] = t.lstrip("/") if t.startswith("/") else "xl/" + t
            self._cells = {}
            for name, part in self.sheets.items():
                self._cells[name] = self._sheet_cells(z.read(part))

    @staticmethod
    def _shared_strings(z):
        root = ET.fromstring(z.read("xl/sharedStrings.xml"))
        out = []
        for si in root.findall("m:si", NS):
            out.append("".join(t.text or "" for t in si.iter("{%s}t" % NS["m"])))
        return out

    @staticmethod
    def _date_styles(z):
        root = ET.fromstring(z.read("xl/styles.xml"))
        custom = {}
        nf = root.find("m:numFmts", NS)
        if nf is not None:
            for f in nf.findall("m:numFmt", NS):
                code = re.sub(r'"[^"]*"|\[[^\]]*\]|\\.', "", f.get("formatCode", ""))
                custom[int(f.get("numFmtId"))] = bool(re.search(r"[dmyhs]", code, re.I)) and not re.search(r"[0#]", code)
        dates = set()
        xfs = root.find("m:cellXfs", NS)
        if xfs is not None:
            for i, xf in enumerate(xfs.findall("m:xf", NS)):
                fid = int(xf.get("numFmtId", "0"))
                if fid in _BUILTIN_DATE_FMTS or custom.get(fid, False):
                    dates.add(i)
        return dates

    def _sheet_cells(self, xml):
        root = ET.fromstring(xml)
        cells = {}
        data = root.find("m:sheetData", NS)
        if data is None:
            return cells
        for row in data.findall("m:row", NS):
            for c in row.findall("m:c", NS):
                m = _CELL.match(c.get("r", ""))
                if not m:
                    continue
                col, r = col_index(m.group(1)), int(m.group(2)) - 1
                t = c.get("t", "n")
                v = c.find("m:v", NS)
                if t == "inlineStr":
                    is_ = c.find("m:is", NS)
                    val = "".join(x.text or "" for x in is_.iter("{%s}t" % NS["m"])) if is_ is not None else None
                elif v is None or v.text is None:
                    continue
                elif t == "s":
                    val = self.shared[int(v.text)]
                elif t in ("str", "e"):
                    val = v.text
                elif t == "b":
                    val = v.text == "1"
                else:
                    num = float(v.text)
                    if int(c.get("s", "0")) in self.date_styles:
                        val = _excel_date(num)
                    else:
                        val = int(num) if num.is_integer() and "." not in v.text and "E" not in v.text.upper() else num
                if val is None or (isinstance(val, str) and val == ""):
                    continue
                cells[(r, col)] = val
        return cells

    def frame(self, sheet, usecols=None, skiprows=0, dtype=None, expect=None):
        """``pd.read_excel(..., header=0, usecols=usecols, skiprows=skiprows)``.

        ``expect``: column names the header row must contain.  When the row at ``skiprows`` does not hold them, the
        rows within three of it are searched: the reference's own example workbooks (Spreadsheet Interface/Examples)
        carry the Edges header one row above where stryke.py:362-367 looks for it."""
        if sheet not in self._cells:
            raise ValueError(f"Worksheet named '{sheet}' not found")
        cells = self._cells[sheet]
        if not cells:
            return pd.DataFrame()
        last_row = max(r for r, _ in cells)
        last_col = max(c for _, c in cells)
        if usecols:
            a, b = usecols.split(":")
            cols = list(range(col_index(a), col_index(b) + 1))
        else:
            cols = list(range(0, last_col + 1))
        header_row = skiprows
        if expect:
            want = set(expect)
            def holds(r):
                return want <= {cells.get((r, c)) for c in cols}
            if not holds(header_row):
                near = [r for r in range(max(0, skiprows - 3), skiprows + 4) if holds(r)]
                if near:
                    header_row = min(near, key=lambda r: abs(r - skiprows))
        names, seen = [], {}
        for j, c in enumerate(cols):
            h = cells.get((header_row, c))
            h = f"Unnamed: {c}" if h is None else (h if isinstance(h, str) else str(h))
            if h in seen:
                seen[h] += 1
                h = f"{h}.{seen[h]}"
            else:
                seen[h] = 0
            names.append(h)
        rows = [[cells.get((r, c), np.nan) for c in cols] for r in range(header_row + 1, last_row + 1)]
        df = pd.DataFrame(rows, columns=names)
        for c in df.columns:      # pandas-style inference: all-numeric columns become int64 / float64
            col = df[c]
            if col.map(lambda x: isinstance(x, (int, float, np.integer, np.floating)) and not isinstance(x, bool)).all():
                if col.notna().all() and col.map(lambda x: isinstance(x, (int, np.integer))).all():
                    df[c] = col.astype(np.int64)
                else:
                    df[c] = col.astype(float)
        for c, t in (dtype or {}).items():
            if c in df.columns:
                if t is str:
                    df[c] = df[c].map(lambda x: x if (isinstance(x, float) and np.isnan(x)) else str(x))
                else:
                    df[c] = df[c].astype(t)
        return df


def _cell_xml(ref, v):
    if v is None or (isinstance(v, float) and np.isnan(v)) or v is pd.NaT:
        return ""
    if isinstance(v, (bool, np.bool_)):
        return f'<c r="{ref}" t="b"><v>{int(v)}</v></c>'
    if isinstance(v, (int, float, np.integer, np.floating)):
        return f'<c r="{ref}"><v>{repr(float(v)) if isinstance(v, (float, np.floating)) else int(v)}</v></c>'
    return f'<c r="{ref}" t="inlineStr"><is><t xml:space="preserve">{escape(str(v))}</t></is></c>'


def _sheet_xml(df, startrow=0, startcol=0, index=True):
    if index:
        df = df.reset_index()
    out = ['<?xml version="1.0" encoding="UTF-8" standalone="yes"?>',
           '<worksheet xmlns="http://schemas.openxmlformats.org/spreadsheetml/2006/main"><sheetData>']
    r0 = startrow + 1
    header = "".join(_cell_xml(f"{col_letters(startcol + j)}{r0}", "" if str(c) == "index" else str(c))
                     for j, c in enumerate(df.columns))
    out.append(f'<row r="{r0}">{header}</row>')
    for i, row in enumerate(df.itertuples(index=False), start=r0 + 1):
        out.append(f'<row r="{i}">' + "".join(_cell_xml(f"{col_letters(startcol + j)}{i}", v) for j, v in enumerate(row)) + "</row>")
    out.append("</sheetData></worksheet>")
    return "".join(out)


def new_workbook(path):
    """Write an empty workbook (no sheets yet) that ``append_sheets`` can extend."""
    with zipfile.ZipFile(path, "w", zipfile.ZIP_DEFLATED) as z:
        z.writestr("[Content_Types].xml",
                   '<?xml version="1.0" encoding="UTF-8" standalone="yes"?><Types xmlns="http://schemas.openxmlformats.org/package/2006/content-types">'
                   '<Default Extension="xml" ContentType="application/xml"/>'
                   '<Default Extension="rels" ContentType="application/vnd.openxmlformats-package.relationships+xml"/>'
                   '<Override PartName="/xl/workbook.xml" ContentType="application/vnd.openxmlformats-officedocument.spreadsheetml.sheet.main+xml"/></Types>')
        z.writestr("_rels/.rels",
                   '<?xml version="1.0" encoding="UTF-8" standalone="yes"?><Relationships xmlns="http://schemas.openxmlformats.org/package/2006/relationships">'
                   '<Relationship Id="rId1" Type="http://schemas.openxmlformats.org/officeDocument/2006/relationships/officeDocument" Target="xl/workbook.xml"/></Relationships>')
        z.writestr("xl/workbook.xml",
                   '<?xml version="1.0" encoding="UTF-8" standalone="yes"?><workbook xmlns="http://schemas.openxmlformats.org/spreadsheetml/2006/main" '
                   'xmlns:r="http://schemas.openxmlformats.org/officeDocument/2006/relationships"><sheets></sheets></workbook>')
        z.writestr("xl/_rels/workbook.xml.rels",
                   '<?xml version="1.0" encoding="UTF-8" standalone="yes"?><Relationships xmlns="http://schemas.openxmlformats.org/package/2006/relationships"></Relationships>')


def append_sheets(path, frames, startrow=0, startcol=0, index=True):
    """Add worksheets to an existing workbook.  ``startrow`` / ``startcol`` may be ints or {sheet: int} dicts."""
    with zipfile.ZipFile(path) as z:
        parts = {n: z.read(n) for n in z.namelist()}
    wb = parts["xl/workbook.xml"].decode("utf-8")
    rels = parts["xl/_rels/workbook.xml.rels"].decode("utf-8")
    ctypes = parts["[Content_Types].xml"].decode("utf-8")
    sheet_ids = [int(x) for x in re.findall(r'sheetId="(\d+)"', wb)]
    rel_ids = [int(x) for x in re.findall(r'Id="rId(\d+)"', rels)]
    files = [int(x) for x in re.findall(r"worksheets/sheet(\d+)\.xml", " ".join(parts))]
    for name, df in frames.items():
        if re.search(r'<sheet [^>]*name="%s"' % re.escape(escape(name)), wb):
            raise ValueError(f"Sheet '{name}' already exists and if_sheet_exists is set to 'error'.")
        sid, rid, fid = max(sheet_ids + [0]) + 1, max(rel_ids + [0]) + 1, max(files + [0]) + 1
        sheet_ids.append(sid); rel_ids.append(rid); files.append(fid)
        pick = lambda v: v.get(name, 0) if isinstance(v, dict) else v
        parts[f"xl/worksheets/sheet{fid}.xml"] = _sheet_xml(df, pick(startrow), pick(startcol), index).encode("utf-8")
        wb = wb.replace("</sheets>", f'<sheet name="{escape(name)}" sheetId="{sid}" r:id="rId{rid}"/></sheets>')
        rels = rels.replace("</Relationships>",
                            f'<Relationship Id="rId{rid}" Type="http://schemas.openxmlformats.org/officeDocument/2006/relationships/worksheet" '
                            f'Target="worksheets/sheet{fid}.xml"/></Relationships>')
        ctypes = ctypes.replace("</Types>",
                                f'<Override PartName="/xl/worksheets/sheet{fid}.xml" '
                                'ContentType="application/vnd.openxmlformats-officedocument.spreadsheetml.worksheet+xml"/></Types>')
    parts["xl/workbook.xml"], parts["xl/_rels/workbook.xml.rels"] = wb.encode("utf-8"), rels.encode("utf-8")
    parts["[Content_Types].xml"] = ctypes.encode("utf-8")
    tmp = tempfile.NamedTemporaryFile(delete=False, suffix=".xlsx").name
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
        for n, b in parts.items():
            z.writestr(n, b)
    shutil.move(tmp, path)
