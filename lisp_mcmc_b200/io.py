"""Input side of the walker API: the reference's file readers restated for the flat FP64 staging the C ABI
takes -- read-file->data (mcmc-fitting.lisp:1452-1477), file->file-specs (:1426-1441), create-walker-data (:827-831).

The one subtlety is the Common Lisp reader: `read-from-string` parses `2.003E+3` as a SINGLE-float (`E`, or no exponent
marker, means *read-default-float-format* = single-float; only `D` makes a double), integers and ratios stay exact, and
walker-create later widens everything with to-double-floats (:833-835).  A value such as -4.172E-7 therefore reaches
the likelihood as the double nearest to the float32 nearest to it -- 1e-8 relative away from what a C parser gives.
`cl_read_number` reproduces that, so a fit started from a file sees the same numbers as the reference."""
import re

import numpy as np

_INT = re.compile(r"^[+-]?\d+\.?$")
_RATIO = re.compile(r"^([+-]?\d+)/(\d+)$")
_FLOAT = re.compile(r"^[+-]?(\d+\.?\d*|\.\d+)(([eEsSfFdDlL])([+-]?\d+))?$")


def cl_read_number(token):
    """What (read-from-string token) yields, already passed through to-double-floats: float64, or None for a token that
    does not read as a number (a symbol, a string, NIL)."""
    # read-from-string returns the FIRST object: the token ends at whitespace or a terminating macro character
    # (`;` starts a comment, so "1;1.5" reads as 1 -- which is how file->file-specs recognises data lines of files
    # that are not tab-separated)
    t = re.split(r"[\s;()\"'`,]", token.strip(), maxsplit=1)[0]
    if not t:
        return None
    if _INT.match(t):
        return float(int(t.rstrip(".")))
    m = _RATIO.match(t)
    if m:
        return float(int(m.group(1))) / float(int(m.group(2))) if int(m.group(2)) else None
    m = _FLOAT.match(t)
    if not m:
        return None
    marker = (m.group(3) or "e").lower()
    text = t if m.group(3) is None else t[:m.start(3)] + "e" + m.group(4)
    v = float(text)
    if marker in ("d", "l"):                 # double-float / long-float (= double in SBCL)
        return v
    return float(np.float32(v))              # e / s / f: single-float, widened exactly


def _split(delim, line):
    return line.split(delim)


def file_specs(filename, delim="\t"):
    """file->file-specs :1426-1441: counts of lines, header lines (everything before the first line whose first
    tab-separated token reads as a number), columns, rows per page (pages are separated by an empty line) and pages."""
    num_lines, found_data, data_length, data_rows = 0, None, None, None
    with open(filename, "r", newline="") as f:
        for raw in f:
            line = raw.rstrip("\n").rstrip("\r")
            if line == "":
                if found_data is not None and data_rows is None:
                    data_rows = num_lines - found_data
                continue
            if found_data is None and cl_read_number(_split("\t", line)[0]) is not None:
                found_data = num_lines
                data_length = len(_split(delim, line))
            num_lines += 1
    if found_data is None:
        raise ValueError("%s holds no line that starts with a number" % filename)
    rows = data_rows if data_rows else num_lines - found_data
    return {"file-lines": num_lines, "header-lines": found_data, "data-length": data_length, "data-rows": rows,
            "num-pages": (num_lines - found_data) // data_rows if data_rows else 1}


def read_file_to_data(filename, specs=None, delim="\t", transpose=True, pages=False):
    """read-file->data :1452-1477 -> list of columns (transpose=True, the default) or of rows, each a float64 array;
    with pages=True a list of such lists, one per page.  Tokens that are not numbers become NaN."""
    specs = specs or file_specs(filename, delim)
    rows = []
    with open(filename, "r", newline="") as f:
        skipped = 0
        for raw in f:
            if skipped < specs["header-lines"]:
                skipped += 1
                continue
            line = raw.rstrip("\n").rstrip("\r")
            if line == "":
                continue
            vals = [cl_read_number(tok) for tok in _split(delim, line)]
            rows.append([np.nan if v is None else v for v in vals])
    width = max(len(r) for r in rows)
    table = np.array([r + [np.nan] * (width - len(r)) for r in rows], dtype=np.float64)

    def shape(block):
        return [np.ascontiguousarray(c) for c in (block.T if transpose else block)]
    if not pages:
        return shape(table)
    n, per = specs["num-pages"], specs["data-rows"]
    return [shape(table[k * per:(k + 1) * per]) for k in range(n)]


def create_walker_data(data, *columns):
    """create-walker-data :827-831: pick the columns a walker needs, e.g. create_walker_data(data, 0, 3) -> [x, y]"""
    return [np.asarray(data[c], dtype=np.float64) for c in columns]
