"""Text formats of the reference, so that inputs / outputs / strategy files are exchanged with fplll unchanged
(SURVEY §8 f3, Appendix D):

  read_matrix / write_matrix   Matrix<T>::read / print, fplll/nr/matrix.cpp:136-203 — "[[a b c]\\n[d e f]]", rows of unequal
                               length are zero-padded on the right like the reference does (matrix.cpp:186-201)
  load_strategies_json         load_strategies_json, fplll/bkz_param.cpp:82-157 — strategies/default.json and friends ->
                               the dict form fplll_b200.bkz takes (block size -> preprocessing sizes, pruning vectors)
  GSODump                      BKZReduction::dump_gso, fplll/bkz.cpp:728-798 (BKZ_DUMP_GSO): a JSON list of
                               {"step", "loop", "time", "norms": [log r_ii ...]} records

Host-side and exact (Python integers): none of this touches the GPU."""
import json
import math
import re

import numpy as np


def read_matrix(path_or_text, dtype=object):
    """Returns a 2-d numpy array (dtype=object keeps arbitrary-precision entries; np.int64 raises OverflowError if one
    does not fit, which is the reference's convert<long, mpz_t> refusal, bkz.cpp:826)."""
    text = path_or_text
    if "[" not in text:
        text = open(path_or_text).read()
    start = text.index("[")
    rows, depth, cur = [], 0, None
    for m in re.finditer(r"\[|\]|-?\d+", text[start:]):
        tok = m.group(0)
        if tok == "[":
            depth += 1
            if depth == 2:
                cur = []
        elif tok == "]":
            depth -= 1
            if depth == 1 and cur is not None:
                rows.append(cur)
                cur = None
            elif depth == 0:
                break
        elif depth == 2:
            cur.append(int(tok))
        else:
            raise ValueError("fplll matrix format: number outside a row")
    width = max((len(r) for r in rows), default=0)
    out = np.zeros((len(rows), width), dtype=object)
    for i, r in enumerate(rows):
        out[i, : len(r)] = r  # short rows are zero-padded (matrix.cpp:186-201)
    if dtype is object:
        return out
    return np.array(out, dtype=dtype)


def write_matrix(path, b, regular=False):
    """Matrix<T>::print (matrix.cpp:136-163): '[[a b c]\n[d e f]]' — the compact mode operator<< uses by default —
    or, regular=True, MAT_PRINT_REGULAR's '[[a b c ]\n[d e f ]\n]'.  path=None returns the text."""
    b = np.asarray(b)
    pad = " " if (regular and b.shape[1] > 0) else ""
    lines = ["[" + " ".join(str(int(x)) for x in row) + pad + "]" for row in b]
    text = "[" + "\n".join(lines) + ("\n" if (regular and len(lines)) else "") + "]\n"
    if path is None:
        return text
    with open(path, "w") as f:
        f.write(text)
    return text


def load_strategies_json(path):
    """strategies/*.json -> {block_size: (preprocessing_block_sizes int32[], gh_factor f64[], expectation f64[],
    coefficients f64[n_prune, block_size])}, the table fplll_b200.BKZParam(strategies=...) takes.  Entries are
    [gh_factor, [coefficients...], expectation] as bkz_param.cpp:121-140 reads them."""
    out = {}
    for e in json.load(open(path)):
        bs = int(e["block_size"])
        pp = e.get("pruning_parameters", [])
        coef = np.array([p[1] for p in pp], dtype=np.float64).reshape(len(pp), bs if pp else 0)
        out[bs] = (np.array(e.get("preprocessing_block_sizes", []), dtype=np.int32),
                   np.array([p[0] for p in pp], dtype=np.float64), np.array([p[2] for p in pp], dtype=np.float64), coef)
    return out


class GSODump:
    """BKZ_DUMP_GSO: append(step, loop, time, r_mant, r_expo) writes one record of bkz.cpp:728-798 — the norms are
    log(r_ii) = log(mantissa) + expo * log 2 with 8 significant digits, the steps "Input", "End of BKZ loop", "Output"."""

    def __init__(self, path):
        self.path = path
        self.first = True

    def append(self, step, loop, time, r_mant, r_expo):
        norms = ", ".join("%.8g" % (math.log(float(m)) + int(e) * math.log(2.0)) for m, e in zip(r_mant, r_expo))
        with open(self.path, "w" if self.first else "a") as f:
            if self.first:
                f.write("[\n")
            f.write(" " * 8 + "{\n")
            f.write(" " * 16 + '"step": "%s",\n' % step)
            f.write(" " * 16 + '"loop": %d,\n' % loop)
            f.write(" " * 16 + '"time": %s,\n' % repr(float(time)))
            f.write(" " * 16 + '"norms": [%s]\n' % norms)
            f.write(" " * 8 + "}")
            f.write("\n]" if step == "Output" else ",\n")
        self.first = False
