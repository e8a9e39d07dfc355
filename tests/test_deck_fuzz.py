"""Deck parser against the reference's own GetNameList (lmp2sqt.c:2014-2110), called in the unmodified reference
library on randomly malformed decks: same values in every field and the same `ok` verdict.  Needs oracle/_ref (built
where /root/reference exists); a frozen sample of the same generator runs everywhere against recorded answers."""
import ctypes as C
import json
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

import pyoracle
from lmp2sqt_b200 import _lib
from lmp2sqt_b200._lib import SqtDeck

HERE = os.path.dirname(os.path.abspath(__file__))
FIELDS_R = ("kVal", "deltaT", "mass", "rVal")
FIELDS_I = ("limitCorrAv", "nCBuffer", "nCSpatial", "nCTime", "nCSkip")
FROZEN = os.path.join(HERE, "golden", "deck_fuzz.json")


def make_decks(n, seed=7):
    """Decks in and around the NameList syntax: shuffled keys, tabs / spaces / commas, integers written as reals and
    reals as integers, exponents, missing / duplicate / unknown keys, extra tokens, blank and comment lines, a missing
    terminator, a long line (the reference reads 80 characters at a time)."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        keys = list(FIELDS_R + FIELDS_I)
        rng.shuffle(keys)
        if i % 4 == 3:
            # a deck GetNameList calls ok (every key once, nothing else -- no header, no terminator), or one flaw away from it
            lines = ["%s\t%s " % (k, rng.choice(["0.5", "2", "1e-1"]) if k in FIELDS_R else rng.choice(["10", "3", "0"])) for k in keys]
            flaw = int(rng.integers(0, 5))
            if flaw == 1:
                lines.pop(int(rng.integers(0, len(lines))))
            elif flaw == 2:
                lines.append(lines[0])
            elif flaw == 3:
                lines[2] += " 7"
            elif flaw == 4:
                lines.append("----")
            out.append("\n".join(lines) + "\n")
            continue
        lines = ["NameList -- data"] if rng.random() < 0.7 else []
        for k in keys:
            r = rng.random()
            if r < 0.08 and k != "nCTime":
                continue                                           # missing key
            if k in FIELDS_R:
                v = rng.choice(["0.0981748", "1", "6.0e-2", ".2", "-0.5", "1e3", "0.200000 ", "abc", "3.5xyz"])
            else:
                v = rng.choice(["10", "150", "0", "7.9", "-3", "1e2", "12abc", "", "2147483648"])
            sep = rng.choice(["\t", "\t\t", " ", "  \t "])
            tail = rng.choice(["", " ", " \n".strip("\n"), " 99", ", 5", "\t# note"]) if rng.random() < 0.3 else " "
            lines.append("%s%s%s%s" % (k, sep, v, tail))
            if rng.random() < 0.05:
                lines.append("%s%s%s" % (k, sep, "42"))             # duplicate key
        if rng.random() < 0.15:
            lines.insert(int(rng.integers(0, len(lines) + 1)), "unknownKey 3")
        if rng.random() < 0.1:
            lines.insert(int(rng.integers(1, len(lines) + 1)), "")  # a blank line ends the reference's scan
        if rng.random() < 0.05:
            lines.insert(1, "kVal " + "0" * 90 + "1")              # longer than the 80-character fgets window
        if rng.random() < 0.8:
            lines.append("----")
        out.append("\n".join(lines) + ("\n" if rng.random() < 0.9 else ""))
    return out


def ours(text, tmp):
    p = os.path.join(tmp, "deck")
    with open(p, "w") as f:
        f.write(text)
    d = SqtDeck()
    ok = _lib.host_lib().sqt_deck_read(p.encode(), C.byref(d))
    return {"ok": ok, **{k: getattr(d, k) for k in FIELDS_R + FIELDS_I}}


def _reference_worker(job):
    """Runs in a fresh process (the reference keeps its deck in file-scope globals)."""
    with open(job, "rb") as f:
        decks, tmp = pickle.load(f)
    L = C.CDLL(pyoracle.REF_SO)
    L.GetNameList.restype = C.c_int
    name = (C.c_char * 100).in_dll(L, "inputFilename")
    res = []
    for i, text in enumerate(decks):
        p = os.path.join(tmp, "d%d" % i)
        with open(p, "w") as f:
            f.write(text)
        name.value = p.encode()
        for k in FIELDS_R:
            C.c_double.in_dll(L, k).value = -7.25                   # sentinels: a key the deck lacks keeps its old value
        for k in FIELDS_I:
            C.c_int.in_dll(L, k).value = -77
        ok = L.GetNameList()
        r = {"ok": ok}
        r.update({k: C.c_double.in_dll(L, k).value for k in FIELDS_R})
        r.update({k: C.c_int.in_dll(L, k).value for k in FIELDS_I})
        res.append(r)
    with open(job + ".out", "wb") as f:
        pickle.dump(res, f)


def reference(decks, tmp):
    job = os.path.join(tmp, "job.pkl")
    with open(job, "wb") as f:
        pickle.dump((decks, tmp), f)
    code = "import sys; sys.path[:0] = [%r, %r, %r]; import test_deck_fuzz as t; t._reference_worker(%r)" % (
        HERE, os.path.join(os.path.dirname(HERE), "oracle"), os.path.dirname(HERE), job)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    with open(job + ".out", "rb") as f:
        return pickle.load(f)


def _same(a, b):
    for k in FIELDS_R + FIELDS_I:
        # a key the deck lacks: the reference keeps whatever its global held (sentinel), ours its own default -- only
        # fields the deck sets are comparable; nCSkip defaults to 0 in both (lmp2sqt.c:2031)
        if b[k] in (-7.25, -77):
            continue
        if a[k] != b[k]:
            return False
    return (a["ok"] == 1) == (b["ok"] == 1)


@pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref not built")
def test_deck_parser_agrees_with_reference_getnamelist_on_random_decks(tmp_path):
    decks = make_decks(400)
    ref = reference(decks, str(tmp_path))
    bad = [(i, ours(t, str(tmp_path)), r) for i, (t, r) in enumerate(zip(decks, ref)) if not _same(ours(t, str(tmp_path)), r)]
    assert not bad, (len(bad), decks[bad[0][0]], bad[0][1], bad[0][2])


def test_deck_parser_on_frozen_sample(tmp_path):
    """The first 60 decks of the generator with the reference's recorded answers (tests/golden/deck_fuzz.json)."""
    rec = json.load(open(FROZEN))
    decks = make_decks(400)[:len(rec)]
    for t, r in zip(decks, rec):
        assert _same(ours(t, str(tmp_path)), r), (t, ours(t, str(tmp_path)), r)


if __name__ == "__main__":                                           # regenerate the frozen sample (needs oracle/_ref)
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        json.dump(reference(make_decks(400)[:60], td), open(FROZEN, "w"), indent=0)
