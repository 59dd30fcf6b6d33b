"""Helpers shared by the CPU (oracle) and GPU (parity) tests: golden loading and DataFrame comparison."""
import json
import os

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_PATH = os.path.join(HERE, "golden", "reference_golden.json")
_G = None


def golden():
    global _G
    if _G is None:
        with open(GOLDEN_PATH) as f:
            _G = json.load(f)
    return _G


def pwms_from_json(j):
    dt = np.dtype(j["dtype"])
    return list(j["Matrix_id"]), [np.array(w, dtype=np.float64).astype(dt) for w in j["weights"]]


def pwm_df(j):
    ids, ws = pwms_from_json(j)
    return pd.DataFrame({"Matrix_id": ids, "weights": ws})


def case_kwargs(case):
    kw = dict(case["kwargs"])
    thr = case["threshold"]
    if isinstance(thr, list):
        thr = np.array(thr)
    return thr, kw


SCORE_RTOL = 1e-5  # north_star: scores within 1e-5 relative


def assert_frame_matches(got, exp, name, float_cols=("score", "frac_score", "max_score", "avg", "sd"),
                         rtol=SCORE_RTOL, atol=0.0):
    """`exp` is a dict of columns (golden JSON) or a DataFrame; order and values must match."""
    if isinstance(exp, pd.DataFrame):
        exp = {c: exp[c].to_numpy() for c in exp.columns}
    assert list(got.columns) == list(exp.keys()), "%s: columns %s != %s" % (name, list(got.columns), list(exp.keys()))
    n = len(next(iter(exp.values()))) if exp else 0
    assert len(got) == n, "%s: %d rows, expected %d" % (name, len(got), n)
    for c, ev in exp.items():
        gv = got[c].to_numpy()
        if c in float_cols:
            np.testing.assert_allclose(gv.astype(np.float64), np.asarray(ev, dtype=np.float64), rtol=rtol, atol=atol,
                                       err_msg="%s: column %s" % (name, c))
        elif c == "bin":
            assert [str(x) for x in gv] == [str(x) for x in ev], "%s: column bin" % name
        else:
            assert list(gv) == list(ev), "%s: column %s differs" % (name, c)


def fix_stats_ddof(key, exp):
    """The golden 'stats' frames were recorded from the reference running under pandas 3, where
    ``.agg([len, np.mean, np.std])`` (scan.py:304) yields ddof=0; the reference's own golden
    tests/data/shuf_stats.csv (pandas < 2) is ddof=1, which is what this framework pins.  Convert."""
    if key != "stats":
        return exp
    exp = dict(exp)
    n = np.asarray(exp["len"], dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        exp["sd"] = list(np.asarray(exp["sd"], dtype=np.float64) * np.sqrt(n / (n - 1)))
    return exp
