"""CPU tests: the C-ABI library loads and exports every symbol include/smeagol_b200.h declares (no compute calls
without a GPU), and the product path fails loudly without CUDA."""
import ctypes
import os
import re

import pytest

from smeagol_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "smeagol_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(smg_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_and_library_exports_the_same_symbols():
    syms = header_symbols()
    assert sorted(_lib.EXPORTS) == syms
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), "libsmeagol_b200.so does not export %s" % s
    assert _lib.load().smg_version() >= 100


def test_struct_sizes_match_header():
    import numpy as np

    from smeagol_b200 import device as dev

    assert dev.ROW_DTYPE.itemsize == 32 and dev.CHUNK_DTYPE.itemsize == 8 and dev.NEAR_DTYPE.itemsize == _lib.NEAR_BYTES
    assert np.dtype(dev.ROW_DTYPE).names == ("word_off", "g_off", "g_len", "n_avail", "n_own")


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import numpy as np
    import pandas as pd

    import smeagol_b200 as sb

    model = sb.models.PWMModel(pd.DataFrame({"Matrix_id": ["a"], "weights": [np.zeros((3, 4))]}))
    with pytest.raises(_lib.SmeagolB200Error):
        model.predict(np.array([[1, 2, 3]]))
    with pytest.raises(_lib.SmeagolB200Error):
        sb.scan.scan_sequences([sb.SeqRecord(sb.Seq("ACGT"), id="a", name="a")], model, 0.5)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "smeagol_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            txt = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), fn
