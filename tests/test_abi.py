"""not-gpu: the C-ABI library loads without a GPU and exports every symbol include/*.h declares"""
import ctypes
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    syms = set()
    for hdr in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = open(hdr).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        for m in re.finditer(r"X264_CUDA_API\s+[^;(]*?\b(x264_\w+)\s*\(", text):
            syms.add(m.group(1))
    return syms


def test_library_exports_every_declared_symbol(pkg):
    lib = ctypes.CDLL(pkg.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 20
    missing = [s for s in sorted(syms) if not hasattr(lib, s)]
    assert not missing, missing


def test_struct_sizes_match_header(pkg):
    assert pkg.ME_JOB.itemsize == 84 and pkg.ME_RESULT.itemsize == 16 and pkg.ME_FINAL.itemsize == 16
    assert pkg.ME_MB_JOB.itemsize == 280 and pkg.MB_COEFFS.itemsize == 816 and pkg.RESID_JOB.itemsize == 8


def test_host_cost_table_matches_oracle(pkg, port):
    import numpy as np
    for qp in (0, 11, 12, 20, 26, 37, 51):
        assert np.array_equal(pkg.host_cost_mv(qp), port.cost_mv_table(qp))
