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


def test_headers_compile_as_c99_and_sizes_agree(pkg, tmp_path):
    """the public headers are plain C (the reference is C99): compile them with gcc -std=c99 -pedantic-errors and print the sizes of the
    records the Python binding mirrors"""
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include "x264_cuda.h"\n#include "x264_cuda_tables.h"\n'
                   'int main(void){ printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(x264_cuda_me_job_t), sizeof(x264_cuda_me_result_t), '
                   'sizeof(x264_cuda_me_mb_job_t), sizeof(x264_cuda_me_mb_result_t), sizeof(x264_cuda_resid_job_t), sizeof(x264_cuda_mb_coeffs_t), '
                   'sizeof(x264_cuda_intra16_job_t), sizeof(x264_cuda_mb_coeffs_i16_t), sizeof(x264_cuda_skip_job_t), sizeof(x264_cuda_lowres_result_t)); return 0; }\n')
    exe = tmp_path / "abi"
    r = subprocess.run(["gcc", "-std=c99", "-pedantic-errors", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    assert sizes == [pkg.ME_JOB.itemsize, pkg.ME_RESULT.itemsize, pkg.ME_MB_JOB.itemsize, pkg.ME_MB_RESULT.itemsize, pkg.RESID_JOB.itemsize,
                     pkg.MB_COEFFS.itemsize, pkg.INTRA16_JOB.itemsize, pkg.MB_COEFFS_I16.itemsize, pkg.SKIP_JOB.itemsize, 16], sizes
