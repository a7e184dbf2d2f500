"""CPU-side tests of the host logic: the C-ABI library loads and exports every symbol the header
declares, the product's H expansion and edge tables match the oracle's, size rules, sharding, and
the no-CPU-fallback rule (compute entry points fail loudly without a device)."""
import ctypes as C
import re

import numpy as np
import pytest

import oracle


def test_library_exports_every_declared_symbol():
    from myldpccppapi_b200 import lib
    L = lib.load()
    names = lib.header_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), "libldpc_b200.so does not export " + n
    assert set(names) == set(lib.SIGNATURES), "lib.py signatures out of sync with include/ldpc_b200.h"
    assert b"sm_100a" in L.ldpc_b200_version()
    # the drop-in C++ Coder's C doorway (include/MyLdpc_c.h, libmyldpc_b200.so)
    LC = lib.load_coder()
    cnames = lib.coder_header_symbols()
    assert len(cnames) >= 20
    for n in cnames:
        assert hasattr(LC, n), "libmyldpc_b200.so does not export " + n
    assert set(cnames) == set(lib.CODER_SIGNATURES), "lib.py signatures out of sync with include/MyLdpc_c.h"


def test_library_is_sm_100a_only():
    """The shared object carries sm_100a SASS for the decode kernels and nothing else."""
    import shutil, subprocess
    from myldpccppapi_b200 import _build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([cuobjdump, "-lelf", str(_build.LIB)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.parametrize("rate,name,num,den", [(0, "1/2", 1, 2), (1, "2/3A", 2, 3), (2, "2/3B", 2, 3),
                                               (3, "3/4A", 3, 4), (4, "3/4B", 3, 4), (5, "5/6", 5, 6)])
@pytest.mark.parametrize("N", [576, 1056, 2304])
def test_wimax_H_matches_oracle(rate, name, num, den, N):
    import myldpccppapi_b200 as m
    K = N * num // den
    rp, ci, M = m.wimax_csr(K, N, rate)
    orp, oci, oM = oracle.wimax_H(N, name)
    assert M == oM and np.array_equal(rp, orp) and np.array_equal(ci, oci)


def test_wimax_argument_errors():
    import myldpccppapi_b200 as m
    with pytest.raises(m.LdpcError):
        m.wimax_csr(432, 577, 4)       # N not a multiple of 24
    with pytest.raises(m.LdpcError):
        m.wimax_csr(400, 576, 4)       # K inconsistent with the rate
    with pytest.raises(m.LdpcError):
        m.wimax_csr(432, 576, 9)       # no such rate


def test_edge_tables_are_in_ascending_row_order(default_code):
    """vn_edge lists every variable's edges in ascending edge id = ascending row: the order of the
    reference's hColFirstPtr/hColNextPtr walk, i.e. the fp32 summation order (MyLdpc.cpp:723-728)."""
    import myldpccppapi_b200 as m
    c = default_code
    cp, ve, rw, cw = m.edge_tables(c["M"], c["N"], c["row_ptr"], c["col_idx"])
    assert (rw, cw) == (15, 6) and cp[-1] == c["row_ptr"][-1]
    rows_of_edge = np.repeat(np.arange(c["M"]), np.diff(c["row_ptr"]))
    for n in range(c["N"]):
        ent = ve[cp[n]:cp[n + 1]]
        chk, pos = ent >> 5, ent & 31
        edges = c["row_ptr"][chk] + pos
        assert np.all(c["col_idx"][edges] == n)
        assert np.all(np.diff(edges) > 0) and np.all(np.diff(chk.astype(np.int64)) > 0)
        assert np.array_equal(rows_of_edge[edges], chk)


def test_malformed_matrices_are_rejected():
    import myldpccppapi_b200 as m
    rp = np.array([0, 2, 4], dtype=np.int32)
    with pytest.raises(m.LdpcError):
        m.edge_tables(2, 3, rp, np.array([0, 1, 1, 3], dtype=np.int32))   # column out of range
    with pytest.raises(m.LdpcError):
        m.edge_tables(2, 3, rp, np.array([0, 0, 1, 2], dtype=np.int32))   # duplicate entry in a row
    with pytest.raises(m.LdpcError):
        m.edge_tables(2, 3, np.array([0, 3, 2], dtype=np.int32), np.array([0, 1, 2], dtype=np.int32))


def test_no_cpu_fallback_without_a_device(default_code):
    """On a box without CUDA the decoder cannot be created: error code LDPC_B200_ERR_CUDA, no silent path."""
    import torch
    import myldpccppapi_b200 as m
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    c = default_code
    with pytest.raises(m.LdpcError) as e:
        m.Decoder(c["M"], c["N"], c["K"], c["row_ptr"], c["col_idx"])
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    with pytest.raises(m.LdpcError):
        m.Coder(c["K"], c["N"], m.rate_3_4_b).forDecoder(4)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under myldpccppapi_b200/ or include/ may name it."""
    import pathlib
    root = pathlib.Path(__file__).resolve().parents[1]
    for p in list((root / "myldpccppapi_b200").rglob("*")) + list((root / "include").rglob("*")):
        if p.suffix in (".py", ".cu", ".cuh", ".cpp", ".h"):
            text = p.read_text()
            assert not re.search(r"^\s*(import|from)\s+oracle\b", text, re.M), p
            assert "liboracle" not in text and "ldpc_oracle" not in text, p


def test_coder_size_helpers_match_reference_rules():
    import myldpccppapi_b200 as m
    c = m.Coder(432, 576, m.rate_3_4_b)
    L = oracle.lib()
    for n in (1, 53, 54, 55, 1000, 54 * 64):
        assert c.getCodeSize(n) == L.oracle_getCodeSize(432, n)
        assert c.getPostCodeLength(n) == L.oracle_getPostCodeLength(432, 576, n)
        assert c.getPriorCodeLength(n) == L.oracle_getPriorCodeLength(432, 576, n)


def test_synthetic_codes_are_well_formed():
    from myldpccppapi_b200 import codes
    M, N, K, rp, ci = codes.regular_code()
    assert (M, N, K) == (4096, 8192, 4096) and np.all(np.diff(rp) == 6) and np.all(np.bincount(ci, minlength=N) == 3)
    M, N, K, rp, ci = codes.ira_code()
    assert (M, N, K) == (32400, 64800, 32400) and rp[-1] == 226799
    u = np.random.default_rng(0).integers(0, 2, (2, K)).astype(np.uint8)
    assert codes.syndrome(M, rp, ci, codes.ira_encode(M, N, K, rp, ci, u)).sum() == 0
    # rows hold distinct ascending columns
    for r in (0, 1, 777, M - 1):
        cols = ci[rp[r]:rp[r + 1]]
        assert np.all(np.diff(cols) > 0)


def test_shard_ranges_cover_and_are_disjoint():
    from myldpccppapi_b200 import shard_ranges
    for ncw in (0, 1, 7, 64, 65536, 1000003):
        for world in (1, 2, 4, 8):
            for align in (1, 8):
                r = shard_ranges(ncw, world, align)
                assert r[0][0] == 0 and r[-1][1] == ncw
                assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
                assert all(b % align == 0 or b == ncw for b, _ in r)


def test_staging_copy_gives_memcpy_bytes(tmp_path):
    """stage_copy_nt (non-temporal staging copy of the pageable host-buffer path): tests/native/stage_copy_test.cpp built
    against the product source with g++ and run here -- every alignment, sizes around the vector steps, no stray byte."""
    import pathlib, shutil, subprocess
    root = pathlib.Path(__file__).resolve().parent.parent
    cxx = shutil.which("g++")
    assert cxx, "g++ missing"
    exe = tmp_path / "stage_copy_test"
    csrc = root / "myldpccppapi_b200" / "csrc"
    r = subprocess.run([cxx, "-O2", "-std=c++17", "-I", str(csrc), "-I", str(root / "include"), str(root / "tests" / "native" / "stage_copy_test.cpp"),
                        str(csrc / "ldpc_tables.cpp"), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "bad=0" in r.stdout, r.stdout + r.stderr
