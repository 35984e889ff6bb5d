"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the
header declares, its structs match the ctypes mirrors, and -- with no GPU -- the product path
fails loudly instead of falling back to any CPU implementation."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

from conftest import ROOT

import irrotational_warp_lab_b200  # noqa: F401  (the shim for irrotational-warp-lab_b200/)
from irrotational_warp_lab_b200 import _capi, backend

HEADER = os.path.join(ROOT, "include", "iwb200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(iwb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _capi.load()
    declared = _declared_symbols()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), f"libiwb200.so does not export {name}"
    assert set(declared) == set(_capi.SYMBOLS), "ctypes table and header disagree"
    assert lib.iwb_version() == 100


def test_struct_layouts_match_the_header():
    """Compile a 20-line C program against include/iwb200.h and compare sizeof/offsetof with ctypes."""
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "iwb200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu\n", sizeof(iwb_params3d), sizeof(iwb_result), sizeof(iwb_params_axisym),
         sizeof(iwb_params_slice), sizeof(iwb_slice_result));
  printf("%zu %zu %zu %zu %zu\n", offsetof(iwb_params3d, x), offsetof(iwb_params3d, shell_r),
         offsetof(iwb_params3d, rho_out), offsetof(iwb_result, shell_cnt), offsetof(iwb_params_axisym, rho_out));
  printf("%d %d %d\n", IWB_MAX_SHELLS, IWB_FLUSH_PLANES, IWB_ROW_WIDTH);
  return 0;
}'''
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "t.c")
        open(src, "w").write(prog)
        exe = os.path.join(td, "t")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        out = subprocess.check_output([exe]).decode().split()
    sizes = [int(v) for v in out]
    assert sizes[:5] == [C.sizeof(_capi.Params3D), C.sizeof(_capi.Result), C.sizeof(_capi.ParamsAxisym),
                         C.sizeof(_capi.ParamsSlice), C.sizeof(_capi.SliceResultC)]
    assert sizes[5:10] == [_capi.Params3D.x.offset, _capi.Params3D.shell_r.offset, _capi.Params3D.rho_out.offset,
                           _capi.Result.shell_cnt.offset, _capi.ParamsAxisym.rho_out.offset]
    assert sizes[10:] == [_capi.MAX_SHELLS, _capi.FLUSH_PLANES, _capi.ROW_WIDTH]


def _no_gpu():
    return _capi.load().iwb_device_count() == 0


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful without a GPU")
def test_no_cpu_fallback_without_gpu():
    h = C.c_void_p()
    rc = _capi.load().iwb_create(0, C.byref(h))
    assert rc == _capi.ERR_NO_DEVICE and not h
    assert "no CPU fallback" in _capi.last_error()
    with pytest.raises(RuntimeError, match="no CUDA device"):
        backend.compute_energy_3d(rho=5.0, sigma=4.0, v=1.0, extent=66.0, n=40)
    with pytest.raises(RuntimeError):
        backend.compute_energy_axisymmetric(rho=5.0, sigma=4.0, v=1.0, extent=66.0, nx=30, ny=20)


def test_unknown_backend_is_a_value_error():
    with pytest.raises(ValueError, match="Unknown backend"):
        backend.resolve_backend("cupyx")
    with pytest.raises(ValueError, match="Unknown backend"):
        backend.compute_energy_3d(rho=5.0, sigma=4.0, v=1.0, extent=66.0, n=40, backend="numpy")


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package may import or dlopen it."""
    pkg = os.path.join(ROOT, "irrotational-warp-lab_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle_np" not in text and "oracle_c" not in text and "liboracle" not in text, f
    code = "import sys; sys.path.insert(0, %r); import irrotational_warp_lab_b200.backend, irrotational_warp_lab_b200.sweeps, " \
           "irrotational_warp_lab_b200.optimize, irrotational_warp_lab_b200.distributed; " \
           "assert not any(m.startswith('oracle') for m in sys.modules), 'oracle imported'" % ROOT
    subprocess.check_call([sys.executable, "-c", code])


def test_span_table_covers_every_unit_once():
    """make_spans (csrc/energy3d.cu) through its host-only view: boundaries start at 0, end at `units`, increase
    strictly, the first `ctas` spans carry the static share in equal parts, the last span is a single unit (it bounds
    the scheduling tail), and a tiny problem degenerates to one unit per span."""
    lib = _capi.load()
    for units, ctas, pct in [(33408, 148, 70), (31552, 148, 70), (3968, 148, 70), (640, 148, 70), (12, 148, 70),
                             (1, 148, 70), (0, 148, 70), (5000, 148, 0), (5000, 148, 100), (5000, 1, 70), (1 << 20, 148, 80)]:
        n = lib.iwb_debug_spans(units, ctas, pct, None, 0)
        buf = (C.c_int * max(n, 1))()
        assert lib.iwb_debug_spans(units, ctas, pct, buf, n) == n
        b = list(buf)[:n]
        assert b[0] == 0 and b[-1] == units, (units, ctas, pct)
        assert all(b[q] < b[q + 1] for q in range(n - 1)), (units, ctas, pct)
        sizes = [b[q + 1] - b[q] for q in range(n - 1)]
        if units >= 4 * ctas and 0 < pct:
            static = sizes[:ctas]
            assert sum(static) == units * pct // 100 and max(static) - min(static) <= 1
            assert all(sizes[q] >= sizes[q + 1] for q in range(ctas, len(sizes) - 1))     # guided: never growing
        if units > 0 and pct < 100:
            assert sizes[-1] == 1
        if 0 < units < 4 * ctas:
            assert sizes == [1] * units


def test_column_tiles_partition_the_axis():
    """tiles_k / tile_begin_k (csrc/energy3d.cuh): interior tiles hold at most 60 output columns, the two tiles at
    the global faces 62, a single tile 64; the tiles partition [0, n) and there is never one more than needed."""
    lib = _capi.load()
    buf = (C.c_int * 256)()
    for n in list(range(3, 700)) + [1000, 1023, 1024, 1025, 2048, 4096, 5000]:
        nt = lib.iwb_debug_tiles_k(n, buf, 256)
        b = list(buf)[: nt + 1]
        assert b[0] == 0 and b[-1] == n
        sizes = [b[t + 1] - b[t] for t in range(nt)]
        assert min(sizes) >= 1
        if nt == 1:
            assert n <= 64
        else:
            assert sizes[0] <= 62 and sizes[-1] <= 62 and all(s <= 60 for s in sizes[1:-1])
            assert max(sizes) - min(sizes) <= 3            # the slack is spread evenly
            cap_fewer = 64 if nt == 2 else 2 * 62 + (nt - 3) * 60
            assert n > cap_fewer                           # one tile fewer would not hold n columns
    assert lib.iwb_debug_tiles_k(1024, buf, 256) == 17 and list(buf)[:3] == [0, 62, 122]


def test_row_tiles_partition_the_axis():
    """tiles_j / tile_begin_j (csrc/energy3d.cuh) for the default 40-row region: an even split into tiles of at most 36
    output rows (tiles of 38 rows at the j faces were measured twice and rejected)."""
    lib = _capi.load()
    buf = (C.c_int * 512)()
    for n in list(range(3, 500)) + [512, 1000, 1023, 1024, 1025, 2048, 4096]:
        nt = lib.iwb_debug_tiles_j(n, buf, 512)
        b = list(buf)[: nt + 1]
        assert b[0] == 0 and b[-1] == n
        sizes = [b[t + 1] - b[t] for t in range(nt)]
        assert min(sizes) >= 1 and max(sizes) <= 36
        assert nt == -(-n // 36)                       # no tile row more than needed
        assert max(sizes) - min(sizes) <= 1            # an even split
    assert lib.iwb_debug_tiles_j(256, buf, 512) == 8 and lib.iwb_debug_tiles_j(1024, buf, 512) == 29
