"""Process-wide engine handle (one CUDA device per process) and thin typed wrappers.

Replaces the role of ``_resolve_backend`` (/root/reference/scripts/reproduce_rodal_exact.py:16-30)
for ``backend="b200"``: instead of returning an array module it hands out the handle the
fused evaluators run on.  Creation fails loudly (RuntimeError) without a usable GPU.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

from . import _capi

_lock = threading.Lock()
_engines: dict[int, "Engine"] = {}


def _as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Engine:
    """Owns one ``iwb_handle`` (streams, scratch) on one device."""

    def __init__(self, device: int = 0):
        self._lib = _capi.load()
        h = C.c_void_p()
        _capi.check(self._lib.iwb_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)
        sm, major, minor = C.c_int(), C.c_int(), C.c_int()
        name = C.create_string_buffer(128)
        _capi.check(self._lib.iwb_device_info(h, C.byref(sm), C.byref(major), C.byref(minor), name, 128))
        self.sm_count = sm.value
        self.cc = (major.value, minor.value)
        self.name = name.value.decode()

    def close(self):
        if self._h:
            self._lib.iwb_destroy(self._h)
            self._h = None

    # -- 3D ---------------------------------------------------------------------------------
    @staticmethod
    def make_params3d(*, x, y, z, dx, dy, dz, rho, sigma, v, dtype, shells=(), eps=1e-12,
                      i_begin=0, i_end=None, rho_out=None, rho_out_is_device=False, mode="exact",
                      tile_stride=1, tile_first=0):
        """Fill an ``iwb_params3d``.  ``x, y, z`` are the host linspace arrays (float32 grids are
        widened to float64 without change of value).  Returns (params, keepalive)."""
        xd, yd, zd = _as_f64(x), _as_f64(y), _as_f64(z)
        shells = [float(s) for s in shells]
        if len(shells) > _capi.MAX_SHELLS:
            raise ValueError(f"at most {_capi.MAX_SHELLS} shell radii per call")
        p = _capi.Params3D()
        if dtype not in ("float64", "float32"):
            raise ValueError(f"Unknown dtype: {dtype!r} (expected: float64|float32)")
        # "float32" is the reference's --dtype float32 as NumPy >= 2 evaluates it (mixed precision), bit for bit
        p.dtype = _capi.F32_REF if dtype == "float32" else _capi.F64
        if mode not in ("exact", "fast"):
            raise ValueError(f"Unknown mode: {mode!r} (expected: exact|fast)")
        p.mode = _capi.MODE_FAST if mode == "fast" else _capi.MODE_EXACT
        p.n0, p.n1, p.n2 = len(xd), len(yd), len(zd)
        p.i_begin = int(i_begin)
        p.i_end = len(xd) if i_end is None else int(i_end)
        p.x, p.y, p.z = _ptr(xd), _ptr(yd), _ptr(zd)
        p.dx, p.dy, p.dz = float(dx), float(dy), float(dz)
        p.rho, p.sigma, p.v, p.eps = float(rho), float(sigma), float(v), float(eps)
        rs = rho * sigma                       # potential.py:50-52: np.sinh / np.cosh of a Python float
        p.sinh_rs, p.cosh_rs = float(np.sinh(rs)), float(np.cosh(rs))
        p.nshell = len(shells)
        for s, r in enumerate(shells):
            p.shell_r[s] = r
        p.tile_stride, p.tile_first = int(tile_stride), int(tile_first)     # interleaved tile shard (multi-GPU)
        keep = [xd, yd, zd]
        if rho_out is not None:
            if rho_out_is_device:
                p.rho_out = int(rho_out)
                p.rho_out_is_device = 1
            else:
                assert rho_out.dtype == np.float64 and rho_out.flags.c_contiguous
                p.rho_out = rho_out.ctypes.data
                keep.append(rho_out)
        return p, keep

    @staticmethod
    def _result_dict(r, nshell):
        return {
            "e_pos": r.e_pos, "e_neg": r.e_neg,
            "cnt_pos": r.cnt_pos, "cnt_neg": r.cnt_neg, "cnt_zero": r.cnt_zero, "cnt_nan": r.cnt_nan,
            "shell_pos": [r.shell_pos[s] for s in range(nshell)],
            "shell_neg": [r.shell_neg[s] for s in range(nshell)],
            "shell_cnt": [r.shell_cnt[s] for s in range(nshell)],
            "kernel_ms": r.kernel_ms, "total_ms": r.total_ms,
        }

    def energy3d(self, **kw):
        p, keep = self.make_params3d(**kw)
        r = _capi.Result()
        _capi.check(self._lib.iwb_energy3d(self._h, C.byref(p), C.byref(r)))
        del keep
        return self._result_dict(r, p.nshell)

    def radial_profile3d(self, edges, **kw):
        """Per-bin sum and count of rho_adm over ``edges[k] <= r < edges[k+1]`` plus the energies of the same
        evaluation (iwb_radial_profile3d)."""
        edges = np.ascontiguousarray(edges, dtype=np.float64)
        nb = len(edges) - 1
        p, keep = self.make_params3d(**kw)
        sums = np.zeros(max(nb, 1), dtype=np.float64)
        counts = np.zeros(max(nb, 1), dtype=np.int64)
        prof = _capi.Profile3D()
        prof.n_bins = nb
        prof.edges = edges.ctypes.data_as(C.POINTER(C.c_double))
        prof.sum = sums.ctypes.data_as(C.POINTER(C.c_double))
        prof.count = counts.ctypes.data_as(C.POINTER(C.c_int64))
        r = _capi.Result()
        _capi.check(self._lib.iwb_radial_profile3d(self._h, C.byref(p), C.byref(prof), C.byref(r)))
        del keep
        out = self._result_dict(r, p.nshell)
        out["bin_sum"], out["bin_count"] = sums[:nb], counts[:nb]
        out["path"] = {1: "fused", 2: "two-pass"}.get(self._lib.iwb_debug_profile_path(self._h), "none")
        return out

    def energy3d_batch(self, points):
        """``points``: list of kwargs for :meth:`make_params3d`; one host sync for the whole batch."""
        n = len(points)
        arr = (_capi.Params3D * n)()
        keep = []
        for q, kw in enumerate(points):
            p, k = self.make_params3d(**kw)
            arr[q] = p
            keep.append(k)
        res = (_capi.Result * n)()
        _capi.check(self._lib.iwb_energy3d_batch(self._h, n, arr, res))
        del keep
        return [self._result_dict(res[q], arr[q].nshell) for q in range(n)]

    def energy3d_slab_async(self, table_ptr, stream_ptr=None, **kw):
        p, keep = self.make_params3d(**kw)
        _capi.check(self._lib.iwb_energy3d_slab_async(self._h, C.byref(p), C.c_void_p(int(table_ptr)),
                                                      C.c_void_p(int(stream_ptr or 0))))
        return keep

    def energy3d_tiles_async(self, chains_ptr, stream_ptr=None, **kw):
        """Interleaved tile shard ``tile_first`` of ``tile_stride``: chain sums into the device buffer
        ``[nfb, TILE_CHAINS // tile_stride, ROW_WIDTH]`` (iwb_energy3d_tiles_async); no synchronisation."""
        p, keep = self.make_params3d(**kw)
        _capi.check(self._lib.iwb_energy3d_tiles_async(self._h, C.byref(p), C.c_void_p(int(chains_ptr)),
                                                       C.c_void_p(int(stream_ptr or 0))))
        return keep

    def reduce_chains(self, gathered_ptr, tile_stride, nfb, nshell, stream_ptr=None):
        """Fixed tree over the gathered chain sums ``[tile_stride, nfb, TILE_CHAINS // tile_stride, ROW_WIDTH]`` and the
        row sum (iwb_reduce_chains); blocking."""
        r = _capi.Result()
        _capi.check(self._lib.iwb_reduce_chains(self._h, C.c_void_p(int(gathered_ptr)), int(tile_stride), int(nfb),
                                                int(nshell), C.c_void_p(int(stream_ptr or 0)), C.byref(r)))
        return self._result_dict(r, nshell)

    def reduce_table_async(self, table_ptr, nrows, row_ptr, stream_ptr=None):
        """k_reduce_rows into the device row ``row_ptr`` (ROW_WIDTH doubles); no copy, no synchronisation."""
        _capi.check(self._lib.iwb_reduce_table_async(self._h, C.c_void_p(int(table_ptr)), int(nrows),
                                                     C.c_void_p(int(stream_ptr or 0)), C.c_void_p(int(row_ptr))))

    def reduce_chains_async(self, gathered_ptr, tile_stride, nfb, scratch_ptr, row_ptr, stream_ptr=None):
        """Chain tree + row sum into the device row ``row_ptr``; ``scratch_ptr``: [nfb, ROW_WIDTH] doubles; no synchronisation."""
        _capi.check(self._lib.iwb_reduce_chains_async(self._h, C.c_void_p(int(gathered_ptr)), int(tile_stride), int(nfb),
                                                      C.c_void_p(int(stream_ptr or 0)), C.c_void_p(int(scratch_ptr)),
                                                      C.c_void_p(int(row_ptr))))

    def read_row(self, row_ptr, nshell, stream_ptr=None):
        """Blocking copy of a device result row (ROW_WIDTH doubles) to the host, unpacked (iwb_read_row)."""
        r = _capi.Result()
        _capi.check(self._lib.iwb_read_row(self._h, C.c_void_p(int(row_ptr)), int(nshell), C.c_void_p(int(stream_ptr or 0)),
                                           C.byref(r)))
        return self._result_dict(r, nshell)

    def peer_group(self, rank, world, nfb_max, exchange_handles):
        """Peer-memory exchange group of ``world`` single-GPU processes on one node (see :class:`PeerGroup`)."""
        return PeerGroup(self, rank, world, nfb_max, exchange_handles)

    def plan3d(self, **kw):
        """Prepared evaluation (coordinates resident on the device); see :class:`Plan3D`."""
        return Plan3D(self, **kw)

    def reduce_table(self, table_ptr, nrows, nshell, stream_ptr=None):
        r = _capi.Result()
        _capi.check(self._lib.iwb_reduce_table(self._h, C.c_void_p(int(table_ptr)), int(nrows), int(nshell),
                                               C.c_void_p(int(stream_ptr or 0)), C.byref(r)))
        return self._result_dict(r, nshell)

    # -- axisymmetric ---------------------------------------------------------------------------
    def energy_axisym(self, *, x, y, dx, dy, rho, sigma, v, dtype, shells=(), eps=1e-12, rho_out=None):
        xd, yd = _as_f64(x), _as_f64(y)
        shells = [float(s) for s in shells]
        if len(shells) > _capi.MAX_SHELLS:
            raise ValueError(f"at most {_capi.MAX_SHELLS} shell radii per call")
        if dtype not in ("float64", "float32"):
            raise ValueError(f"Unknown dtype: {dtype!r} (expected: float64|float32)")
        p = _capi.ParamsAxisym()
        p.dtype = _capi.F32_REF if dtype == "float32" else _capi.F64
        p.mode = _capi.MODE_EXACT
        p.nx, p.ny = len(xd), len(yd)
        p.x, p.y = _ptr(xd), _ptr(yd)
        p.dx, p.dy = float(dx), float(dy)
        p.rho, p.sigma, p.v, p.eps = float(rho), float(sigma), float(v), float(eps)
        rs = rho * sigma
        p.sinh_rs, p.cosh_rs = float(np.sinh(rs)), float(np.cosh(rs))
        p.nshell = len(shells)
        for s, r in enumerate(shells):
            p.shell_r[s] = r
        if rho_out is not None:
            assert rho_out.dtype == np.float64 and rho_out.flags.c_contiguous and rho_out.shape == (len(yd), len(xd))
            p.rho_out = _ptr(rho_out)
        r = _capi.Result()
        _capi.check(self._lib.iwb_energy_axisym(self._h, C.byref(p), C.byref(r)))
        return self._result_dict(r, p.nshell)

    # -- 2D slice -----------------------------------------------------------------------------------
    def slice_z0(self, *, x, y, dx, dy, rho, sigma, v, eps=1e-12):
        xd, yd = _as_f64(x), _as_f64(y)
        n = len(xd)
        if len(yd) != n:
            raise ValueError("slice_z0 needs a square grid")
        fields = {k: np.empty((n, n), dtype=np.float64) for k in ("phi", "beta_x", "beta_y", "rho_adm")}
        p = _capi.ParamsSlice()
        p.n = n
        p.x, p.y = _ptr(xd), _ptr(yd)
        p.dx, p.dy = float(dx), float(dy)
        p.rho, p.sigma, p.v, p.eps = float(rho), float(sigma), float(v), float(eps)
        p.phi, p.beta_x, p.beta_y, p.rho_adm = (_ptr(fields[k]) for k in ("phi", "beta_x", "beta_y", "rho_adm"))
        r = _capi.SliceResultC()
        _capi.check(self._lib.iwb_slice_z0(self._h, C.byref(p), C.byref(r)))
        fields.update(pos=r.pos, neg=r.neg, net=r.net, cnt_pos=r.cnt_pos, cnt_neg=r.cnt_neg,
                      cnt_zero=r.cnt_zero, kernel_ms=r.kernel_ms, total_ms=r.total_ms)
        return fields

    # -- diagnostics ----------------------------------------------------------------------------
    def measure_fma_peak(self, fp32=False):
        t, mhz = C.c_double(), C.c_double()
        _capi.check(self._lib.iwb_measure_fma_peak(self._h, 1 if fp32 else 0, C.byref(t), C.byref(mhz)))
        return {"tflops": t.value, "sm_mhz_effective": mhz.value}

    def selftest_divsqrt(self, count=1 << 26, seed=1, wide=False):
        bd, bs, sk = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _capi.check(self._lib.iwb_selftest_divsqrt(self._h, int(count), int(seed), 1 if wide else 0,
                                                   C.byref(bd), C.byref(bs), C.byref(sk)))
        return {"div_mismatches": bd.value, "sqrt_mismatches": bs.value, "skipped": sk.value}

    def selftest_phi(self, *, rho=5.0, sigma=4.0, v=1.0, eps=1e-12, coord_lo=0.05, coord_hi=66.0, count=1 << 26, seed=1):
        bad, wall = C.c_uint64(), C.c_uint64()
        _capi.check(self._lib.iwb_selftest_phi(self._h, float(rho), float(sigma), float(v), float(eps), float(coord_lo),
                                               float(coord_hi), int(count), int(seed), C.byref(bad), C.byref(wall)))
        return {"mismatches": bad.value, "wall_points": wall.value}

    def selftest_constdiv(self, divisor, count=1 << 26, seed=1):
        bad = C.c_uint64()
        _capi.check(self._lib.iwb_selftest_constdiv(self._h, float(divisor), int(count), int(seed), C.byref(bad)))
        return bad.value


class Plan3D:
    """``iwb_plan3d``: upload once, launch many times without host<->device copies."""

    def __init__(self, eng: Engine, **kw):
        self._eng = eng
        p, self._keep = Engine.make_params3d(**kw)
        self.nshell = p.nshell
        self.n0, self.i_begin, self.i_end = p.n0, p.i_begin, p.i_end
        h = C.c_void_p()
        _capi.check(eng._lib.iwb_plan3d_create(eng._h, C.byref(p), C.byref(h)))
        self._h = h

    def run_async(self, table_ptr, stream_ptr=None):
        _capi.check(self._eng._lib.iwb_plan3d_run_async(self._h, C.c_void_p(int(table_ptr)), C.c_void_p(int(stream_ptr or 0))))

    def run_exchange_async(self, peers: "PeerGroup", row_ptr, stream_ptr=None):
        """Collective: this rank's prepared tile shard with the fused peer exchange; result row in ``row_ptr``."""
        _capi.check(self._eng._lib.iwb_plan3d_run_exchange_async(self._h, peers._h, C.c_void_p(int(stream_ptr or 0)),
                                                                  C.c_void_p(int(row_ptr))))

    def close(self):
        if self._h:
            self._eng._lib.iwb_plan3d_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PeerGroup:
    """``iwb_peer``: the buffers through which the ranks of a tile-sharded evaluation exchange their chain sums as NVLink
    peer stores (no NCCL launch).  ``exchange_handles(blob) -> list[bytes]`` is the host collective that returns every
    rank's 64-byte IPC handle in rank order (e.g. ``torch.distributed.all_gather_object``); creation is collective."""

    def __init__(self, eng: Engine, rank: int, world: int, nfb_max: int, exchange_handles):
        self._eng = eng
        self.rank, self.world, self.nfb_max = int(rank), int(world), int(nfb_max)
        h = C.c_void_p()
        blob = C.create_string_buffer(_capi.PEER_HANDLE_BYTES)
        _capi.check(eng._lib.iwb_peer_create(eng._h, self.rank, self.world, self.nfb_max, C.byref(h), blob))
        self._h = h
        if exchange_handles is None:          # emulated ranks of one process: connect_local() follows (tests)
            return
        handles = exchange_handles(bytes(blob.raw))
        if len(handles) != self.world or any(len(b) != _capi.PEER_HANDLE_BYTES for b in handles):
            raise RuntimeError("peer exchange: the handle collective returned a malformed list")
        joined = C.create_string_buffer(b"".join(handles), self.world * _capi.PEER_HANDLE_BYTES)
        _capi.check(eng._lib.iwb_peer_connect(self._h, joined))

    @staticmethod
    def connect_local(groups):
        """Test hook: connect the emulated ranks ``groups`` (all created in this process with ``exchange_handles=None``)."""
        arr = (C.c_void_p * len(groups))(*[g._h for g in groups])
        for g in groups:
            _capi.check(g._eng._lib.iwb_debug_peer_connect_local(g._h, arr))

    def energy3d_tiles_exchange_async(self, row_ptr, stream_ptr=None, **kw):
        """Collective: kernel + peer push of the chain sums + combine; the result row lands in ``row_ptr`` (device)."""
        p, keep = self.prepare(**kw)
        self.run_prepared(p, row_ptr, stream_ptr)
        return keep

    def prepare(self, **kw):
        """This rank's ``iwb_params3d`` of a tile-sharded evaluation (and the arrays it points into): a caller that repeats
        an evaluation keeps the pair and calls :meth:`run_prepared`."""
        return Engine.make_params3d(tile_stride=self.world, tile_first=self.rank, **kw)

    def run_prepared(self, p, row_ptr, stream_ptr=None):
        _capi.check(self._eng._lib.iwb_energy3d_tiles_exchange_async(self._eng._h, self._h, C.byref(p),
                                                                     C.c_void_p(int(stream_ptr or 0)), C.c_void_p(int(row_ptr))))

    def close(self):
        if self._h:
            self._eng._lib.iwb_peer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def get_engine(device: int | None = None) -> Engine:
    """The engine of this process (device = LOCAL_RANK under torchrun, else 0)."""
    if device is None:
        device = int(os.environ.get("IWB_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    with _lock:
        eng = _engines.get(device)
        if eng is None:
            eng = _engines[device] = Engine(device)
        return eng
