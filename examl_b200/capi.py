"""ctypes binding of the C ABI declared in include/examl_b200.h.

`Engine` wraps libexaml_b200.so (the CUDA engine).  The class is written against a
(library, prefix) pair so that the test suite can bind its CPU checker (same signatures,
prefix `orc_`) through the very same code path -- the product itself never does:
`load_library()` below only ever opens the CUDA library and raises if it is missing.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from .phylo import Entry

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libexaml_b200.so")

EXPORTS = [
    "exb_create", "exb_destroy", "exb_last_error", "exb_set_partition", "exb_commit",
    "exb_set_model", "exb_newview", "exb_evaluate", "exb_makenewz_prepare", "exb_core", "exb_makenewz",
    "exb_comm_unique_id", "exb_comm_init", "exb_read_clv", "exb_read_scaler",
    "exb_read_sumtable", "exb_last_kernel_ms", "exb_launch_count", "exb_set_tile_sites", "exb_set_timing",
    "exb_set_rate_model", "exb_set_categories", "exb_set_model_lg4", "exb_evaluate_partial", "exb_comm_is_p2p",
    "exb_last_phase_ms", "exb_set_aa_precision", "exb_set_sev", "exb_sev_stats",
    "exb_score_insertions", "exb_can_score_insertions", "exb_max_entries",
]
SCORE = 4                 # EXB_SCORE: tipCase bit of a scoring entry
SCRATCH_NODES = 64        # EXB_SCRATCH_NODES
UNIQUE_ID_BYTES = 128


class ExbEntry(C.Structure):
    _fields_ = [("tipCase", C.c_int), ("pNumber", C.c_int), ("qNumber", C.c_int), ("rNumber", C.c_int)]


class EngineError(RuntimeError):
    pass


_lib_cache = {}


def load_library(path: str = LIB_PATH) -> C.CDLL:
    """Open the CUDA engine.  There is no fallback: a missing library is an error.  (EXB_LIBRARY: an alternative build of
    the same CUDA library, e.g. one compiled with other tuning macros -- experiments only.)"""
    path = os.environ.get("EXB_LIBRARY", path)
    if path in _lib_cache:
        return _lib_cache[path]
    if not os.path.exists(path):
        raise EngineError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  examl_b200 has no CPU fallback.")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    _lib_cache[path] = lib
    return lib


def _dptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def pack_traversal(ti: Sequence[Entry], num_branches: int):
    """Flatten a list of phylo.Entry into (exb_entry[count], qz[count*nb], rz[count*nb])."""
    n = len(ti)
    arr = (ExbEntry * max(n, 1))()
    qz = np.zeros(max(n, 1) * num_branches)
    rz = np.zeros(max(n, 1) * num_branches)
    for i, e in enumerate(ti):
        arr[i].tipCase, arr[i].pNumber, arr[i].qNumber, arr[i].rNumber = e.tipCase, e.pNumber, e.qNumber, e.rNumber
        qz[i * num_branches:(i + 1) * num_branches] = e.qz[:num_branches]
        if e.rz:
            rz[i * num_branches:(i + 1) * num_branches] = e.rz[:num_branches]
    return arr, qz, rz


class PackedTraversal:
    """A traversal descriptor already flattened to the C-ABI arrays (reusable across calls)."""
    __slots__ = ("arr", "qz", "rz", "count")

    def __init__(self, ti: Sequence[Entry], num_branches: int):
        self.arr, self.qz, self.rz = pack_traversal(ti, num_branches)
        self.count = len(ti)


class Engine:
    """One likelihood engine (one GPU).  Method names follow the reference functions they replace."""

    def __init__(self, mxtips: int, num_partitions: int, num_branches: int = 1, device: int = 0,
                 lib: Optional[C.CDLL] = None, prefix: str = "exb_", tile_sites: Optional[int] = None):
        self.lib = lib if lib is not None else load_library()
        self.prefix = prefix
        self.mxtips, self.nparts, self.nbranches = mxtips, num_partitions, num_branches
        self.states = [0] * num_partitions
        self.widths = [0] * num_partitions
        self.cat = False                  # CAT / PSR: one rate category per site (span = states)
        self._h = C.c_void_p()
        if prefix == "exb_":
            self._check(self._f("create")(C.byref(self._h), C.c_int(device), C.c_int(mxtips),
                                          C.c_int(num_partitions), C.c_int(num_branches)))
            if tile_sites is not None:
                self._check(self._f("set_tile_sites")(self._h, C.c_int(tile_sites)))
        else:
            self._check(self._f("create")(C.byref(self._h), C.c_int(mxtips), C.c_int(num_partitions),
                                          C.c_int(num_branches)))

    # ---- plumbing -------------------------------------------------------------------------
    def _f(self, name: str):
        fn = getattr(self.lib, self.prefix + name)
        fn.restype = C.c_int
        return fn

    def _check(self, rc: int) -> None:
        if rc != 0:
            msg = ""
            if self.prefix == "exb_":
                f = self.lib.exb_last_error
                f.restype = C.c_char_p
                msg = (f() or b"").decode()
            raise EngineError(f"{self.prefix} call failed (rc={rc}): {msg}")

    def close(self) -> None:
        if self._h:
            d = getattr(self.lib, self.prefix + "destroy")
            d.restype = None
            d(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- set-up ---------------------------------------------------------------------------
    def set_partition(self, model: int, states: int, wgt: np.ndarray, tips: np.ndarray) -> None:
        """tips: uint8 [mxtips, width] taxon-major (pInfo.yResource); wgt: int32 [width]."""
        wgt = np.ascontiguousarray(wgt, dtype=np.int32)
        tips = np.ascontiguousarray(tips, dtype=np.uint8)
        width = int(wgt.shape[0])
        assert tips.shape == (self.mxtips, width), (tips.shape, self.mxtips, width)
        self.states[model], self.widths[model] = states, width
        self._check(self._f("set_partition")(self._h, C.c_int(model), C.c_int(states), C.c_size_t(width),
                                             _dptr(wgt), _dptr(tips)))

    def set_rate_model(self, rate_het_model: int, max_categories: int = 25) -> None:
        """rateHetModel: 0 = CAT / PSR (one rate category per site), 1 = GAMMA (axml.h:349-350).  Before set_partition."""
        self.cat = rate_het_model == 0
        self._check(self._f("set_rate_model")(self._h, C.c_int(rate_het_model), C.c_int(max_categories)))

    def set_sev(self, on: bool = True) -> None:
        """-S: no CLV storage for (node, tile) pairs that are all-gap below the node.  Before set_partition / commit."""
        self._check(self._f("set_sev")(self._h, C.c_int(1 if on else 0)))

    def sev_stats(self):
        """(blocks in use, blocks a dense arena would hold, bytes of slabs allocated)."""
        a, b, c = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
        self._check(self._f("sev_stats")(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return int(a.value), int(b.value), int(c.value)

    def set_aa_precision(self, mode: int) -> None:
        """0 = FP64 (default), 1 = 3xTF32 on tcgen05 (mixed precision; include/examl_b200.h).  Before set_partition / commit."""
        self._check(self._f("set_aa_precision")(self._h, C.c_int(mode)))

    def set_categories(self, model: int, per_site_rates, rate_category) -> None:
        """CAT / PSR: pInfo.perSiteRates[numberOfCategories], pInfo.rateCategory[width]."""
        r = np.ascontiguousarray(per_site_rates, dtype=np.float64)
        rc = np.ascontiguousarray(rate_category, dtype=np.int32)
        assert rc.shape[0] == self.widths[model]
        self._check(self._f("set_categories")(self._h, C.c_int(model), C.c_int(r.shape[0]), _dptr(r), _dptr(rc)))

    def commit(self) -> None:
        if self.prefix == "exb_":
            self._check(self._f("commit")(self._h))

    def set_model(self, model: int, mc) -> None:
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in
                (mc.EIGN, mc.EV, mc.EI, mc.tipVector, mc.gammaRates)]
        self._check(self._f("set_model")(self._h, C.c_int(model), *[_dptr(a) for a in arrs]))

    def set_model_lg4(self, model: int, EIGN, EV, EI, tipVector, gammaRates, weights) -> None:
        """LG4M / LG4X: per-category eigen-systems, flattened [4][20], [4][400], [4][400], [4][460], and category weights."""
        arrs = [np.ascontiguousarray(a, dtype=np.float64).reshape(-1) for a in (EIGN, EV, EI, tipVector, gammaRates, weights)]
        assert [a.size for a in arrs] == [80, 1600, 1600, 1840, 4, 4]
        self._check(self._f("set_model_lg4")(self._h, C.c_int(model), *[_dptr(a) for a in arrs]))

    # ---- hot path ---------------------------------------------------------------------------
    @staticmethod
    def _mask(execute_model, n):
        if execute_model is None:
            return None
        return np.ascontiguousarray(np.asarray(execute_model, dtype=np.uint8).reshape(n))

    def pack(self, ti) -> PackedTraversal:
        return ti if isinstance(ti, PackedTraversal) else PackedTraversal(ti, self.nbranches)

    def newviewIterative(self, ti, execute_model=None) -> None:
        t = self.pack(ti)
        m = self._mask(execute_model, self.nparts)
        self._check(self._f("newview")(self._h, t.arr, C.c_int(t.count), _dptr(t.qz), _dptr(t.rz), _dptr(m)))

    def evaluateIterative(self, ti, execute_model=None, out: Optional[np.ndarray] = None) -> np.ndarray:
        t = self.pack(ti)
        m = self._mask(execute_model, self.nparts)
        lh = np.zeros(self.nparts) if out is None else out
        self._check(self._f("evaluate")(self._h, t.arr, C.c_int(t.count), _dptr(t.qz), _dptr(t.rz), _dptr(m), _dptr(lh)))
        return lh

    def makenewzIterative(self, ti, execute_model=None) -> None:
        t = self.pack(ti)
        m = self._mask(execute_model, self.nparts)
        self._check(self._f("makenewz_prepare")(self._h, t.arr, C.c_int(t.count), _dptr(t.qz), _dptr(t.rz), _dptr(m)))

    def execCore(self, lz: Sequence[float], execute_model=None):
        lzv = np.ascontiguousarray(np.asarray(lz, dtype=np.float64).reshape(self.nbranches))
        m = self._mask(execute_model, self.nparts)
        d1 = np.zeros(self.nbranches)
        d2 = np.zeros(self.nbranches)
        self._check(self._f("core")(self._h, _dptr(lzv), _dptr(m), _dptr(d1), _dptr(d2)))
        return d1, d2

    def makenewz(self, ti: Sequence[Entry], z0: Sequence[float], maxiter: int, execute_model=None):
        """topLevelMakenewz on the device (single rank): returns (z, curvatOK, iterations)."""
        t = self.pack(ti)
        arr, qz, rz = t.arr, t.qz, t.rz
        m = self._mask(execute_model, self.nparts)
        z0v = np.ascontiguousarray(np.asarray(z0, dtype=np.float64).reshape(self.nbranches))
        res = np.zeros(self.nbranches)
        cok = np.zeros(self.nbranches, dtype=np.int32)
        it = C.c_int(0)
        self._check(self._f("makenewz")(self._h, arr, C.c_int(t.count), _dptr(qz), _dptr(rz), _dptr(m), _dptr(z0v),
                                        C.c_int(maxiter), _dptr(res), _dptr(cok), C.byref(it)))
        return res, cok, int(it.value)

    def evaluatePartial(self, ti, models, sites, rates) -> np.ndarray:
        """Batched evaluatePartialGeneric (evaluatePartialGenericSpecial.c:259): for every (partition, site, rate)
        triple the site's log likelihood (times its pattern weight) with that single rate, recomputed from the
        tips along the full traversal `ti` (ti[0] = root branch whose pNumber is a tip).  CAT / PSR engines only."""
        t = self.pack(ti)
        mo = np.ascontiguousarray(models, dtype=np.int32)
        si = np.ascontiguousarray(sites, dtype=np.int32)
        ra = np.ascontiguousarray(rates, dtype=np.float64)
        assert mo.shape == si.shape == ra.shape
        out = np.zeros(ra.shape[0])
        self._check(self._f("evaluate_partial")(self._h, t.arr, C.c_int(t.count), _dptr(t.qz), _dptr(t.rz),
                                                C.c_size_t(ra.shape[0]), _dptr(mo), _dptr(si), _dptr(ra), _dptr(out)))
        return out

    def scoreInsertions(self, ti, ncand: int, execute_model=None) -> np.ndarray:
        """Batched lazy-SPR scoring (exb_score_insertions): ti[0] = the pruned subtree (pNumber, qz = its branch), then
        newview entries (scratch nodes 2*mxtips-1.. allowed) and scoring entries (tipCase | SCORE, pNumber = candidate).
        Returns lnl[ncand, partitions]."""
        t = self.pack(ti)
        m = self._mask(execute_model, self.nparts)
        out = np.zeros((ncand, self.nparts))
        self._check(self._f("score_insertions")(self._h, t.arr, C.c_int(t.count), _dptr(t.qz), _dptr(t.rz), _dptr(m),
                                                C.c_int(ncand), _dptr(out)))
        return out

    def max_entries(self) -> int:
        """Entries one descriptor may hold."""
        return 1 << 30 if self.prefix != "exb_" else int(self._f("max_entries")(self._h))

    def can_score_insertions(self) -> bool:
        return self.prefix != "exb_" or bool(self._f("can_score_insertions")(self._h))

    # ---- multi-GPU -------------------------------------------------------------------------
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(UNIQUE_ID_BYTES)
        self._check(self._f("comm_unique_id")(buf))
        return buf.raw

    def comm_init(self, rank: int, nranks: int, uid: bytes) -> None:
        buf = C.create_string_buffer(uid, UNIQUE_ID_BYTES)
        self._check(self._f("comm_init")(self._h, C.c_int(rank), C.c_int(nranks), buf))

    def comm_is_p2p(self) -> bool:
        return bool(self._f("comm_is_p2p")(self._h))

    # ---- introspection ------------------------------------------------------------------------
    def read_clv(self, model: int, node_number: int) -> np.ndarray:
        R = 1 if self.cat else 4
        out = np.zeros(self.widths[model] * R * self.states[model])
        self._check(self._f("read_clv")(self._h, C.c_int(model), C.c_int(node_number), _dptr(out)))
        return out.reshape(self.widths[model], R, self.states[model])

    def read_scaler(self, model: int, node_number: int) -> int:
        v = C.c_uint(0)
        self._check(self._f("read_scaler")(self._h, C.c_int(model), C.c_int(node_number), C.byref(v)))
        return int(v.value)

    def read_sumtable(self, model: int) -> np.ndarray:
        R = 1 if self.cat else 4
        out = np.zeros(self.widths[model] * R * self.states[model])
        self._check(self._f("read_sumtable")(self._h, C.c_int(model), _dptr(out)))
        return out.reshape(self.widths[model], R, self.states[model])

    def last_kernel_ms(self) -> float:
        v = C.c_float(0)
        self._check(self._f("last_kernel_ms")(self._h, C.byref(v)))
        return float(v.value)

    def last_phase_ms(self):
        """(tables, traverse, finish) milliseconds of the most recent traversal call (timing must be on)."""
        v = (C.c_float * 3)()
        self._check(self._f("last_phase_ms")(self._h, v))
        return float(v[0]), float(v[1]), float(v[2])

    def set_timing(self, on: bool) -> None:
        """CUDA events around the kernels of every hot-path call (last_kernel_ms); off on latency-critical paths."""
        self._check(self._f("set_timing")(self._h, C.c_int(1 if on else 0)))

    def launch_count(self) -> int:
        f = self.lib.exb_launch_count
        f.restype = C.c_long
        return int(f(self._h))
